# flameConfig.cmake -- what `find_package(flame REQUIRED)` of robustrobotics/flame_ros resolves to
# (/root/reference/CMakeLists.txt:57) when the B200 library stands in for the flame core.
# Exports the two variables the reference's build uses: flame_INCLUDE_DIRS
# (/root/reference/CMakeLists.txt:205, include_directories) and flame_LIBRARIES
# (/root/reference/src/CMakeLists.txt:11,34,58, target_link_libraries).
#
# Layout expected (source tree or install prefix):
#   <root>/include/flame/*.h, <root>/include/flame_b200.h
#   <root>/flame_ros_b200/lib/libflame_b200.so   (source tree)   or   <root>/lib/libflame_b200.so (installed)
# Use:  catkin build flame_ros -Dflame_DIR=<root>/cmake
get_filename_component(_flame_root "${CMAKE_CURRENT_LIST_DIR}/.." ABSOLUTE)
set(flame_INCLUDE_DIRS "${_flame_root}/include")
find_library(flame_LIBRARY NAMES flame_b200
             PATHS "${_flame_root}/flame_ros_b200/lib" "${_flame_root}/lib" NO_DEFAULT_PATH)
if(NOT flame_LIBRARY)
  set(flame_FOUND FALSE)
  if(flame_FIND_REQUIRED)
    message(FATAL_ERROR "flameConfig.cmake: libflame_b200.so not found under ${_flame_root} (build it: python -m flame_ros_b200.build)")
  endif()
else()
  set(flame_LIBRARIES "${flame_LIBRARY}")
  set(flame_FOUND TRUE)
  if(NOT TARGET flame::flame)
    add_library(flame::flame SHARED IMPORTED)
    set_target_properties(flame::flame PROPERTIES
      IMPORTED_LOCATION "${flame_LIBRARY}"
      INTERFACE_INCLUDE_DIRECTORIES "${flame_INCLUDE_DIRS}")
  endif()
endif()
unset(_flame_root)
