// flame/utils/assert.h -- FLAME_ASSERT (/root/reference/src/flame_nodelet.cc:85-88, src/flame_offline_tum.cc:691-694)
#pragma once
#include <cstdio>
#include <cstdlib>
#define FLAME_ASSERT(cond)                                                                  \
  do {                                                                                      \
    if (!(cond)) {                                                                          \
      std::fprintf(stderr, "FLAME_ASSERT failed: %s (%s:%d)\n", #cond, __FILE__, __LINE__); \
      std::abort();                                                                         \
    }                                                                                       \
  } while (0)
