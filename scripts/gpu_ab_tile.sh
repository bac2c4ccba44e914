# A/B of the tile solver inside fb_update: the working-tree kernel vs scratch_old_tile.cuh (the previous version).
timeout 300 python -m pytest tests/test_gpu_nltgv2.py tests/test_gpu_update.py -q -x 2>&1 | tail -2
for i in 1 2 3; do timeout 120 python scripts/profile_update.py 8 0 2>&1 | cut -c1-120; done
cp flame_ros_b200/csrc/nltgv2_tile.cuh /tmp/new_tile.cuh
cp scratch_old_tile.cuh flame_ros_b200/csrc/nltgv2_tile.cuh
python -m flame_ros_b200.build > /dev/null 2>&1 && echo "rebuilt with the previous tile kernel"
for i in 1 2 3; do timeout 120 python scripts/profile_update.py 8 0 2>&1 | cut -c1-120; done
