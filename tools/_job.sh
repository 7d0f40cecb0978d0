CEDR_B200_TRANSPOSED_MIN=1 timeout 600 python -m pytest tests/test_gpu_parity.py -m gpu -x -q -k "fast_path_blocks_bitwise or default_path_full_size or subtree_partition or replayed_as_cuda_graph" 2>&1 | tail -3
python tools/kernel_times.py qlt ne120x128x40 1280 | grep "mid\|down\|sum"
for gy in 4 10 20 40; do echo gy=$gy; CEDR_B200_MID_GY=$gy python tools/kernel_times.py qlt ne120x128x40 1280 | grep "mid"; done
python tools/kernel_times.py qlt ne120x128x40 5120 | grep "mid\|sum"
python tools/time_run.py qlt ne30x72x40
