set -x
export CEDR_B200_TRANSPOSED_MIN=1
timeout 600 python -m pytest tests/test_gpu_parity.py -m gpu -x -q -k "fast_path_blocks_bitwise or default_path_full_size or subtree_partition or replayed_as_cuda_graph" 2>&1 | tail -5
unset CEDR_B200_TRANSPOSED_MIN
python tools/kernel_times.py qlt ne120x128x40 1280
for g in 4 8 16 32; do CEDR_B200_GROUP=$g python tools/kernel_times.py qlt ne120x128x40 1280 | grep "down\|sum"; done
python tools/time_run.py qlt ne120x128x40 5120
CEDR_B200_TRANSPOSED=0 python tools/time_run.py qlt ne120x128x40 5120
