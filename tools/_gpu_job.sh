set -x
timeout 900 python -m pytest tests/test_gpu_parity.py -x -q -m gpu -k "one_field or fast_path_blocks or subtree_partition or randomized or general_rank or ring_mixed" 2>&1 | tail -5
timeout 300 python tools/mixed_classes.py 86400 1280 > gpurun_out/mixed_classes.txt 2>&1; tail -5 gpurun_out/mixed_classes.txt
CEDR_B200_RING_TRACE=1 timeout 300 python tools/ring_trace.py caas 86400 640 > gpurun_out/ring_trace_caas.txt 2>&1; tail -3 gpurun_out/ring_trace_caas.txt
CEDR_B200_RING_TRACE=1 timeout 300 python tools/ring_trace.py qlt 86400 640 > gpurun_out/ring_trace_qlt.txt 2>&1; tail -3 gpurun_out/ring_trace_qlt.txt
