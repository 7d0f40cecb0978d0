set -x
CEDR_B200_TRANSPOSED_MIN=1 timeout 600 python -m pytest tests/test_multigpu.py -m gpu -q -rA 2>&1 | tail -9 > gpurun_out/r02b_pytest_multigpu_2gpu.log
tail -8 gpurun_out/r02b_pytest_multigpu_2gpu.log
timeout 600 python -m pytest tests/test_gpu_parity.py -m gpu -q -x -k "partition or rank_map or run_replayed" 2>&1 | tail -3
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 --steps 5 --warmup 3 --no-e2e --no-cpu-baseline > gpurun_out/r02b_bench_ne120_n2_noe2e.json 2> gpurun_out/r02b_bench_n2.err
python - <<'PY'
import json
for l in open("gpurun_out/r02b_bench_ne120_n2_noe2e.json"):
    if l.startswith("{"):
        d=json.loads(l); print(d["ms_per_step"], d["qlt"]["ms_per_run"], d["caas"]["ms_per_run"], d["output_digest"])
PY
