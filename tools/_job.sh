run () {
timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 8 --steps 10 --warmup 3 --no-e2e --no-cpu-baseline > gpurun_out/tmp_n8.json 2> gpurun_out/tmp_n8.err
python - <<'PY'
import json
for l in open("gpurun_out/tmp_n8.json"):
    if l.startswith("{"):
        d=json.loads(l); print(d["n_gpus"], round(d["ms_per_step"],4), round(d["qlt"]["ms_per_run"],4), round(d["caas"]["ms_per_run"],4), d["output_digest"]["qlt"])
PY
}
echo default; run
echo no_side; CEDR_B200_NO_SIDE_STREAM=1 run
echo down2; CEDR_B200_TRANSPOSED=0 run
