set -e
cd "$(dirname "$0")/.."
python tools/prof_run.py qlt ne120x128x40 1 640 > gpurun_out/prof_plain_qlt.log 2>&1 &&
ncu --set full --clock-control none --import-source on -f -o gpurun_out/r02_qlt_ne120 python tools/prof_run.py qlt ne120x128x40 1 640 > gpurun_out/prof_ncu_qlt.log 2>&1
python tools/prof_run.py caas ne120x128x40 1 640 > gpurun_out/prof_plain_caas.log 2>&1 &&
ncu --set full --clock-control none --import-source on -f -o gpurun_out/r02_caas_ne120 python tools/prof_run.py caas ne120x128x40 1 640 > gpurun_out/prof_ncu_caas.log 2>&1
python bench.py --steps 2 --warmup 3 --no-e2e --no-cpu-baseline > gpurun_out/prof_plain_bench.log 2>&1 &&
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/r02_launches_bench_ne120.csv python bench.py --steps 2 --warmup 3 --no-e2e --no-cpu-baseline > gpurun_out/prof_ncu_bench.log 2>&1
ls -la gpurun_out/r02_*
