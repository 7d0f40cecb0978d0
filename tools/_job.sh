set -x
python bench.py --steps 5 --warmup 3 > gpurun_out/r02b_bench_ne120_n1.json 2> gpurun_out/r02b_bench_ne120_n1.err
tail -c 600 gpurun_out/r02b_bench_ne120_n1.json
python bench.py --steps 5 --warmup 3 --workload ne30x72x40 > gpurun_out/r02b_bench_ne30_n1.json 2>/dev/null
python bench.py --steps 5 --warmup 3 --workload ne256x128x10 --no-cpu-baseline > gpurun_out/r02b_bench_ne256_n1.json 2>/dev/null
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/r02b_launches_bench_ne120.csv python bench.py --steps 2 --warmup 1 --no-e2e --no-cpu-baseline > gpurun_out/ncu_bench.log 2>&1
ncu --set full --clock-control none --import-source on -c 12 -o gpurun_out/r02b_qlt_ne120 python tools/prof_run.py qlt ne120x128x40 1 640 > gpurun_out/ncu_q.log 2>&1
tail -2 gpurun_out/ncu_q.log
python tools/mixed_classes.py > gpurun_out/r02b_mixed_classes.txt 2>&1; tail -5 gpurun_out/r02b_mixed_classes.txt
