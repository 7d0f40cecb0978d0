cd "$(dirname "$0")/.."
python bench.py --steps 5 --warmup 3 > gpurun_out/bench_r02_ne120.json 2> gpurun_out/bench_r02_ne120.err
python bench.py --steps 5 --warmup 3 --workload ne30x72x40 > gpurun_out/bench_r02_ne30.json 2> gpurun_out/bench_r02_ne30.err
python bench.py --impl reference --steps 2 --warmup 1 > gpurun_out/bench_r02_ne120_reference.json 2> gpurun_out/bench_r02_ne120_reference.err
tail -c 600 gpurun_out/bench_r02_ne30.json; echo; tail -2 gpurun_out/bench_r02_ne120.err
