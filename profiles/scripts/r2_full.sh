# full GPU test-suite + C2 / C4 bench lines of the current build
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -x -q -s 2>&1 | tail -40 > gpurun_out/r2_pytest2.log
tail -8 gpurun_out/r2_pytest2.log
timeout 600 python bench.py --config C2 --no-cpu > gpurun_out/r2b_bench_C2.json 2> gpurun_out/r2b_bench_C2.err; tail -c 900 gpurun_out/r2b_bench_C2.json
timeout 900 python bench.py --config C4 --steps 2 --no-cpu > gpurun_out/r2b_bench_C4.json 2> gpurun_out/r2b_bench_C4.err; tail -c 900 gpurun_out/r2b_bench_C4.json
