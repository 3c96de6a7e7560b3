set -x
mkdir -p gpurun_out
python -m pytest tests -m gpu -x -q -s 2>&1 | tail -40 > gpurun_out/r2_pytest1.log
tail -5 gpurun_out/r2_pytest1.log
for c in C2 C1 C3; do
  python bench.py --config $c > gpurun_out/r2_bench_$c.json 2> gpurun_out/r2_bench_$c.err; tail -c 600 gpurun_out/r2_bench_$c.json
done
python bench.py --config C5 --steps 3 > gpurun_out/r2_bench_C5.json 2> gpurun_out/r2_bench_C5.err; tail -c 600 gpurun_out/r2_bench_C5.json
python bench.py --config C4 --steps 2 > gpurun_out/r2_bench_C4.json 2> gpurun_out/r2_bench_C4.err; tail -c 600 gpurun_out/r2_bench_C4.json
