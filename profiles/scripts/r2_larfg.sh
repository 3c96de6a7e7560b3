mkdir -p gpurun_out
for a in "SLA_LDR_PIVOT=panel" "SLA_LDR_PIVOT=strict"; do
  env $a timeout 120 python profiles/ldr_bench_graded.py 256 64 100 2>&1 | tail -1
  env $a timeout 120 python profiles/ldr_bench_graded.py 400 512 100 2>&1 | tail -1
  env $a timeout 120 python profiles/ldr_bench.py 256 128 c 2>&1 | tail -1
  env $a timeout 200 python profiles/ldr_bench_graded.py 1024 32 100 2>&1 | tail -1
done
timeout 2400 python -m pytest tests -m gpu -x -q 2>&1 | tail -6 | cut -c1-400
