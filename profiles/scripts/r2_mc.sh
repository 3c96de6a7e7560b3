# multi-CTA streaming LDR kernel: correctness (default dispatch at n > 512 + forced child runs), timing against streaming
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_parity.py -m gpu -x -q -k "ldr_n512_and_n1024" 2>&1 | tail -8
for a in "SLA_LDR_MC=0" "SLA_LDR_MC=1" "SLA_LDR_ALGO=mc SLA_LDR_MC_CS=2" "SLA_LDR_ALGO=mc SLA_LDR_MC_CS=8"; do
  echo "== $a"; env $a timeout 300 python profiles/ldr_bench_graded.py 1024 32 100 2>&1 | tail -1
done
env SLA_LDR_MC=0 timeout 300 python profiles/ldr_bench_graded.py 1024 4 100 2>&1 | tail -1
timeout 300 python profiles/ldr_bench_graded.py 1024 4 100 2>&1 | tail -1
timeout 1200 python -m pytest tests/test_gpu_parity.py -m gpu -x -q -k "other_ldr_kernels_subprocess and mc or chain_C4 or sweep_C2" 2>&1 | tail -8
timeout 900 python bench.py --config C4 --steps 2 --no-cpu > gpurun_out/r2c_bench_C4.json 2> gpurun_out/r2c_bench_C4.err; tail -c 700 gpurun_out/r2c_bench_C4.json; tail -2 gpurun_out/r2c_bench_C4.err
