mkdir -p gpurun_out
timeout 300 python -m pytest tests/test_gpu_parity.py -m gpu -x -q -k "gemm" 2>&1 | tail -6 | cut -c1-300
for b in 128 64 1280; do
SLA_GEMM_SA=0 timeout 120 python profiles/gemm_bench.py 256 $b s 2>&1 | tail -2
SLA_GEMM_SA=1 timeout 120 python profiles/gemm_bench.py 256 $b s 2>&1 | head -1
done
SLA_LDR_ALGO=mc timeout 600 python -m pytest tests/test_gpu_parity.py -m gpu -x -q -k "ldr_random or ldr_graded or ldr_hybrid or ldr_identity" 2>&1 | tail -4 | cut -c1-300
timeout 1200 python -m pytest tests/test_gpu_parity.py -m gpu -x -q -k "chain_C2 or chain_C1 or sweep_C2 or chain_C4" 2>&1 | tail -4 | cut -c1-300
timeout 600 python bench.py --config C2 --no-cpu > gpurun_out/r2d_bench_C2.json 2> gpurun_out/r2d_bench_C2.err; tail -c 400 gpurun_out/r2d_bench_C2.json; tail -2 gpurun_out/r2d_bench_C2.err
