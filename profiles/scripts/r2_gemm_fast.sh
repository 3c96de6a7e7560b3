# A/B of the precomputed-address loader (FAST) of gemm_kernel, one B200
mkdir -p gpurun_out
for b in 64 128 1280; do
  for f in 0 1; do
    SLA_GEMM_SA=0 SLA_GEMM_FAST=$f timeout 120 python profiles/gemm_bench.py 256 $b s 2>&1 | head -1 | sed "s/^/fast=$f /"
  done
done
for t in 12864 128 6432 6444 6424; do
  for b in 128 1280; do
    SLA_GEMM_SA=0 SLA_GEMM_TILE=$t timeout 120 python profiles/gemm_bench.py 256 $b s 2>&1 | head -1 | sed "s/^/fast=1 /"
  done
done
for f in 0 1; do
  SLA_GEMM_SA=0 SLA_GEMM_FAST=$f timeout 120 python profiles/gemm_bench.py 400 512 2>&1 | sed "s/^/fast=$f /"
  SLA_GEMM_SA=0 SLA_GEMM_FAST=$f timeout 120 python profiles/gemm_bench.py 256 256 c 2>&1 | sed "s/^/fast=$f /"
  SLA_GEMM_SA=0 SLA_GEMM_FAST=$f timeout 120 python profiles/gemm_bench.py 1024 64 2>&1 | sed "s/^/fast=$f /"
done
timeout 600 python -m pytest tests/test_gpu_parity.py -q -x -m gpu -k "gemm or chain_C1 or small_complex or chain_C2_sample" 2>&1 | tail -3
for f in 0 1; do
  SLA_GEMM_FAST=$f timeout 300 python bench.py --config C2 --steps 10 --warmup 3 --no-cpu 2>gpurun_out/r2f_bench_C2_fast$f.err | tee gpurun_out/r2f_bench_C2_fast$f.json | python -c "import sys,json; d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print('fast=$f C2', d['ms_per_step'], d['value'], d['roofline']['achieved'], d['parity'])"
done
timeout 900 python -m pytest tests/test_gpu_parity.py -q -x -m gpu -k "panel_pivot_mode_subprocess" 2>&1 | tail -15 | cut -c1-300
