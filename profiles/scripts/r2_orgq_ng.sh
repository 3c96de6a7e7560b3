# A/B: orgq with 32 columns per CTA, clusters of 8, two CTAs per SM (SLA_ORGQ_NG=4) vs 64 columns, clusters of 4; GEMM stage / sub-chunk variants on the C2 step
mkdir -p gpurun_out
for ng in 8 4; do
  SLA_ORGQ_NG=$ng timeout 120 python profiles/ldr_bench_graded.py 256 64 100 2>&1 | tail -1 | sed "s/^/ng=$ng /"
  SLA_ORGQ_NG=$ng timeout 120 python profiles/ldr_bench_graded.py 200 64 100 2>&1 | tail -1 | sed "s/^/ng=$ng /"
done
SLA_ORGQ_NG=4 timeout 600 python -m pytest tests/test_gpu_parity.py -q -x -m gpu -k "ldr_hybrid_shape or chain_C2_sample or chain_C2_full or ldr_random or ldr_graded" 2>&1 | tail -3
b() { env "$@" timeout 300 python bench.py --config C2 --steps 10 --warmup 3 --no-cpu 2>/dev/null | python -c "import sys,json; d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print('$*', d['ms_per_step'], d['value'])"; }
b SLA_ORGQ_NG=8
b SLA_ORGQ_NG=4
b SLA_GEMM_TILE=6424
b SLA_OVERLAP_SUB=4
b SLA_OVERLAP_SUB=5
b SLA_OVERLAP_SUB=1
b SLA_ORGQ_NG=4 SLA_GEMM_TILE=6424
