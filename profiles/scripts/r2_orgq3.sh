# orgq: split cluster barriers, 3-buffer ring with two CTA barriers per chunk
mkdir -p gpurun_out
for ng in 4 8; do
  SLA_ORGQ_NG=$ng timeout 120 python profiles/ldr_bench_graded.py 256 64 100 2>&1 | tail -1 | sed "s/^/ng=$ng /"
  SLA_ORGQ_NG=$ng timeout 120 python profiles/ldr_bench_graded.py 136 64 100 2>&1 | tail -1 | sed "s/^/ng=$ng /"
done
timeout 900 python -m pytest tests/test_gpu_parity.py -q -x -m gpu -k "ldr_hybrid_shape or chain_C2 or ldr_random or ldr_graded or sweep_small or chain_graph or forced_algo" 2>&1 | tail -3
SLA_ORGQ_NG=8 timeout 600 python -m pytest tests/test_gpu_parity.py -q -x -m gpu -k "ldr_hybrid_shape or chain_C2_sample" 2>&1 | tail -3
b() { env "$@" timeout 300 python bench.py --config C2 --steps 10 --warmup 3 --no-cpu 2>/dev/null | python -c "import sys,json; d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print('$*', d['ms_per_step'], d['value'], d['gpu_launches'])"; }
b SLA_ORGQ_NG=4
b SLA_ORGQ_NG=8
