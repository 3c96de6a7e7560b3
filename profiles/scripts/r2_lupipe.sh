# LU inverse: forward sweep by the cluster peers during the factorisation (PIPE) vs after it
mkdir -p gpurun_out
for v in 1 0; do
  SLA_LU_PIPE=$v timeout 60 python profiles/inv_bench.py 256 64 2>&1 | tail -3 | sed "s/^/pipe=$v /"
done
SLA_LU_PIPE=1 timeout 120 python profiles/inv_bench.py 1024 32 2>&1 | tail -3 | sed "s/^/pipe=1 /"
SLA_LU_PIPE=0 timeout 120 python profiles/inv_bench.py 1024 32 2>&1 | tail -3 | sed "s/^/pipe=0 /"
timeout 600 python -m pytest tests/test_gpu_parity.py -q -x -m gpu -k "lu or inv or ldiv or rdiv or chain_C1 or chain_C2_sample or chain_small or sweep_small or f32_ldr or tiny or n512" 2>&1 | tail -3
b() { env "$@" timeout 300 python bench.py --config C2 --steps 10 --warmup 3 --no-cpu 2>/dev/null | python -c "import sys,json; d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print('$*', d['ms_per_step'], d['value'], d['gpu_launches'])"; }
b SLA_LU_PIPE=1
b SLA_LU_PIPE=0
