mkdir -p gpurun_out
b() { env "$@" timeout 300 python bench.py --config C2 --steps 10 --warmup 3 --no-cpu 2>/dev/null | python -c "import sys,json; d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print('$*', d['ms_per_step'], d['value'], d['gpu_launches'])"; }
b SLA_SIDES=2 SLA_OVERLAP_SUB=1
b SLA_SIDES=3 SLA_OVERLAP_SUB=1
b SLA_SIDES=4 SLA_OVERLAP_SUB=1
b SLA_SIDES=3 SLA_OVERLAP_SUB=2
b SLA_SIDES=1 SLA_OVERLAP_SUB=2
b SLA_SIDES=2 SLA_OVERLAP_SUB=1
timeout 900 python -m pytest tests/test_gpu_parity.py -q -x -m gpu -k "chain_C1 or chain_C2_sample or chain_small or chain_graph" 2>&1 | tail -3
