# A/B: group-product sub-chunks alternating between two side streams
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_parity.py -q -x -m gpu -k "chain_C1 or chain_C2 or chain_small or chain_C3_sample or chain_graph or sharding or chain_C5_size" 2>&1 | tail -3
b() { env "$@" timeout 300 python bench.py --config C2 --steps 10 --warmup 3 --no-cpu 2>/dev/null | python -c "import sys,json; d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print('$*', d['ms_per_step'], d['value'], d['gpu_launches'])"; }
b SLA_SIDE2=1
b SLA_SIDE2=0
b SLA_SIDE2=1 SLA_OVERLAP_SUB=1
b SLA_SIDE2=1 SLA_OVERLAP_SUB=3
b SLA_SIDE2=1 SLA_GRAPH=0
b SLA_SIDE2=0 SLA_GRAPH=0
