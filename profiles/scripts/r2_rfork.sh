# A/B: R0*Rv of a stabilisation on its own stream beside the Q formation
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_parity.py -q -x -m gpu -k "chain_C1 or chain_C2 or chain_small or chain_C3_sample or chain_graph or sharding or chain_C5_size or f32_chain or c32_chain" 2>&1 | tail -3
b() { env "$@" timeout 300 python bench.py --config C2 --steps 10 --warmup 3 --no-cpu 2>/dev/null | python -c "import sys,json; d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print('$*', d['ms_per_step'], d['value'], d['gpu_launches'])"; }
b SLA_RFORK=1
b SLA_RFORK=0
b SLA_RFORK=1
b SLA_RFORK=1 SLA_GRAPH=0
