# LU: column-oriented substitutions
mkdir -p gpurun_out
timeout 120 python profiles/inv_bench.py 256 64 2>&1 | tail -4
timeout 120 python profiles/inv_bench.py 256 128 c 2>&1 | tail -4
timeout 120 python profiles/inv_bench.py 400 512 2>&1 | tail -4
timeout 900 python -m pytest tests/test_gpu_parity.py -q -x -m gpu -k "lu or inv or ldiv or rdiv or chain_C1 or chain_C2_sample or chain_small or sweep_small or f32_ldr" 2>&1 | tail -3
b() { env "$@" timeout 300 python bench.py --config C2 --steps 10 --warmup 3 --no-cpu 2>/dev/null | python -c "import sys,json; d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print('$*', d['ms_per_step'], d['value'], d['gpu_launches'])"; }
b SLA_X=1
