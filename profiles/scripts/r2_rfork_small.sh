# R product of a stabilisation on its own stream for small problems (C1): chain parity tests, C1 bench with and without
mkdir -p gpurun_out
timeout 300 python -m pytest tests -m gpu -x -q -k "chain_C1 or chain_few or small_complex or graph_replay or f32 or c32_chain or golden or sharding_invariance or sweep_small or sweep_C1 or tiny_sizes or c_client" 2>&1 | tail -25 > gpurun_out/r2r_pytest.log
tail -4 gpurun_out/r2r_pytest.log
timeout 200 python bench.py --config C1 > gpurun_out/r2r_C1.json 2> gpurun_out/r2r_C1.err; tail -c 400 gpurun_out/r2r_C1.err
SLA_RFORK=0 timeout 100 python bench.py --config C1 --no-cpu > gpurun_out/r2r_C1_off.json 2> gpurun_out/r2r_C1_off.err
python - <<'P'
import json
for f in ['r2r_C1','r2r_C1_off']:
    d=json.loads(open('gpurun_out/%s.json'%f).read().strip().splitlines()[-1])
    print(f, d['value'], d['ms_per_step'], d['e2e']['value'], d['gpu_launches'], d['parity'], d['clocks'])
P
