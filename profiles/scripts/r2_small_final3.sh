# small LU inverse, third version (pivot search over the published column with all 32 lanes of the warp): parity tests, timings, C1 bench
mkdir -p gpurun_out
timeout 400 python -m pytest tests -m gpu -x -q -k "lu_small or det_inv_lu or chain_C1 or chain_few or small_complex or graph_replay or sweep_small or sweep_C1 or f32 or test_ref_ or tiny_sizes or golden or info_survives or lu_singular or ldiv_lu or c32_chain or c_client" 2>&1 | tail -25 > gpurun_out/r2q_pytest.log
tail -8 gpurun_out/r2q_pytest.log
timeout 60 python profiles/lu_small_bench.py > gpurun_out/r2q_lu_on.log 2>&1; cat gpurun_out/r2q_lu_on.log
timeout 200 python bench.py --config C1 > gpurun_out/r2q_C1.json 2> gpurun_out/r2q_C1.err; tail -c 400 gpurun_out/r2q_C1.err; python - <<'P'
import json
d=json.loads(open('gpurun_out/r2q_C1.json').read().strip().splitlines()[-1])
print('C1', d['value'], d['ms_per_step'], d['e2e']['value'], d['ldr_kernel'], d['parity'], d['clocks'])
P
