# final pass of the small-matrix kernels (LDR version 3 as the default, small LU inverse): parity tests at n <= 64 without any
# environment override, LU timings old/new, full C1 bench (CPU leg + parity), smoke
mkdir -p gpurun_out
timeout 400 python -m pytest tests -m gpu -x -q -k "lu_small or det_inv_lu or ldr_small or ldr_random or ldr_graded or ldr_identity or chain_C1 or chain_few or small_complex or graph_replay or sweep_small or sweep_C1 or f32 or test_ref_ or tiny_sizes or range_guard or golden or info_survives or lu_singular or ldiv_lu or c32_chain or sharding_invariance or forced_algo or forced_ldr or c_client or unsupported_size or walker_sums" 2>&1 | tail -25 > gpurun_out/r2x_pytest.log
tail -8 gpurun_out/r2x_pytest.log
timeout 60 python profiles/lu_small_bench.py > gpurun_out/r2x_lu_on.log 2>&1; cat gpurun_out/r2x_lu_on.log
SLA_LU_SMALL=0 timeout 60 python profiles/lu_small_bench.py > gpurun_out/r2x_lu_off.log 2>&1; cat gpurun_out/r2x_lu_off.log
timeout 200 python bench.py --config C1 > gpurun_out/r2x_C1.json 2> gpurun_out/r2x_C1.err; tail -c 400 gpurun_out/r2x_C1.err; python - <<'P'
import json
d=json.loads(open('gpurun_out/r2x_C1.json').read().strip().splitlines()[-1])
print('C1', d['value'], d['ms_per_step'], d['e2e']['value'], d['ldr_kernel'], d['parity'], d['clocks'])
P
timeout 120 python -c "import __graft_entry__ as g; g.smoke(); print('smoke ok')" 2>&1 | tail -3
