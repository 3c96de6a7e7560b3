# small-matrix LDR kernel, version 3 (speculative ?larfg under the pivot search, fused norms, ballot shortcut): parity tests at
# n <= 64 with the version forced through the environment, timings, phase counters, C1 bench, launch list of the C1 step
mkdir -p gpurun_out
export SLA_LDR_SMALL=3
timeout 300 python -m pytest tests -m gpu -x -q -k "ldr_small or ldr_random or ldr_graded or ldr_identity or chain_C1 or chain_few or small_complex or graph_replay or sweep_small or sweep_C1 or f32 or test_ref_ or tiny_sizes or range_guard or golden or info_survives or c32_chain or sharding_invariance or forced_algo or forced_ldr or c_client or unsupported_size or walker_sums" 2>&1 | tail -25 > gpurun_out/r2w_pytest.log
tail -8 gpurun_out/r2w_pytest.log
timeout 90 python profiles/ldr_small_bench.py > gpurun_out/r2w_ldr_v3.log 2>&1; tail -10 gpurun_out/r2w_ldr_v3.log
SLA_B200_LIB=$PWD/stablelinearalgebra.jl_b200/csrc/abl/libsla_b200_smallprof.so timeout 50 python profiles/ldr_small_phases.py 64 1 > gpurun_out/r2w_phases.log 2>&1; cat gpurun_out/r2w_phases.log
timeout 120 python bench.py --config C1 --no-cpu > gpurun_out/r2w_C1.json 2> gpurun_out/r2w_C1.err; tail -c 400 gpurun_out/r2w_C1.err; python - <<'P'
import json
d=json.loads(open('gpurun_out/r2w_C1.json').read().strip().splitlines()[-1])
print('C1', d['value'], d['ms_per_step'], d['e2e']['value'], d['ldr_kernel'], d['clocks'])
P
