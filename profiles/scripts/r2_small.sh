# small-matrix LDR kernel (sla_ldr_small.cu): parity tests at n <= 64, A/B timing against the previous dispatch, C1 bench
mkdir -p gpurun_out
timeout 240 python -m pytest tests/test_gpu_parity.py -m gpu -x -q -k "ldr_small or ldr_random or ldr_graded or ldr_identity or chain_C1 or chain_few or small_complex or graph_replay or sweep_small or sweep_C1 or f32 or test_ref_" 2>&1 | tail -12 > gpurun_out/r2s_pytest.log
tail -6 gpurun_out/r2s_pytest.log
timeout 90 python profiles/ldr_small_bench.py > gpurun_out/r2s_ldr_on.log 2>&1; cat gpurun_out/r2s_ldr_on.log | tail -10
SLA_LDR_SMALL=0 timeout 90 python profiles/ldr_small_bench.py > gpurun_out/r2s_ldr_off.log 2>&1; cat gpurun_out/r2s_ldr_off.log | tail -10
timeout 120 python bench.py --config C1 --no-cpu > gpurun_out/r2s_C1.json 2> gpurun_out/r2s_C1.err; tail -c 900 gpurun_out/r2s_C1.json
SLA_LDR_SMALL=0 timeout 120 python bench.py --config C1 --no-cpu > gpurun_out/r2s_C1_off.json 2> gpurun_out/r2s_C1_off.err; tail -c 300 gpurun_out/r2s_C1_off.json
