mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_parity.py tests/test_abi.py -m gpu -x -q -k "graph_replay or c_client or chain_C1 or chain_C2_sample or ldr_n512 or chain_C4_size_short" 2>&1 | tail -6 | cut -c1-300
for c in C1 C2; do
SLA_GRAPH=0 timeout 600 python bench.py --config $c --no-cpu > gpurun_out/r2e_bench_${c}_nograph.json 2> gpurun_out/r2e_bench_${c}_nograph.err; python -c "
import json,sys
d=json.loads([l for l in open('gpurun_out/r2e_bench_${c}_nograph.json') if l.startswith('{')][0]); print('$c nograph', d['value'], d['ms_per_step'], d['e2e']['value'], d['gpu_launches'])"
timeout 600 python bench.py --config $c --no-cpu > gpurun_out/r2e_bench_${c}.json 2> gpurun_out/r2e_bench_${c}.err; python -c "
import json,sys
d=json.loads([l for l in open('gpurun_out/r2e_bench_${c}.json') if l.startswith('{')][0]); print('$c graph  ', d['value'], d['ms_per_step'], d['e2e']['value'], d['gpu_launches'])"
tail -2 gpurun_out/r2e_bench_${c}.err | cut -c1-300
done
timeout 600 python bench.py --config C5 --steps 3 --no-cpu > gpurun_out/r2e_bench_C5.json 2> gpurun_out/r2e_bench_C5.err; python -c "
import json,sys
d=json.loads([l for l in open('gpurun_out/r2e_bench_C5.json') if l.startswith('{')][0]); print('C5', d['value'], d['ms_per_step'], d['e2e']['value'], d['roofline_ldr']['launch_ms'])"
timeout 600 python bench.py --config C3 --steps 5 --no-cpu > gpurun_out/r2e_bench_C3.json 2> gpurun_out/r2e_bench_C3.err; python -c "
import json,sys
d=json.loads([l for l in open('gpurun_out/r2e_bench_C3.json') if l.startswith('{')][0]); print('C3', d['value'], d['ms_per_step'], d['e2e']['value'], d['roofline_ldr']['launch_ms'])"
