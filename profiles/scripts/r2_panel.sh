mkdir -p gpurun_out
SLA_LDR_PIVOT=panel timeout 300 python profiles/blk_check.py 2>&1 | tail -8
timeout 1800 python -m pytest tests/test_gpu_parity.py -m gpu -x -q -k "panel_pivot_mode" 2>&1 | tail -25 | cut -c1-400
for a in "SLA_LDR_PIVOT=strict" "SLA_LDR_PIVOT=panel"; do
  env $a timeout 120 python profiles/ldr_bench_graded.py 256 64 100 2>&1 | tail -1
  env $a timeout 120 python profiles/ldr_bench_graded.py 400 512 100 2>&1 | tail -1
  env $a timeout 120 python profiles/ldr_bench.py 256 128 c 2>&1 | tail -1
  env $a timeout 200 python profiles/ldr_bench_graded.py 1024 32 100 2>&1 | tail -1
done
timeout 600 python bench.py --config C2 --ldr-pivoting panel > gpurun_out/r2_panel_C2.json 2> gpurun_out/r2_panel_C2.err; python -c "
import json
d=json.loads([l for l in open('gpurun_out/r2_panel_C2.json') if l.startswith('{')][0]); print('C2 panel', d['value'], d['ms_per_step'], d['e2e']['value'], d['parity'], d['ldr_kernel'])"; tail -2 gpurun_out/r2_panel_C2.err | cut -c1-300
