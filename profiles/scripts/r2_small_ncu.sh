# ncu --set full of the small-matrix LDR (version 3) and LU-inverse kernels, batch 64 of 64 x 64; C2 bench line of the final build
mkdir -p gpurun_out
timeout 60 python profiles/small_kernels_once.py > gpurun_out/r2z_plain.log 2>&1 && timeout 150 ncu --set full --clock-control none --import-source on -k regex:small --launch-skip 2 --launch-count 2 -f -o gpurun_out/prof_small_r2z python profiles/small_kernels_once.py > gpurun_out/r2z_ncu.log 2>&1; tail -2 gpurun_out/r2z_plain.log gpurun_out/r2z_ncu.log
timeout 200 python bench.py --config C2 --no-cpu > gpurun_out/r2z_C2.json 2> gpurun_out/r2z_C2.err; tail -c 300 gpurun_out/r2z_C2.err; python - <<'P'
import json
d=json.loads(open('gpurun_out/r2z_C2.json').read().strip().splitlines()[-1])
print('C2', d['value'], d['ms_per_step'], d['e2e']['value'], d['ldr_kernel'], d['roofline']['frac'], d['clocks'])
P
