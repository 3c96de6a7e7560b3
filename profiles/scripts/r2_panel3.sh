mkdir -p gpurun_out
C=stablelinearalgebra.jl_b200/csrc
nvcc -std=c++17 -O3 -lineinfo -gencode arch=compute_100a,code=sm_100a -Xcompiler -fPIC -Xcompiler -fvisibility=hidden -DSLA_PP_PROFILE -c $C/sla_ldr_pp.cu -o /tmp/pp_prof.o
OBJS=$(ls $C/build/*.o | grep -v "sla_ldr_pp.o")
nvcc -shared -o /tmp/libsla_ppprof.so $OBJS /tmp/pp_prof.o -gencode arch=compute_100a,code=sm_100a -Xcompiler -fPIC -cudart static
SLA_B200_LIB=/tmp/libsla_ppprof.so SLA_LDR_PIVOT=panel timeout 120 python profiles/ldr_bench_graded.py 256 64 100 2>&1 | tail -2 | cut -c1-700
SLA_B200_LIB=/tmp/libsla_ppprof.so SLA_LDR_PIVOT=panel timeout 120 python profiles/ldr_bench.py 256 128 c 2>&1 | tail -2 | cut -c1-700
for a in "SLA_LDR_PIVOT=panel"; do
  env $a timeout 120 python profiles/ldr_bench_graded.py 256 64 100 2>&1 | tail -1
  env $a timeout 120 python profiles/ldr_bench_graded.py 400 512 100 2>&1 | tail -1
  env $a timeout 120 python profiles/ldr_bench.py 256 128 c 2>&1 | tail -1
done
timeout 1800 python -m pytest tests/test_gpu_parity.py -m gpu -x -q -k "panel_pivot_mode or (other_ldr_kernels_subprocess and (streaming or mc)) or ldr_n512" 2>&1 | tail -6 | cut -c1-400
for c in C2 C3 C5; do
timeout 900 python bench.py --config $c --ldr-pivoting panel --no-cpu $( [ $c = C5 ] && echo --steps 3 ) > gpurun_out/r2_panel_$c.json 2> gpurun_out/r2_panel_$c.err; python -c "
import json
d=json.loads([l for l in open('gpurun_out/r2_panel_$c.json') if l.startswith('{')][0]); print('$c panel', d['value'], d['ms_per_step'], d['e2e']['value'], d['ldr_kernel'], d['roofline_ldr']['launch_ms'])"
done
