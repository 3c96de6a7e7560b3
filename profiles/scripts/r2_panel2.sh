mkdir -p gpurun_out
SLA_LDR_PIVOT=panel timeout 300 python profiles/blk_check.py 2>&1 | tail -8
C=stablelinearalgebra.jl_b200/csrc
nvcc -std=c++17 -O3 -lineinfo -gencode arch=compute_100a,code=sm_100a -Xcompiler -fPIC -Xcompiler -fvisibility=hidden -DSLA_PP_PROFILE -c $C/sla_ldr_pp.cu -o /tmp/pp_prof.o
OBJS=$(ls $C/build/*.o | grep -v "sla_ldr_pp.o")
nvcc -shared -o /tmp/libsla_ppprof.so $OBJS /tmp/pp_prof.o -gencode arch=compute_100a,code=sm_100a -Xcompiler -fPIC -cudart static
SLA_B200_LIB=/tmp/libsla_ppprof.so SLA_LDR_PIVOT=panel timeout 120 python profiles/ldr_bench_graded.py 256 64 100 2>&1 | tail -2 | cut -c1-700
SLA_B200_LIB=/tmp/libsla_ppprof.so SLA_LDR_PIVOT=panel timeout 120 python profiles/ldr_bench_graded.py 400 148 100 2>&1 | tail -2 | cut -c1-700
for a in "SLA_LDR_PIVOT=panel"; do
  env $a timeout 120 python profiles/ldr_bench_graded.py 256 64 100 2>&1 | tail -1
  env $a timeout 120 python profiles/ldr_bench_graded.py 400 512 100 2>&1 | tail -1
  env $a timeout 120 python profiles/ldr_bench.py 256 128 c 2>&1 | tail -1
done
timeout 1800 python -m pytest tests/test_gpu_parity.py -m gpu -x -q -k "panel_pivot_mode" 2>&1 | tail -12 | cut -c1-400
