C=stablelinearalgebra.jl_b200/csrc
nvcc -std=c++17 -O3 -lineinfo -gencode arch=compute_100a,code=sm_100a -Xcompiler -fPIC -Xcompiler -fvisibility=hidden -DSLA_PP_PROFILE -c $C/sla_ldr_pp.cu -o /tmp/pp_prof.o
OBJS=$(ls $C/build/*.o | grep -v "sla_ldr_pp.o")
nvcc -shared -o /tmp/libsla_ppprof.so $OBJS /tmp/pp_prof.o -gencode arch=compute_100a,code=sm_100a -Xcompiler -fPIC -cudart static
SLA_B200_LIB=/tmp/libsla_ppprof.so SLA_LDR_PIVOT=panel timeout 120 python profiles/ldr_bench_graded.py 256 64 100 2>&1 | tail -2 | cut -c1-900
SLA_B200_LIB=/tmp/libsla_ppprof.so SLA_LDR_PIVOT=panel timeout 120 python profiles/ldr_bench_graded.py 256 1 100 2>&1 | tail -2 | cut -c1-900
