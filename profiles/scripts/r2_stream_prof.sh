# phase counters of the L2-streaming LDR kernel (one CTA per matrix) at the sizes of C2 / C3 / C5
C=stablelinearalgebra.jl_b200/csrc
nvcc -std=c++17 -O3 -lineinfo -gencode arch=compute_100a,code=sm_100a -Xcompiler -fPIC -Xcompiler -fvisibility=hidden -DSLA_LDRS_PROFILE -c $C/sla_ldr.cu -o /tmp/ldrs_prof.o
OBJS=$(ls $C/build/*.o | grep -v "sla_ldr.o")
nvcc -shared -o /tmp/libsla_sprof.so $OBJS /tmp/ldrs_prof.o -gencode arch=compute_100a,code=sm_100a -Xcompiler -fPIC -cudart static
export SLA_B200_LIB=/tmp/libsla_sprof.so SLA_LDR_ALGO=streaming
timeout 120 python profiles/ldr_bench_graded.py 256 64 100 2>&1 | tail -2
timeout 120 python profiles/ldr_bench_graded.py 400 148 100 2>&1 | tail -2
timeout 120 python profiles/ldr_bench.py 256 128 c 2>&1 | tail -2
