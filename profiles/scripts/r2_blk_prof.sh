# blocked LDR kernel: correctness for every n, then a -DSLA_BLK_PROFILE build of the same sources for the phase counters
mkdir -p gpurun_out
SLA_LDR_ALGO=blk timeout 300 python profiles/blk_check.py 2>&1 | tail -8
C=stablelinearalgebra.jl_b200/csrc
nvcc -std=c++17 -O3 -lineinfo -gencode arch=compute_100a,code=sm_100a -Xcompiler -fPIC -Xcompiler -fvisibility=hidden -DSLA_BLK_PROFILE -c $C/sla_ldr_blk.cu -o /tmp/blk_prof.o
OBJS=$(ls $C/build/*.o | grep -v sla_ldr_blk.o)
nvcc -shared -o /tmp/libsla_prof.so $OBJS /tmp/blk_prof.o -gencode arch=compute_100a,code=sm_100a -Xcompiler -fPIC -cudart static
for rep in 1 2; do
SLA_B200_LIB=/tmp/libsla_prof.so SLA_LDR_ALGO=blk timeout 120 python profiles/blk_phase_cycles.py 256 64 100
done
SLA_B200_LIB=/tmp/libsla_prof.so SLA_LDR_ALGO=blk timeout 120 python profiles/blk_phase_cycles.py 256 2 100
SLA_B200_LIB=/tmp/libsla_prof.so SLA_LDR_ALGO=blk timeout 120 python profiles/blk_phase_cycles.py 256 64 0
