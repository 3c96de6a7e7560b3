# correctness + timing of the blocked LDR kernel (sla_ldr_blk.cu) against the hybrid shape, one box
mkdir -p gpurun_out
SLA_LDR_ALGO=blk timeout 300 python profiles/blk_check.py 2>&1 | tail -12
SLA_LDR_ALGO=hyb timeout 300 python profiles/blk_check.py 2>&1 | tail -12
for rep in 1 2; do
SLA_LDR_ALGO=hyb timeout 120 python profiles/ldr_bench_graded.py 256 64 100 2>&1 | tail -1
SLA_LDR_ALGO=blk timeout 120 python profiles/ldr_bench_graded.py 256 64 100 2>&1 | tail -1
done
SLA_LDR_ALGO=blk timeout 120 python profiles/ldr_bench_graded.py 200 64 100 2>&1 | tail -1
SLA_LDR_ALGO=hyb timeout 120 python profiles/ldr_bench_graded.py 200 64 100 2>&1 | tail -1
SLA_LDR_ALGO=blk timeout 120 python profiles/ldr_bench_graded.py 256 16 100 2>&1 | tail -1
SLA_LDR_ALGO=reg timeout 120 python profiles/ldr_bench_graded.py 256 16 100 2>&1 | tail -1
timeout 600 python -m pytest tests/test_gpu_parity.py tests/test_abi.py -m gpu -x -q -k "ldr_hybrid or ldr_random or ldr_graded or chain_C2 or norm_range or c_client" 2>&1 | tail -15
timeout 300 python bench.py --config C2 --no-cpu > gpurun_out/r2_blk_C2.json 2> gpurun_out/r2_blk_C2.err; tail -c 1500 gpurun_out/r2_blk_C2.json; tail -3 gpurun_out/r2_blk_C2.err
