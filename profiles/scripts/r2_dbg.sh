mkdir -p gpurun_out
SLA_LDR_PIVOT=panel SLA_EXPECT_ALGO=panel timeout 1500 python -m pytest tests/test_gpu_parity.py -q -x -m gpu -k "ldr_random or ldr_graded or ldr_identity or ldr_hybrid_shape or ldr_n512 or test_ref_ or chain_C1 or small_complex or chain_C2_sample or chain_C3_sample or chain_C5_size_sample or chain_C4_size_short or sweep_small or sweep_C1 or f32_chain or forced_algo_ran" 2>&1 | grep -v "^$" | tail -30 | cut -c1-300
timeout 1200 python -m pytest tests/test_gpu_parity.py -q -x -m gpu -k "other_ldr_kernels_subprocess and (streaming or mc) or ldr_n512 or chain_C3_full or chain_C5_per or chain_C4_full" 2>&1 | tail -5 | cut -c1-300
for a in "SLA_LDR_PIVOT=strict"; do
  env $a timeout 120 python profiles/ldr_bench_graded.py 400 512 100 2>&1 | tail -1
  env $a timeout 120 python profiles/ldr_bench.py 256 128 c 2>&1 | tail -1
  env $a timeout 200 python profiles/ldr_bench_graded.py 1024 32 100 2>&1 | tail -1
done
