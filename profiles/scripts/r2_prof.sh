# round-2 evidence pass: remaining tests, sweep bench line, launch list of a C2 step, full captures of the changed kernels
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_parity.py -m gpu -x -q -s -k "sweep_C2 or tma_kernel" 2>&1 | grep -v "^$" | tail -6 | cut -c1-400
timeout 600 python bench.py --workload sweep --config C2 > gpurun_out/r2_sweep_C2.json 2> gpurun_out/r2_sweep_C2.err; tail -c 1200 gpurun_out/r2_sweep_C2.json; tail -3 gpurun_out/r2_sweep_C2.err
python profiles/run_step.py C2 > gpurun_out/plain_r2.log 2>&1 &&
ncu --metrics gpu__time_duration.sum --clock-control none -s 200 -c 200 --csv --log-file gpurun_out/launches_r2_C2_step.csv python profiles/run_step.py C2 > gpurun_out/ncu_r2_a.log 2>&1
python profiles/inv_bench.py 256 64 > gpurun_out/plain_r2_lu.log 2>&1 &&
ncu --set full --clock-control none --import-source on -k regex:lu_kernel -s 4 -c 1 -o gpurun_out/prof_lu_r2 python profiles/inv_bench.py 256 64 > gpurun_out/ncu_r2_b.log 2>&1
python profiles/gemm_bench.py 256 128 s > gpurun_out/plain_r2_gemm.log 2>&1 &&
ncu --set full --clock-control none --import-source on -k regex:gemm_kernel -s 3 -c 1 -o gpurun_out/prof_gemm_r2 python profiles/gemm_bench.py 256 128 s > gpurun_out/ncu_r2_c.log 2>&1
python profiles/ldr_bench_graded.py 1024 32 100 > gpurun_out/plain_r2_mc.log 2>&1 &&
ncu --set full --clock-control none --import-source on -k regex:ldr_mc_kernel -s 1 -c 1 -o gpurun_out/prof_ldrmc_r2 python profiles/ldr_bench_graded.py 1024 32 100 > gpurun_out/ncu_r2_d.log 2>&1
python profiles/ldr_bench_graded.py 256 64 100 > gpurun_out/plain_r2_hyb.log 2>&1 &&
ncu --set full --clock-control none --import-source on -k regex:ldr_reg_kernel -s 2 -c 1 -o gpurun_out/prof_ldrhyb_r2 python profiles/ldr_bench_graded.py 256 64 100 > gpurun_out/ncu_r2_e.log 2>&1
ls -la gpurun_out/*.ncu-rep gpurun_out/launches_r2_C2_step.csv 2>&1 | tail -8
