# round-2 evidence pass after the GEMM fast loader / orgq changes: launch list of a C2 step, full captures of gemm (fast) and orgq
mkdir -p gpurun_out
python profiles/run_step.py C2 > gpurun_out/plain_r2g.log 2>&1 &&
ncu --metrics gpu__time_duration.sum --clock-control none -s 198 -c 198 --csv --log-file gpurun_out/launches_r2g_C2_step.csv python profiles/run_step.py C2 > gpurun_out/ncu_r2g_a.log 2>&1
python profiles/gemm_bench.py 256 128 s > gpurun_out/plain_r2g_gemm.log 2>&1 &&
ncu --set full --clock-control none --import-source on -k regex:gemm_kernel -s 3 -c 1 -o gpurun_out/prof_gemm_r2g python profiles/gemm_bench.py 256 128 s > gpurun_out/ncu_r2g_c.log 2>&1
python profiles/ldr_bench_graded.py 256 64 100 > gpurun_out/plain_r2g_hyb.log 2>&1 &&
ncu --set full --clock-control none --import-source on -k regex:orgq_dmma -s 2 -c 1 -o gpurun_out/prof_orgq_r2g python profiles/ldr_bench_graded.py 256 64 100 > gpurun_out/ncu_r2g_e.log 2>&1
ls -la gpurun_out/*r2g* 2>&1 | tail -8
