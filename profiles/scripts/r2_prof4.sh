# evidence pass of the final build: launch list of a C2 step (289 launches), full captures of the pipelined LU inverse and of one batch-64 GEMM launch
mkdir -p gpurun_out
python profiles/run_step.py C2 > gpurun_out/plain_r2h.log 2>&1 &&
ncu --metrics gpu__time_duration.sum --clock-control none -s 289 -c 289 --csv --log-file gpurun_out/launches_r2h_C2_step.csv python profiles/run_step.py C2 > gpurun_out/ncu_r2h_a.log 2>&1
python profiles/inv_bench.py 256 64 > gpurun_out/plain_r2h_lu.log 2>&1 &&
ncu --set full --clock-control none --import-source on -k regex:lu_kernel -s 4 -c 1 -o gpurun_out/prof_lu_r2h python profiles/inv_bench.py 256 64 > gpurun_out/ncu_r2h_b.log 2>&1
python profiles/gemm_bench.py 256 64 s > gpurun_out/plain_r2h_gemm.log 2>&1 &&
ncu --set full --clock-control none --import-source on -k regex:gemm_kernel -s 3 -c 1 -o gpurun_out/prof_gemm64_r2h python profiles/gemm_bench.py 256 64 s > gpurun_out/ncu_r2h_c.log 2>&1
ls -la gpurun_out/*r2h* 2>&1 | tail -8
