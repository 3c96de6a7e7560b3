# diagnostics build (-DSLA_LDR_INTERVALS): how long does ldr! (factorisation + Q) take on the main stream inside the step?
for v in "SLA_OVERLAP=0" "SLA_SIDES=1 SLA_OVERLAP_SUB=2" "SLA_SIDES=2 SLA_OVERLAP_SUB=1"; do
  env $v SLA_GRAPH=0 timeout 300 python bench.py --config C2 --steps 10 --warmup 3 --no-cpu --no-parity 2>gpurun_out/iv.err | python -c "import sys,json; d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print('$v', d['ms_per_step'])"
  grep "intervals" gpurun_out/iv.err | tail -2
done
