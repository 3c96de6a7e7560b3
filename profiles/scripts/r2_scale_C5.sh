# C5 (N=400, L=200) over 2/4/8 GPUs of one box: weak (512 chains per GPU) and strong (4096 chains in total)
mkdir -p gpurun_out
for N in 2 4 8; do
  for S in weak strong; do
    timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port $((29500 + N)) \
      bench.py --gpus $N --config C5 --scaling $S --steps 2 --warmup 3 --no-cpu > gpurun_out/r2_scale_C5_${S}_$N.json 2> gpurun_out/r2_scale_C5_${S}_$N.err
    tail -c 300 gpurun_out/r2_scale_C5_${S}_$N.json; echo; tail -2 gpurun_out/r2_scale_C5_${S}_$N.err | cut -c1-300
  done
done
for N in 2 8; do
  timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port $((29600 + N)) \
    bench.py --gpus $N --config C2 --steps 20 --warmup 3 --no-cpu > gpurun_out/r2_scale_C2_weak_$N.json 2> gpurun_out/r2_scale_C2_weak_$N.err
  tail -c 200 gpurun_out/r2_scale_C2_weak_$N.json; echo
done
