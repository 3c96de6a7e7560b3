# C2 weak scaling of the final build on the GPUs of this box (N = number of visible GPUs, and N = 1 on the same box)
mkdir -p gpurun_out
NG=$(python -c "import torch; print(torch.cuda.device_count())")
timeout 300 python bench.py --gpus 1 --steps 20 --warmup 3 --no-cpu > gpurun_out/r2f_scale_C2_weak_1of$NG.json 2> gpurun_out/r2f_scale_1.err
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $NG --master-addr 127.0.0.1 --master-port $((29700 + NG)) \
  bench.py --gpus $NG --config C2 --steps 20 --warmup 3 --no-cpu > gpurun_out/r2f_scale_C2_weak_$NG.json 2> gpurun_out/r2f_scale_$NG.err
for f in gpurun_out/r2f_scale_C2_weak_1of$NG.json gpurun_out/r2f_scale_C2_weak_$NG.json; do
python - <<P
import json
d=json.loads(open("$f").read().strip().splitlines()[-1])
print("$f", d["n_gpus"], d["value"], d["ms_per_step"], "e2e", d["e2e"]["value"], d["e2e"].get("ms_per_step"), d.get("clocks"))
P
done
tail -2 gpurun_out/r2f_scale_$NG.err | cut -c1-300
