# Round-2 scaling runs on N GPUs of one box (run under `gpurun --gpus N`): weak (c2 per GPU; c3 at N = 8) and c5 strong.
N=$1
T="python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29517"
mkdir -p gpurun_out
timeout 300 $T bench.py --gpus $N --steps 20 --warmup 3 --no-cpu-baseline --no-gpu-reference 2>gpurun_out/r02_final_n${N}_weak.err | tail -1 > gpurun_out/r02_final_n${N}_weak.json
timeout 300 $T bench.py --gpus $N --steps 20 --warmup 3 --P 500000 --views 8 --res 1024 --phase raster --strong --no-cpu-baseline --no-gpu-reference 2>gpurun_out/r02_final_c5_n${N}.err | tail -1 > gpurun_out/r02_final_c5_n${N}.json
python - <<PY
import json
for f in ["gpurun_out/r02_final_n${N}_weak.json", "gpurun_out/r02_final_c5_n${N}.json"]:
    d = json.loads(open(f).read()); print(f, round(d["value"], 2), round(d["ms_per_step"], 3), d["phase_ms"], d.get("allreduce_us"))
PY
