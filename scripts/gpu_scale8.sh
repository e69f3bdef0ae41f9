#!/bin/bash
# 8-GPU call: scaling bench 1/2/4/8, slab check at N=8, 2048^3 slab timing
mkdir -p gpurun_out
nvidia-smi -L | head -8
for N in 1 2 4 8; do
  if [ $N = 1 ]; then python bench.py --gpus 1 --no-cpu-baseline > gpurun_out/scale8_n1.json 2> gpurun_out/scale8.err
  else python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 2951$N bench.py --gpus $N > gpurun_out/scale8_n$N.json 2>> gpurun_out/scale8.err; fi
done
python - <<'P'
import json,glob
for f in sorted(glob.glob("gpurun_out/scale8_n*.json")):
    try:
        d=json.loads(open(f).read().strip().splitlines()[-1]); print(f, d["n_gpus"], round(d["value"],1), round(d["e2e"]["value"],1), d["ms_per_step"])
    except Exception as e: print(f, "ERR", e)
P
tail -3 gpurun_out/scale8.err
python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29521 scripts/slab_gpu.py --size 128 --check 2>&1 | grep -E "SLAB|Error|error" | head
python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29523 scripts/slab_gpu.py --size 128 --check --peer 2>&1 | grep -E "SLAB|Error|error" | head
python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29522 scripts/slab_gpu.py --size 2048 --reps 2 --peer > gpurun_out/slab_2048_n8.json 2> gpurun_out/slab8.err; tail -1 gpurun_out/slab_2048_n8.json; tail -3 gpurun_out/slab8.err
python scripts/batch256.py > gpurun_out/batch256_n1.json 2>> gpurun_out/scale8.err; tail -1 gpurun_out/batch256_n1.json
python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29524 scripts/batch256.py > gpurun_out/batch256_n8.json 2>> gpurun_out/scale8.err; tail -1 gpurun_out/batch256_n8.json
grep -h numa gpurun_out/scale8_n8.json | head -c 0; python - <<'P'
import json
d=json.loads(open("gpurun_out/scale8_n8.json").read().strip().splitlines()[-1]); print("numa:", d["e2e"].get("numa"))
P
