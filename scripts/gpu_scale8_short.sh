#!/bin/bash
# 8-GPU call, trimmed: scaling bench at 2/4/8 ranks (N=1 comes from the 1-GPU profile run) and the
# 2048^3 slab-decomposed volume with peer halo planes.
mkdir -p gpurun_out
nvidia-smi -L | wc -l
for N in 8 4 2; do
  python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 2951$N bench.py --gpus $N --no-cpu-baseline > gpurun_out/scale8_n$N.json 2>> gpurun_out/scale8.err
done
python - <<'P'
import json,glob
for f in sorted(glob.glob("gpurun_out/scale8_n[248].json")):
    try:
        d=json.loads(open(f).read().strip().splitlines()[-1]); print(f, d["n_gpus"], round(d["value"],1), round(d["e2e"]["value"],1), d["ms_per_step"])
    except Exception as e: print(f, "ERR", e)
P
tail -3 gpurun_out/scale8.err
python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29523 scripts/slab_gpu.py --size 128 --check --peer 2>&1 | grep -E "SLAB|Error|error" | head -5
python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29522 scripts/slab_gpu.py --size 2048 --reps 2 --peer > gpurun_out/slab_2048_n8.json 2> gpurun_out/slab8.err; tail -1 gpurun_out/slab_2048_n8.json; tail -3 gpurun_out/slab8.err
