#!/bin/bash
# multi-GPU call: parity tests (incl. the NCCL slab test), scaling bench, slab timing
N=${1:-2}
mkdir -p gpurun_out
python -m pytest tests -x -q -m gpu 2>&1 | tail -5
python bench.py --gpus 1 --no-cpu-baseline > gpurun_out/scale_n1.json 2> gpurun_out/scale.err
python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus $N > gpurun_out/scale_n$N.json 2>> gpurun_out/scale.err
python - <<'P'
import json,glob
for f in sorted(glob.glob("gpurun_out/scale_n*.json")):
    try:
        d=json.loads(open(f).read().strip().splitlines()[-1]); print(f, d["n_gpus"], round(d["value"],1), round(d["e2e"]["value"],1), d["ms_per_step"])
    except Exception as e: print(f, "ERR", e)
P
tail -5 gpurun_out/scale.err
python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29512 scripts/slab_gpu.py --size 512 > gpurun_out/slab_512_n$N.json 2> gpurun_out/slab.err; tail -2 gpurun_out/slab_512_n$N.json; tail -3 gpurun_out/slab.err
python scripts/bench_ops.py --shape 512 --batch 1 --reps 3 --only edt,ccl26,through26 --out gpurun_out/ops_512.json 2>&1 | cut -c1-400
