#!/bin/bash
# 4-GPU call: what limits the end-to-end number when every rank uploads at once -- host threads x
# pinned-memory flavour (plain / write-combined), ordered uploads on (the default)
mkdir -p gpurun_out
nvidia-smi -L | wc -l
N=${1:-4}
i=0
for cfg in "2" "3" "2 --wc" "3 --wc"; do
  set -- $cfg
  i=$((i+1))
  python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 2961$i bench.py --gpus $N --steps 10 --warmup 3 --no-cpu-baseline --no-slab --threads $1 $2 > gpurun_out/scale${N}_e2e_t$1${2:+_wc}.json 2>> gpurun_out/scale4.err
done
python - <<'P'
import json,glob
for f in sorted(glob.glob("gpurun_out/scale*_e2e_t*.json")):
    try:
        d=json.loads(open(f).read().strip().splitlines()[-1]); print(f, d["n_gpus"], round(d["value"],1), round(d["e2e"]["value"],1), round(d["e2e"]["ms_per_step"],2), d["e2e"]["h2d_gbs_this_rank_all_ranks_copying"], d["e2e"]["numa"])
    except Exception as e: print(f, "ERR", e)
P
tail -3 gpurun_out/scale4.err
