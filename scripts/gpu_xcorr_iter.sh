python -m pytest tests -x -q -m gpu -k "cross or gtv or texture or golden" 2>&1 | tail -3
python scripts/bench_ops.py --shape brats --batch 16 --reps 5 --only xcorr --out gpurun_out/ops_xcorr.json > /dev/null 2>&1
python - <<P
import json
for r in json.load(open("gpurun_out/ops_xcorr.json")): print(r["op"], r["ms"], r["kernels_ms"])
P
mkdir -p gpurun_out/ncu; C="python scripts/bench_ops.py --reps 1 --shape brats --batch 16 --only rho5_k100_flair"
ncu --set full --clock-control none --import-source on -k regex:xcorr_tma -s 2 -c 1 -o gpurun_out/ncu/xcorr -f $C > gpurun_out/ncu/xcorr.log 2>&1; echo "ncu rc=$?"
