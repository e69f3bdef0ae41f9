#!/bin/bash
# One gpurun call: parity tests, the bench line, the ncu launch list of the bench command
# and one full ncu capture of the dominant kernel.  Outputs go to gpurun_out/.
set -x
mkdir -p gpurun_out
python -m pytest tests -x -q -m gpu > gpurun_out/pytest_gpu.log 2>&1; echo "pytest rc=$?" >> gpurun_out/pytest_gpu.log
tail -3 gpurun_out/pytest_gpu.log
python bench.py > gpurun_out/bench_default.json 2> gpurun_out/bench_default.err; echo "bench rc=$?"
python bench.py --batch 8 --no-cpu-baseline > gpurun_out/bench_b8.json 2>> gpurun_out/bench_default.err
CMD="python bench.py --steps 2 --warmup 1 --batch 8 --no-cpu-baseline"
$CMD > gpurun_out/plain.log 2>&1 &&
ncu --metrics gpu__time_duration.sum --clock-control none -c 1500 --csv --log-file gpurun_out/launches.csv $CMD > gpurun_out/ncu_list.log 2>&1
CMD2="python scripts/bench_ops.py --shape brats --batch 8 --reps 1 --only xcorr"
$CMD2 > gpurun_out/plain2.log 2>&1 &&
ncu --set full --clock-control none --import-source on -k regex:xcorr_ring -s 2 -c 1 -o gpurun_out/xcorr_full -f $CMD2 > gpurun_out/ncu_full.log 2>&1
python scripts/bench_ops.py --shape brats --batch 16 --reps 5 --out gpurun_out/ops_brats_b16.json > /dev/null 2>&1
python scripts/bench_ops.py --shape 512 --batch 1 --reps 5 --out gpurun_out/ops_512.json > /dev/null 2>&1
ls -la gpurun_out
