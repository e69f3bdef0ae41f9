#!/bin/bash
mkdir -p gpurun_out
C1="python scripts/bench_ops.py --shape 512 --batch 1 --reps 1 --only bernoulli"
C2="python scripts/bench_ops.py --shape 512 --batch 1 --reps 1 --only edt"
C3="python scripts/bench_ops.py --shape brats --batch 8 --reps 1 --only percentiles"
$C1 > gpurun_out/p1.log 2>&1 && $C2 > gpurun_out/p2.log 2>&1 && $C3 > gpurun_out/p3.log 2>&1 &&
ncu --set full --clock-control none --import-source on -k regex:ccl_merge -s 6 -c 1 -o gpurun_out/ccl_merge_full -f $C1 > gpurun_out/n1.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:edt_env -s 4 -c 2 -o gpurun_out/edt_env_full -f $C2 > gpurun_out/n2.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:rs_scatter -s 8 -c 1 -o gpurun_out/rs_scatter_full -f $C3 > gpurun_out/n3.log 2>&1
tail -2 gpurun_out/n1.log gpurun_out/n2.log gpurun_out/n3.log; cat gpurun_out/p3.log | cut -c1-300
