#!/bin/bash
# One ncu --set full capture per kernel family (the plain command runs first, as the recipe asks).
mkdir -p gpurun_out/ncu
run() {  # name, kernel regex, skip, bench_ops args...
  local name=$1 regex=$2 skip=$3; shift 3
  local CMD="python scripts/bench_ops.py --reps 1 $*"
  $CMD > gpurun_out/ncu/$name.plain.log 2>&1 &&
  ncu --set full --clock-control none --import-source on -k regex:$regex -s $skip -c 1 -o gpurun_out/ncu/$name -f $CMD > gpurun_out/ncu/$name.log 2>&1
  echo "$name rc=$?"
}
python -m pytest tests -x -q -m gpu 2>&1 | tail -3
run pointwise_threshold pointwise_program 6 --shape brats --batch 16 --only threshold
run morph16 morph16 2 --shape brats --batch 16 --only near26
run reduce_f32 reduce_f32 4 --shape brats --batch 16 --only max
run ccl_init ccl_init 6 --shape brats --batch 16 --only through26
run ccl_merge ccl_merge 6 --shape brats --batch 16 --only through26
run ccl_select ccl_select 6 --shape brats --batch 16 --only through26
run ball_or ball_or 2 --shape brats --batch 16 --only distlt
run rs_scatter rs_scatter 8 --shape brats --batch 16 --only percentiles
run pc_rank pc_rank 2 --shape brats --batch 16 --only percentiles
run edt_env edt_env 4 --shape 512 --batch 1 --only edt
run xcorr xcorr_ring 2 --shape brats --batch 16 --only xcorr
python scripts/bench_ops.py --shape 512 --batch 1 --reps 3 --only edt 2>&1 | cut -c1-300
ls gpurun_out/ncu | head -40
