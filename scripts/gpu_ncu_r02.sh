#!/bin/bash
# the ncu --set full captures of scripts/gpu_profile_r02.sh alone
mkdir -p gpurun_out/ncu
run() {  # name, kernel regex, skip, script args...
  local name=$1 regex=$2 skip=$3; shift 3
  local C="python $*"
  $C > gpurun_out/ncu/$name.plain.log 2>&1 &&
  ncu --set full --clock-control none --import-source on -k regex:$regex -s $skip -c 1 -o gpurun_out/ncu/$name -f $C > gpurun_out/ncu/$name.log 2>&1
  echo "$name rc=$?"
}
OPS="scripts/bench_ops.py --reps 1 --shape brats --batch 16"
run xcorr xcorr_tma 2 $OPS --only rho5_k100_flair
run ccl_merge ccl_merge 2 $OPS --only through26_bg
run ccl_nodes ccl_nodes 2 $OPS --only through26_bg
run ccl_seed ccl_seed 2 $OPS --only through26_bg
run ccl_select ccl_select 2 $OPS --only through26_bg
run ball_or ball_or 2 $OPS --only distlt_r=5
run pack_bits pack_bits 2 $OPS --only distlt_r=5
run morph_bits morph_bits 2 $OPS --only near26
run boolean_program boolean_program 2 $OPS --only and_u8
run pointwise_threshold pointwise_program 2 $OPS --only threshold
run code_compare code_compare 2 $OPS --only percentiles_then_cmp_16-bit
run pq_hist pq_hist 2 $OPS --only percentiles_brain_16-bit
run pq_apply pq_apply 2 $OPS --only percentiles_brain_16-bit
run xc_bin xc_bin 2 $OPS --only rho5_k100_16-bit
run rs_scatter rs_scatter 8 $OPS --only percentiles_brain_f32
run pc_rank pc_rank 2 $OPS --only percentiles_brain_f32
run edt_window edt_window_kernel 1 scripts/bench_ops.py --reps 1 --shape 512 --batch 1 --only edt_blobs
run edt_env edt_env 1 scripts/bench_ops.py --reps 1 --shape 512 --batch 1 --only edt_sparse
ls gpurun_out/ncu | head -60
