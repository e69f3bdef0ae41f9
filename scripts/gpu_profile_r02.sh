#!/bin/bash
# Round 2, one gpurun call: parity tests, the bench line (both arms), the ncu launch list of the bench
# command, full ncu captures of the kernels of the GTV step and the per-operator tables.
# Outputs go to gpurun_out/ (scratch); scripts/collect_profiles.py r02 copies the summaries to profiles/.
mkdir -p gpurun_out gpurun_out/ncu
rm -f gpurun_out/ncu/*
python -m pytest tests -x -q -m gpu > gpurun_out/pytest_gpu.log 2>&1; echo "pytest rc=$?" >> gpurun_out/pytest_gpu.log
tail -3 gpurun_out/pytest_gpu.log
python bench.py --steps 20 --warmup 5 > gpurun_out/bench_default.json 2> gpurun_out/bench_default.err; echo "bench rc=$?"
python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/bench_reference.json 2>> gpurun_out/bench_default.err; echo "reference rc=$?"
python bench.py --batch 8 --no-cpu-baseline > gpurun_out/bench_b8.json 2>> gpurun_out/bench_default.err
CMD="python bench.py --steps 2 --warmup 1 --batch 8 --no-cpu-baseline --threads 1"
$CMD > gpurun_out/plain.log 2>&1 &&
ncu --metrics gpu__time_duration.sum --clock-control none -c 2500 --csv --log-file gpurun_out/launches.csv $CMD > gpurun_out/ncu_list.log 2>&1
echo "launch list rc=$?"
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
python scripts/bench_ops.py --shape brats --batch 16 --reps 5 --out gpurun_out/ops_brats_b16.json > /dev/null 2>&1; echo "ops rc=$?"
python scripts/bench_ops.py --shape 512 --batch 1 --reps 5 --out gpurun_out/ops_512.json > /dev/null 2>&1; echo "ops512 rc=$?"
python scripts/bench_texture.py --batch 4 --steps 5 --out gpurun_out/texture_b4.json > /dev/null 2>&1; echo "texture rc=$?"
ls gpurun_out/ncu | head -60
