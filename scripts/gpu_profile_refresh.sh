#!/bin/bash
# After a change to the crossCorrelation / comparison kernels only: parity tests, the bench line, the launch
# list and the two ncu captures that changed.  scripts/collect_profiles.py r02 then refreshes profiles/.
mkdir -p gpurun_out gpurun_out/ncu
python -m pytest tests -x -q -m gpu > gpurun_out/pytest_gpu.log 2>&1; echo "pytest rc=$?" >> gpurun_out/pytest_gpu.log
tail -3 gpurun_out/pytest_gpu.log
python bench.py --steps 20 --warmup 5 > gpurun_out/bench_default.json 2> gpurun_out/bench_default.err; echo "bench rc=$?"
python bench.py --batch 8 --no-cpu-baseline > gpurun_out/bench_b8.json 2>> gpurun_out/bench_default.err
CMD="python bench.py --steps 2 --warmup 1 --batch 8 --no-cpu-baseline --threads 1"
$CMD > gpurun_out/plain.log 2>&1 &&
ncu --metrics gpu__time_duration.sum --clock-control none -c 2500 --csv --log-file gpurun_out/launches.csv $CMD > gpurun_out/ncu_list.log 2>&1
echo "launch list rc=$?"
run() {
  local name=$1 regex=$2 skip=$3; shift 3
  local C="python scripts/bench_ops.py --reps 1 $*"
  $C > gpurun_out/ncu/$name.plain.log 2>&1 &&
  ncu --set full --clock-control none --import-source on -k regex:$regex -s $skip -c 1 -o gpurun_out/ncu/$name -f $C > gpurun_out/ncu/$name.log 2>&1
  echo "$name rc=$?"
}
run xcorr xcorr_tma 2 --shape brats --batch 16 --only rho5_k100_flair
run code_compare code_compare 2 --shape brats --batch 16 --only percentiles_then_cmp_16-bit
python scripts/bench_ops.py --shape brats --batch 16 --reps 5 --out gpurun_out/ops_brats_b16.json > /dev/null 2>&1; echo "ops rc=$?"
python scripts/bench_texture.py --batch 4 --steps 5 --out gpurun_out/texture_b4.json > /dev/null 2>&1; echo "texture rc=$?"
python scripts/bench_ops.py --shape 512 --batch 1 --reps 5 --out gpurun_out/ops_512.json > /dev/null 2>&1; echo "ops512 rc=$?"
python -c "import __graft_entry__ as g; g.smoke(); print('smoke ok')" 2>&1 | tail -2
