#!/bin/bash
# One gpurun call while iterating: parity tests, a bench line (no CPU leg), the per-operator tables.
mkdir -p gpurun_out
python -m pytest tests -x -q -m gpu > gpurun_out/pytest_gpu.log 2>&1; echo "pytest rc=$?" >> gpurun_out/pytest_gpu.log
tail -5 gpurun_out/pytest_gpu.log
python bench.py --no-cpu-baseline > gpurun_out/bench_iter.json 2> gpurun_out/bench_iter.err; echo "bench rc=$?"
python scripts/bench_ops.py --shape brats --batch 16 --reps 5 --out gpurun_out/ops_brats_b16.json > /dev/null 2>&1; echo "ops rc=$?"
python scripts/bench_ops.py --shape 512 --batch 1 --reps 5 --out gpurun_out/ops_512.json > /dev/null 2>&1; echo "ops512 rc=$?"
python - <<'P'
import json
b=json.load(open('gpurun_out/bench_iter.json'))
print('value',b['value'],'e2e',b['e2e']['value'])
for k,v in b['kernels'].items(): print(' ',k,v['launches'],v['ms_total'],v['share'])
P
