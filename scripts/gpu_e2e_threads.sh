#!/bin/bash
# (1) e2e pipeline experiment with the current library: host threads x ordered / independent uploads
# (2) the next build of the library (VOXCUDA_LIB): the whole GPU suite, then the bench
mkdir -p gpurun_out
for t in 2 3 4; do
  for mode in "" "--independent-uploads"; do
    python bench.py --steps 20 --warmup 3 --no-cpu-baseline --threads $t $mode 2>gpurun_out/e2e_err.log | tail -1 > gpurun_out/e2e_t${t}${mode:+_indep}.json
  done
done
export VOXCUDA_LIB=$PWD/voxlogica_b200/libvoxcuda_next.so
timeout 900 python -m pytest tests -q -m gpu -x 2>&1 | tail -15 > gpurun_out/next_tests.log
for t in 2 3; do
  python bench.py --steps 20 --warmup 3 --no-cpu-baseline --threads $t 2>gpurun_out/next_err.log | tail -1 > gpurun_out/next_t${t}.json
done
python - <<'PY'
import json,glob
for f in sorted(glob.glob('gpurun_out/e2e_t*.json'))+sorted(glob.glob('gpurun_out/next_t*.json')):
    try:
        d=json.loads(open(f).read())
        print(f, round(d['value'],1), round(d['e2e']['value'],1), round(d['e2e']['ms_per_step'],3), d['kernels_serial']['ms_per_step'])
    except Exception as e:
        print(f, 'ERR', e)
PY
cat gpurun_out/next_tests.log
