#!/bin/bash
# Same-box A/B of two builds of libonerf_sm100.so inside ONE gpurun call (box-to-box variance is larger than most kernel
# deltas).  Usage, from the repo root in the build container:
#   cp object_nerf_b200/libonerf_sm100.so ab/libonerf_old.so      # the baseline build
#   (edit kernels, make -C object_nerf_b200/csrc)                 # the candidate build stays in place
#   gpurun --timeout 900 -- 'bash tools/ab_bench.sh'
# Prints: bf16 parity tests of the candidate, then rays/s, fused-kernel TFLOP/s, fraction of peak and ms per fine-pass
# launch for new / old / new (the library file is swapped between runs and restored at the end).
set -u
run() {
  timeout 300 python bench.py --steps 3 --no-cpu 2>&1 | tail -1 | python -c "import sys,json; d=json.loads(sys.stdin.read()); print('$1', d['value'], d['roofline']['achieved'], d['roofline']['frac'], d['roofline']['ms_per_launch'])"
}
LIB=object_nerf_b200/libonerf_sm100.so
cp $LIB /tmp/new.so
timeout 300 python -m pytest tests -m gpu -q -x -k "bf16" 2>&1 | tail -2
run new
if [ -f ab/libonerf_old.so ]; then
  cp ab/libonerf_old.so $LIB
  run old
  cp /tmp/new.so $LIB
  run new
fi
