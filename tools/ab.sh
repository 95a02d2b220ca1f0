#!/bin/bash
# A/B timing of two prebuilt libraries in one gpurun call (boxes differ by >10 %, so never compare across calls)
# usage: tools/ab.sh lib_old.so lib_new.so [bench args]
A=$1; B=$2; shift 2
for v in $A $B $A $B; do
  cp $v halo2-liam-eagen-msm_b200/libeagen_msm.so
  python bench.py --steps 6 --warmup 2 --no-cpu-baseline --no-e2e "$@" | python -c "
import json,sys; d=json.loads(sys.stdin.read()); print('$v', round(d['ms_per_step'],1), {k:round(v,3) for k,v in list(d['kernel_shares'].items())[:5]}, d['clocks'])"
done
