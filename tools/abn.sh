#!/bin/bash
# A/B/C... timing of several prebuilt libraries in one gpurun call, two interleaved rounds (boxes differ by >10 %, never compare across calls)
# usage: tools/abn.sh lib1.so lib2.so ...   (the last library listed stays installed)
for round in 1 2; do
for v in "$@"; do
  cp $v halo2-liam-eagen-msm_b200/libeagen_msm.so
  python bench.py --steps 6 --warmup 2 --no-cpu-baseline --no-e2e | python -c "
import json,sys; d=json.loads(sys.stdin.read()); print('$v', round(d['ms_per_step'],1), {k:round(v,3) for k,v in list(d['kernel_shares'].items())[:5]})"
done
done
