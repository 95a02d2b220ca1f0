#!/bin/bash
# Build a development variant of the library for A/B timing inside one gpurun call (Pallas engine only, own object files):
#   tools/variant.sh NAME [extra nvcc flags, e.g. -DEAGEN_SOME_SWITCH=1]   ->  gpurun_variants/libeagen_NAME.so
# then on the box:  tools/abn.sh gpurun_variants/libeagen_A.so gpurun_variants/libeagen_B.so
NAME=$1; shift
mkdir -p gpurun_variants
EAGEN_OUT=$PWD/gpurun_variants/libeagen_$NAME.so EAGEN_OBJ_SUFFIX=.$NAME EAGEN_NVCC_EXTRA="-DEAGEN_DEV_PALLAS_ONLY $*" python halo2-liam-eagen-msm_b200/build.py --force
