#!/bin/bash
# Multi-GPU evidence in ONE gpurun --gpus 8 call:  tools/scale_round.sh r02
#   sharded parity tests (world 1/2/4/8), weak series N = 2/4/8 (2^20 points per GPU), strong series N = 1/2/4/8 (2^20 points in total),
#   BASELINE config 4 (Vesta, 2^24 points on 8 GPUs) and its parity evidence.  Everything lands in gpurun_out/<tag>_*.
TAG=${1:-r02}
run() { # N, extra args..., output name
  N=$1; shift; OUT=$1; shift
  if [ "$N" = 1 ]; then python bench.py --gpus 1 "$@" > gpurun_out/${TAG}_${OUT}.json 2> gpurun_out/${TAG}_${OUT}.err
  else python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port $((29600 + N)) bench.py --gpus $N "$@" \
       > gpurun_out/${TAG}_${OUT}.json 2> gpurun_out/${TAG}_${OUT}.err; fi
  python -c "import json,sys; d=json.load(open('gpurun_out/${TAG}_${OUT}.json')); print('${OUT}', d['n_gpus'], d['scaling'], round(d['ms_per_step'],1), 'ms', round(d['value']/1e6,2), 'Mpts/s  e2e', round(d['e2e']['ms_per_step'],1) if d.get('e2e') else None)"
}
python -m pytest tests/test_gpu_d_sharded.py -m gpu -q > gpurun_out/${TAG}_gputest_sharded_8gpu.log 2>&1; tail -2 gpurun_out/${TAG}_gputest_sharded_8gpu.log
for N in 1 2 4 8; do run $N strong_${N}gpu --scaling strong --log-n 20 --steps 5 --warmup 3 --no-cpu-baseline; done
for N in 2 4 8; do run $N weak_${N}gpu --steps 5 --warmup 3 --no-cpu-baseline; done
run 8 config4_vesta_2p24_8gpu --curve vesta --log-n 21 --steps 3 --warmup 2 --no-cpu-baseline
python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29513 tools/check_config4.py --log-n 21 \
    > gpurun_out/${TAG}_config4_parity_8gpu.log 2> gpurun_out/${TAG}_config4_parity_8gpu.err
cat gpurun_out/${TAG}_config4_parity_8gpu.log
echo scale_round done
