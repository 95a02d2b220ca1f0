#!/bin/bash
# Round-end evidence in ONE gpurun call (1 GPU):  tools/profile_round.sh rNN
#   1. plain bench (the number), reference arm, config-5 sweep            -> gpurun_out/<tag>_bench.json, _bench_reference.json, _sweep.json
#   2. ncu launch list of one bench step (time + DRAM bytes per launch)   -> gpurun_out/<tag>_launches.csv
#   3. ncu --set full of the top kernels (NTT passes, pointwise, binv; K1, K3, den) -> gpurun_out/<tag>_full_*.ncu-rep
#      read here with: ncu -i <rep> --page raw --csv > profiles/<tag>_..._full_raw.csv
# A number printed by a run under ncu is never a bench value; steps 2-3 only run if step 1 exited 0.
TAG=${1:-r01}
set -o pipefail
python bench.py > gpurun_out/${TAG}_bench.json 2> gpurun_out/${TAG}_bench.err || exit 1
python bench.py --impl reference --steps 2 --warmup 1 > gpurun_out/${TAG}_bench_reference.json 2>> gpurun_out/${TAG}_bench.err
python tools/sweep.py > gpurun_out/${TAG}_sweep.json 2> gpurun_out/${TAG}_sweep.err
CMD="python bench.py --steps 1 --warmup 1 --no-cpu-baseline --no-e2e"
$CMD > gpurun_out/${TAG}_plain.log 2>&1 &&
ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none --csv \
    --log-file gpurun_out/${TAG}_launches.csv $CMD > gpurun_out/${TAG}_ncu1.log 2>&1
$CMD > gpurun_out/${TAG}_plain2.log 2>&1 &&
# one bench step has 246 launches matching the first regex (112 NTT passes, 19 pointwise, 115 binv_down): skip the warm-up step
ncu --set full --clock-control none --import-source on -k regex:'k_ntt_pass|k_pointwise|k_binv_down' -s 330 -c 12 \
    -o gpurun_out/${TAG}_full_ntt_pw $CMD > gpurun_out/${TAG}_ncu2.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:'k_negbase|k_digit_sums|k_den' -c 4 \
    -o gpurun_out/${TAG}_full_k1_k3 $CMD > gpurun_out/${TAG}_ncu3.log 2>&1
echo profile_round done
