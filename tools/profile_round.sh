#!/bin/bash
# Round-end evidence in ONE gpurun call (1 GPU):  tools/profile_round.sh rNN
#   1. plain bench (the number), reference arm, config-5 sweep            -> gpurun_out/<tag>_bench.json, _bench_reference.json, _sweep.json
#   2. ncu launch list of one witness step (time + DRAM bytes per launch) -> gpurun_out/<tag>_launches.csv
#   3. ncu --set full of every kernel >= 1 % of the step (NTT passes both directions, pointwise, binv up/down/base, fixup,
#      merge_desc, pair_finish, den, digit sums, K1)                      -> gpurun_out/<tag>_full_*_raw.csv (the .ncu-rep stay in /tmp: gpurun_out is capped at 64 MiB)
# A number printed by a run under ncu is never a bench value; steps 2-3 only run if the same command exited 0 without ncu.
TAG=${1:-r02}
set -o pipefail
python bench.py > gpurun_out/${TAG}_bench.json 2> gpurun_out/${TAG}_bench.err || exit 1
python bench.py --impl reference --steps 2 --warmup 1 > gpurun_out/${TAG}_bench_reference.json 2>> gpurun_out/${TAG}_bench.err
python tools/sweep.py > gpurun_out/${TAG}_sweep.json 2> gpurun_out/${TAG}_sweep.err
CMD="python tools/one_step.py"
$CMD > gpurun_out/${TAG}_plain.log 2>&1 &&
ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none --csv \
    --log-file gpurun_out/${TAG}_launches.csv $CMD > gpurun_out/${TAG}_ncu1.log 2>&1
$CMD > gpurun_out/${TAG}_plain2.log 2>&1 &&
# one witness step has ~60 NTT launches, 19 pointwise, ~110 binv_down: skip the warm-up step and the small bottom levels
ncu --set full --clock-control none --import-source on -k regex:'k_ntt_pass|k_pointwise|k_binv_down|k_binv_up' -s 290 -c 14 \
    -o /tmp/${TAG}_full_ntt_pw $CMD > gpurun_out/${TAG}_ncu2.log 2>&1
# the small kernels, each family at the launches of the SECOND step (the first step of one_step.py is the warm-up)
ncu --set full --clock-control none --import-source on -k regex:'k_negbase|k_digit_sums|k_scatter_points|k_multiples_proj' -s 4 -c 4 \
    -o /tmp/${TAG}_full_small $CMD > gpurun_out/${TAG}_ncu3.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:'k_den|k_fixup' -s 40 -c 6 \
    -o /tmp/${TAG}_full_den_fixup $CMD > gpurun_out/${TAG}_ncu4.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:'k_merge_desc|k_pair_finish|k_pair_den|k_leaf_lines|k_binv_base' -s 78 -c 6 \
    -o /tmp/${TAG}_full_pyramid $CMD > gpurun_out/${TAG}_ncu5.log 2>&1
for r in ntt_pw small den_fixup pyramid; do
  ncu -i /tmp/${TAG}_full_${r}.ncu-rep --page raw --csv > gpurun_out/${TAG}_full_${r}_raw.csv 2>/dev/null
done
echo profile_round done
