#!/bin/bash
# ncu evidence of the current build, run on the GPU box (one GPU):  bash tools/capture_ncu.sh <tag>
#   1. --set full captures of the attention kernels at S1 / S2 (tools/kernel_only.py) and of the epilogue kernels at S1 / S2
#      (tools/epilogue_only.py), exported to raw CSV / details text / source CSV (the .ncu-rep files stay on the box);
#   2. the launch list (gpu__time_duration.sum) of a short bench.py run.
TAG=${1:-r02f}
OUT=gpurun_out
mkdir -p $OUT
NCU="ncu --set full --clock-control none --import-source on"
cap() {   # name, kernel regex, skip, count, command...
  local name=$1 rx=$2 skip=$3 cnt=$4; shift 4
  timeout 600 $NCU -k regex:"$rx" -s $skip -c $cnt -f -o /tmp/$name "$@" > $OUT/${name}_ncu.log 2>&1
  ncu -i /tmp/$name.ncu-rep --page raw --csv > $OUT/${name}_raw.csv 2>/dev/null
  ncu -i /tmp/$name.ncu-rep --page details > $OUT/${name}_details.txt 2>/dev/null
  ncu -i /tmp/$name.ncu-rep --page source --csv > $OUT/${name}_source.csv 2>/dev/null
}
cap ${TAG}_ncu_full_S1 "vil_tc_" 5 5 python tools/kernel_only.py S1 2
cap ${TAG}_ncu_full_S2 "vil_tc_" 5 5 python tools/kernel_only.py S2 2
cap ${TAG}_ncu_full_epilogue_S1 "addnorm|bias_act|colsum_reduce" 8 8 python tools/epilogue_only.py S1 2
cap ${TAG}_ncu_full_epilogue_S2 "addnorm|bias_act|colsum_reduce" 8 8 python tools/epilogue_only.py S2 2
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file $OUT/${TAG}_launches.csv \
  python bench.py --steps 2 --warmup 3 --no-cpu-baseline --no-microbench > $OUT/${TAG}_launches_bench.log 2>&1
ls -la $OUT | tail -30
