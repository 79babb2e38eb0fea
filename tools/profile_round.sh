#!/bin/bash
# One GPU call that refreshes the ncu evidence under gpurun_out/ (summaries are then written into profiles/ offline):
#   bash tools/profile_round.sh r02
# --set full captures of the shipped kernels (one launch each, ~40 replay passes) and the launch list of a short bench.
R=${1:-r02}
O=gpurun_out
cap() {  # name, kernel regex, launches to skip, command...
  local name=$1 rex=$2 skip=$3; shift 3
  ncu --set full --clock-control none --import-source on -k "regex:$rex" --launch-skip $skip -c 1 -f -o $O/prof_${R}_$name "$@" > $O/${R}_ncu_$name.log 2>&1
  tail -1 $O/${R}_ncu_$name.log
}
cap stream_const rhs_stream_kernel 3 python tools/run_variant.py const_recompute 6
cap stream_general rhs_stream_kernel 3 python tools/run_variant.py general_recompute 6
cap stream_system rhs_stream_kernel 10 python tools/sys_rhs.py 4096
cap spmv spmv_tile_kernel 3 python bench.py --steps 3 --warmup 1 --quick --no-extras --no-strong --no-cpu-baseline --variant const_recompute
ncu --metrics gpu__time_duration.sum --clock-control none -c 700 --csv --log-file $O/${R}_launches_bench.csv \
    python bench.py --steps 2 --warmup 3 --quick --no-strong --no-cpu-baseline > $O/${R}_launches_bench.log 2>&1
python tools/launch_summary.py $O/${R}_launches_bench.csv | head -30
