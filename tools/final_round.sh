#!/bin/bash
# One GPU call at the end of a round: the GPU test suite, smoke(), the default bench line, then the ncu evidence of the
# shipped kernels (summaries are copied into profiles/ offline).   bash tools/final_round.sh r02
R=${1:-r02}
O=gpurun_out
mkdir -p $O
( time timeout 420 python -m pytest tests -m gpu -x -q ) > $O/${R}_final_pytest.log 2>&1; tail -3 $O/${R}_final_pytest.log
timeout 90 python __graft_entry__.py --smoke > $O/${R}_final_smoke.log 2>&1; tail -1 $O/${R}_final_smoke.log
( time timeout 420 python bench.py ) > $O/${R}_final_bench.log 2>&1; grep -c '^{"metric"' $O/${R}_final_bench.log
cap() {  # name, kernel regex, launches to skip, command...
  local name=$1 rex=$2 skip=$3; shift 3
  timeout 150 ncu --set full --clock-control none --import-source on -k "regex:$rex" --launch-skip $skip -c 1 -f -o $O/prof_${R}_$name "$@" > $O/${R}_ncu_$name.log 2>&1
  python tools/ncu_summary.py $O/prof_${R}_$name.ncu-rep > $O/${R}_ncu_$name.txt 2>/dev/null; head -3 $O/${R}_ncu_$name.txt
}
cap stream_const rhs_stream_kernel 3 python tools/run_variant.py const_recompute 6
timeout 200 ncu --metrics gpu__time_duration.sum --clock-control none -c 700 --csv --log-file $O/${R}_launches_bench.csv \
    python bench.py --steps 2 --warmup 3 --quick --no-strong --no-cpu-baseline > $O/${R}_launches_bench.log 2>&1
python tools/launch_summary.py $O/${R}_launches_bench.csv > $O/${R}_launches_bench_summary.txt 2>/dev/null; head -12 $O/${R}_launches_bench_summary.txt
cap stream_general rhs_stream_kernel 3 python tools/run_variant.py general_recompute 6
