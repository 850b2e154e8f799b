#!/bin/bash
# One GPU call that collects the round's ncu evidence (run under gpurun from the repo root):
#   bash tools/profile_round.sh r1e
# Every program is run plain first (exit 0 without ncu), then under ncu.  Outputs -> gpurun_out/.
set -u
TAG=${1:-r1}
OUT=gpurun_out
mkdir -p $OUT
NCU="ncu --set full --clock-control none --import-source on"
run() { echo "== $*"; "$@"; echo "rc=$?"; }

# K1 stand-alone, saturating batch: small EV and large EV on the bench's prices (mode 0: optimistic phase),
# large EV on closed-loop-scale sparse prices (mode 1: safeguarded loop)
for cfg in small:0 large:0 large:1; do
  ev=${cfg%:*}; mode=${cfg#*:}; name=$ev; [ $cfg = large:1 ] && name=large_sparse
  run python tools/profile_solve.py --ev $ev --batch 262144 --reps 3 --mode $mode > $OUT/${TAG}_plain_$name.log 2>&1 &&
  $NCU -k regex:lompc_solve -s 1 -c 1 -o $OUT/${TAG}_solve_$name -f python tools/profile_solve.py --ev $ev --batch 262144 --reps 3 --mode $mode > $OUT/${TAG}_ncu_$name.log 2>&1
done
# fused price loop + BiMPC inside the closed loop (64 stations, 4 steps; the captured launches are from step 3)
run python tools/run_fleet.py --stations 64 --steps 4 > $OUT/${TAG}_plain_fleet.log 2>&1 &&
$NCU -k regex:price_station_chain -s 4 -c 1 -o $OUT/${TAG}_price_loop -f python tools/run_fleet.py --stations 64 --steps 4 > $OUT/${TAG}_ncu_price_loop.log 2>&1
$NCU -k regex:bimpc_solve -s 2 -c 1 -o $OUT/${TAG}_bimpc -f python tools/run_fleet.py --stations 256 --steps 3 > $OUT/${TAG}_ncu_bimpc.log 2>&1
# launch list of the bench command
run python bench.py --steps 2 --warmup 1 --closed-loop-stations 64 --closed-loop-steps 3 --no-cpu-baseline > $OUT/${TAG}_plain_bench.log 2>&1 &&
ncu --metrics gpu__time_duration.sum --clock-control none -c 2000 --csv --log-file $OUT/${TAG}_launches_bench.csv python bench.py --steps 2 --warmup 1 --closed-loop-stations 64 --closed-loop-steps 3 --no-cpu-baseline > $OUT/${TAG}_ncu_bench.log 2>&1
# gpurun merges at most 64 MiB back: keep the raw metric pages of every capture (what tools/ncu_summary.py and
# tools/ncu_stalls.py read) and only the two K1 reports themselves
for r in $OUT/${TAG}_*.ncu-rep; do ncu -i $r --page raw --csv > ${r%.ncu-rep}.raw.csv 2>/dev/null; done
rm -f $OUT/${TAG}_solve_large_sparse.ncu-rep $OUT/${TAG}_price_loop.ncu-rep $OUT/${TAG}_bimpc.ncu-rep
ls -la $OUT | tail -20
