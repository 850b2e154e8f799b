#!/bin/bash
# Round-2 ncu evidence, one GPU call (run under gpurun from the repo root):   bash tools/profile_round2.sh r2k
# Every command is run plain first (exit 0 without ncu), then under ncu.  Outputs -> gpurun_out/.
set -u
TAG=${1:-r2}
OUT=gpurun_out
mkdir -p $OUT
NCU="ncu --set full --clock-control none --import-source on"
BENCH="python bench.py --steps 3 --warmup 3 --closed-loop-stations 0 --no-saturated --no-cpu-baseline --sharded-iters 0"
# the headline kernel: lompc_solve_warp_kernel, 512 small + 512 large QPs in ONE launch (the bench's solve set)
$BENCH > $OUT/${TAG}_plain_bench_set.log 2>&1 &&
$NCU -k regex:lompc_solve_warp -s 10 -c 2 -o $OUT/${TAG}_solve_warp_set -f $BENCH > $OUT/${TAG}_ncu_solve_warp_set.log 2>&1
# the same kernel at a medium batch (16,384 QPs, one EV type): tools/time_k1.py launches it via the variant switch
python tools/time_k1.py --batches 16384 --variants 8 --reps 3 > $OUT/${TAG}_plain_time_k1.log 2>&1 &&
$NCU -k regex:lompc_solve_warp -s 4 -c 1 -o $OUT/${TAG}_solve_warp_16k -f python tools/time_k1.py --batches 16384 --variants 8 --reps 3 > $OUT/${TAG}_ncu_solve_warp_16k.log 2>&1
# launch list of the default bench command (short closed loop, short sharded loop)
LL="python bench.py --steps 2 --warmup 3 --closed-loop-stations 64 --closed-loop-steps 3 --no-cpu-baseline --sharded-iters 12"
$LL > $OUT/${TAG}_plain_bench.log 2>&1 &&
ncu --metrics gpu__time_duration.sum --clock-control none -c 3000 --csv --log-file $OUT/${TAG}_launches_bench.csv $LL > $OUT/${TAG}_ncu_bench.log 2>&1
for r in $OUT/${TAG}_*.ncu-rep; do ncu -i $r --page raw --csv > ${r%.ncu-rep}.raw.csv 2>/dev/null; done
ls -la $OUT | grep ${TAG}_
