#!/bin/bash
# Round-2 ncu evidence, part 2 (the closed loop after the parametric price loop): bash tools/profile_round2b.sh r2w
set -u
TAG=${1:-r2w}
OUT=gpurun_out
mkdir -p $OUT
NCU="ncu --set full --clock-control none --import-source on"
# the parametric chain kernel inside a fleet-scale step (4,096 stations; the launches of step 2: small EV, large EV)
FLEET="python tools/run_fleet.py --stations 4096 --steps 3"
$FLEET > $OUT/${TAG}_plain_fleet.log 2>&1 &&
$NCU -k regex:price_station_chain_warp -s 4 -c 2 -o $OUT/${TAG}_chain_warp -f $FLEET > $OUT/${TAG}_ncu_chain_warp.log 2>&1
# launch list of the default bench command (short legs)
LL="python bench.py --steps 5 --warmup 3 --closed-loop-stations 256 --closed-loop-steps 3 --no-cpu-baseline --sharded-iters 12"
$LL > $OUT/${TAG}_plain_bench.log 2>&1 &&
ncu --metrics gpu__time_duration.sum --clock-control none -c 4000 --csv --log-file $OUT/${TAG}_launches_bench.csv $LL > $OUT/${TAG}_ncu_bench.log 2>&1
for r in $OUT/${TAG}_*.ncu-rep; do ncu -i $r --page raw --csv > ${r%.ncu-rep}.raw.csv 2>/dev/null; done
rm -f $OUT/${TAG}_chain_warp.ncu-rep
ls -la $OUT | grep ${TAG}_
