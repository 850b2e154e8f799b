#!/bin/bash
# Round-2 ncu evidence, part 3 (K1 at saturation after the bulk-copy rows): bash tools/profile_round2c.sh r2y
# Each program runs plain first, then under ncu; raw metric pages are exported on the box (reports are too big to merge).
set -u
TAG=${1:-r2y}
OUT=gpurun_out
mkdir -p $OUT
NCU="ncu --set full --clock-control none --import-source on"
for ev in small large; do
  for v in 0 4; do   # 0 = automatic (bulk-copy rows at this size), 4 = per-thread loads, 64 x 4 shape (small EV's former default)
    P="python tools/profile_solve.py --ev $ev --batch 262144 --reps 3 --variant $v"
    $P > $OUT/${TAG}_plain_${ev}_v$v.log 2>&1 &&
    $NCU -k regex:lompc_solve_reg -s 2 -c 1 -o $OUT/${TAG}_k1_${ev}_v$v -f $P > $OUT/${TAG}_ncu_${ev}_v$v.log 2>&1
  done
done
for r in $OUT/${TAG}_*.ncu-rep; do
  ncu -i $r --page raw --csv > ${r%.ncu-rep}.raw.csv 2>/dev/null
  rm -f $r
done
ls -la $OUT | grep ${TAG}_
