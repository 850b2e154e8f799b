#!/bin/bash
# ncu --set full of the two K1 kernels (saturating batch); usage: bash tools/_ncu_k1.sh TAG
TAG=${1:-x}; OUT=gpurun_out; mkdir -p $OUT
NCU="ncu --set full --clock-control none --import-source on"
for ev in small large; do
  mode=0; [ $ev = large ] && mode=1
  python tools/profile_solve.py --ev $ev --batch 262144 --reps 3 --mode $mode > $OUT/${TAG}_plain_$ev.log 2>&1 &&
  $NCU -k regex:lompc_solve -s 1 -c 1 -o $OUT/${TAG}_solve_$ev -f python tools/profile_solve.py --ev $ev --batch 262144 --reps 3 --mode $mode > $OUT/${TAG}_ncu_$ev.log 2>&1
  tail -1 $OUT/${TAG}_plain_$ev.log
done
