#!/bin/bash
# throughput of every kernel variant on the four reference regimes (run on the GPU box)
for ev in small large; do for m in 0 1; do for v in 1 2 3 4 5 6 7; do
  echo -n "ev=$ev mode=$m variant=$v : "; python tools/profile_solve.py --ev $ev --batch 1048576 --reps 4 --mode $m --variant $v | tail -1
done; done; done
