for m in 0 1; do for v in 4 6 7; do echo -n "ev=large mode=$m variant=$v : "; python tools/profile_solve.py --ev large --batch 1048576 --reps 4 --mode $m --variant $v | tail -1; done; done
