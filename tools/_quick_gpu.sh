for ev in small large; do for m in 0 1; do echo -n "ev=$ev mode=$m : "; python tools/profile_solve.py --ev $ev --batch 1048576 --reps 4 --mode $m | tail -1; done; done
python tools/profile_solve.py --ev small --batch 512 --reps 4 | tail -1
python tools/profile_solve.py --ev large --batch 512 --reps 4 | tail -1
