python -m pytest tests -m gpu -x -q 2>&1 | tail -3
for ev in small large; do for m in 0 1; do echo -n "ev=$ev mode=$m : "; python tools/profile_solve.py --ev $ev --batch 1048576 --reps 4 --mode $m | tail -1; done; done
bash tools/_ncu_k1.sh r1j
