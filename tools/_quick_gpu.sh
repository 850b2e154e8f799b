python -m pytest tests -m gpu -x -q 2>&1 | tail -3
python tools/run_horizon_sweep.py 2>&1 | tail -4 | tee gpurun_out/r1k_horizon_sweep.jsonl
python tools/profile_solve.py --ev small --batch 262144 --reps 3 --variant 1 | tail -1
python tools/profile_solve.py --ev large --batch 262144 --reps 3 --variant 1 | tail -1
python tools/profile_solve.py --ev small --batch 131072 --N 48 --reps 3 | tail -1
python tools/profile_solve.py --ev large --batch 131072 --N 48 --reps 3 | tail -1
python tools/profile_solve.py --ev small --batch 65536 --N 96 --reps 3 | tail -1
python tools/profile_solve.py --ev large --batch 65536 --N 96 --reps 3 | tail -1
