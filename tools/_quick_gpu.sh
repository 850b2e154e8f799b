python -m pytest tests -m gpu -x -q 2>&1 | tail -3
for ev in small large; do for m in 0 1; do echo -n "ev=$ev mode=$m : "; python tools/profile_solve.py --ev $ev --batch 1048576 --reps 4 --mode $m | tail -1; done; done
python tools/time_single_group.py
python tools/run_fleet.py --stations 1024 --steps 24 2>&1 | tail -1 | cut -c1-330
python bench.py --steps 50 --warmup 5 --no-cpu-baseline --closed-loop-stations 64 --closed-loop-steps 3 2>/dev/null | python -c "
import json,sys
d=json.loads(sys.stdin.readlines()[-1]); print('value',d['value'],'ms',d['ms_per_step'],'e2e',d['e2e']['value'],'sat',{k:v['qp_per_s']/1e6 for k,v in d['saturated']['per_type'].items()})"
