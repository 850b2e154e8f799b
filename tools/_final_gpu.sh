python -m pytest tests -m gpu -x -q 2>&1 | tail -3
python tools/dump_failures.py 2>&1 | tail -22
bash tools/profile_round.sh r1k > gpurun_out/r1k_profile_round.log 2>&1
tail -3 gpurun_out/r1k_profile_round.log
