bash tools/profile_round.sh r1l > gpurun_out/r1l_profile_round.log 2>&1
python bench.py > gpurun_out/r1l_bench_n1.json 2> gpurun_out/r1l_bench_n1.err; tail -c 300 gpurun_out/r1l_bench_n1.err
python tools/run_fleet.py --stations 4096 --steps 96 2>/dev/null | tail -1 > gpurun_out/r1l_fleet_4096x96.json; cut -c1-300 gpurun_out/r1l_fleet_4096x96.json
python tools/dump_failures.py 2>&1 | grep -c "failures 0"
python tools/stress_fleet.py 2>&1 | tail -2
