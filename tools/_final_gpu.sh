python bench.py > gpurun_out/r1k_bench_n1.json 2> gpurun_out/r1k_bench_n1.err; tail -c 600 gpurun_out/r1k_bench_n1.err
python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/r1k_bench_ref.json 2>/dev/null; cat gpurun_out/r1k_bench_ref.json | cut -c1-400
python tools/run_fleet.py --stations 4096 --steps 96 2>/dev/null | tail -1 > gpurun_out/r1k_fleet_4096x96.json; cut -c1-400 gpurun_out/r1k_fleet_4096x96.json
python tools/stress_fleet.py 2>&1 | tail -8 > gpurun_out/r1k_stress_fleet.txt; cat gpurun_out/r1k_stress_fleet.txt
