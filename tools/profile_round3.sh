#!/bin/bash
# Evidence of the last session of round 2 (profiles/r3*), one GPU (run under gpurun from the repo root):
#     bash tools/profile_round3.sh r3
# Every command is run plain first (exit 0 without ncu), then under ncu.  Outputs -> gpurun_out/.
set -u
TAG=${1:-r3}
OUT=gpurun_out
mkdir -p $OUT
# final bench line + reference arm
python bench.py > $OUT/${TAG}j_bench_n1.json 2> $OUT/${TAG}j_bench_n1.err
python bench.py --impl reference > $OUT/${TAG}f_bench_ref.json 2> $OUT/${TAG}f_bench_ref.err
# A/B switches of the phase-split loop (sharded leg only)
B="python bench.py --steps 5 --warmup 3 --closed-loop-stations 0 --no-saturated --no-cpu-baseline"
LOMPC_SHARD_OVERLAP=0 $B > $OUT/${TAG}c_ov0.json 2> /dev/null
LOMPC_SHARD_FUSED_BK=0 $B > $OUT/${TAG}h_bk0.json 2> /dev/null
$B > $OUT/${TAG}h_bk1.json 2> /dev/null
# launch list of the default bench command (short closed loop, short sharded loop)
LL="python bench.py --steps 5 --warmup 3 --closed-loop-stations 256 --closed-loop-steps 3 --no-cpu-baseline --sharded-iters 12"
$LL > $OUT/${TAG}j_plain_bench.log 2>&1 &&
ncu --metrics gpu__time_duration.sum --clock-control none -c 2000 --csv --log-file $OUT/${TAG}j_launches_bench.csv $LL > $OUT/${TAG}j_ncu_bench.log 2>&1
# launch list of the sharded leg alone with warm caches (per-kernel times of one iteration)
SL="python bench.py --steps 2 --warmup 3 --closed-loop-stations 0 --no-saturated --no-cpu-baseline --sharded-iters 40"
ncu --metrics gpu__time_duration.sum --clock-control none --cache-control none -c 1500 --csv --log-file $OUT/${TAG}g_launches_sharded.csv $SL > $OUT/${TAG}g_ncu.log 2>&1
# ncu --set full of the loop's three kernels (large-EV half)
ncu --set full --clock-control none --import-source on -k regex:"lompc_solve_reg_kernel|group_step_kernel|sc_solve_bookkeep" -s 150 -c 6 -o $OUT/${TAG}k_sharded -f $SL > $OUT/${TAG}k_ncu.log 2>&1
ncu -i $OUT/${TAG}k_sharded.ncu-rep --page raw --csv > $OUT/${TAG}k_sharded.raw.csv 2>/dev/null
# 2 and 8 GPUs (gpurun --gpus N): python -m torch.distributed.run --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1 \
#     --master-port 29541 bench.py --gpus N --steps 50 --warmup 5 --closed-loop-stations 0 --no-saturated --no-cpu-baseline
ls -la $OUT | grep ${TAG}
