#!/bin/bash
# Round-end evidence on one B200 (run under gpurun from the repo root); outputs land in gpurun_out/.  Every step is bounded.
set -x
R=${1:-r2}
# one eager cfg-2 step (fwd + bwd + clip + AdamW): duration + DRAM traffic of every launch (roofline.traffic, kernel shares)
timeout 600 ncu --profile-from-start off --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none \
    --csv --log-file gpurun_out/${R}_launches_step.csv python tools/step_once.py > gpurun_out/${R}_step_once.log 2>&1
# launch list of the bench command itself: two graph-replayed steps of the timed region
timeout 700 ncu --metrics gpu__time_duration.sum --clock-control none -s ${SKIP:-30000} -c 7200 --csv --log-file gpurun_out/${R}_launches_bench.csv \
    python bench.py --steps 2 --warmup 3 --no-cpu-baseline --no-roofline > gpurun_out/${R}_bench_under_ncu.log 2>&1
# full captures of the dominant kernel: 3x3 conv with epilogue statistics, the small-K feed-forward projection, a 1x1 wgrad
timeout 200 ncu --set full --clock-control none --import-source on -k regex:gemm_tc -s 3 -c 1 -f -o gpurun_out/${R}_full_gemm_conv320_stats \
    python tools/gemm_one.py fwd 16 32 32 320 320 3 3 5 stats > /dev/null 2>&1
timeout 200 ncu --set full --clock-control none --import-source on -k regex:gemm_tc -s 3 -c 1 -f -o gpurun_out/${R}_full_gemm_ff_proj \
    python tools/gemm_one.py fwd 1 1 16384 320 2560 1 1 > /dev/null 2>&1
timeout 200 ncu --set full --clock-control none --import-source on -k regex:gemm_tc -s 3 -c 1 -f -o gpurun_out/${R}_full_gemm_wgrad320 \
    python tools/gemm_one.py wgrad 1 1 16384 320 320 1 1 > /dev/null 2>&1
# GroupNorm apply (forward, statistics from the producer) and the optimizer update
timeout 200 ncu --set full --clock-control none --import-source on -k regex:gn_stream -s 4 -c 3 -f -o gpurun_out/${R}_full_groupnorm \
    python tools/gn_profile.py > /dev/null 2>&1
timeout 300 python tools/shape_bench.py --top 150 > gpurun_out/${R}_shape_bench_final.txt 2>&1
ls -la gpurun_out | tail -14
