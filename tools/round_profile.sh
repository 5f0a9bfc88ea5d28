#!/bin/bash
# Round-end evidence on one B200 (run under gpurun from the repo root); outputs land in gpurun_out/.  Every step is bounded.
set -x
R=${1:-r2}
# one eager cfg-2 step (fwd + bwd + clip + AdamW): duration + DRAM traffic of every launch (roofline.traffic, kernel shares)
timeout 450 ncu --profile-from-start off --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none \
    --csv --log-file gpurun_out/${R}_launches_step.csv python tools/step_once.py > gpurun_out/${R}_step_once.log 2>&1
# launch list of the bench command itself: the kernels of its timed region (one graph-replayed step between cudaProfilerStart/Stop).
# ncu spends ~150 ms per graph kernel node: 3,006 launches need ~470 s (the r2 capture was cut at 450 s after 2,898 of them)
timeout 560 ncu --profile-from-start off --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/${R}_launches_bench.csv \
    python bench.py --steps 1 --warmup 3 --no-cpu-baseline --no-roofline --profile-timed-region > gpurun_out/${R}_bench_under_ncu.log 2>&1
# full captures of the dominant kernel: the small-K feed-forward projection and a 1x1 weight gradient with its fused bias gradient
timeout 150 ncu --set full --clock-control none --import-source on -k regex:gemm_tc -s 3 -c 1 -f -o gpurun_out/${R}_full_gemm_ff_proj_final \
    python tools/gemm_one.py fwd 1 1 16384 320 2560 1 1 > /dev/null 2>&1
timeout 150 ncu --set full --clock-control none --import-source on -k regex:gemm_tc -s 3 -c 1 -f -o gpurun_out/${R}_full_gemm_wgrad320_dbias \
    python tools/gemm_one.py wgrad 1 1 16384 320 320 1 1 5 dbias > /dev/null 2>&1
ls -la gpurun_out | tail -8
