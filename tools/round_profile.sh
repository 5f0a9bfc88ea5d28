#!/bin/bash
# Round-end evidence on one B200 (run under gpurun from the repo root); outputs land in gpurun_out/.
set -x
R=${1:-r1}
timeout 600 python -m pytest tests -m gpu -q 2>&1 | tail -4 > gpurun_out/${R}_tests_gpu.txt
timeout 400 python bench.py --steps 20 --warmup 3 > gpurun_out/${R}_bench_n1.json 2> gpurun_out/${R}_bench_n1.err
# launch list of the bench command (serialised, cold-cache per-launch times: shares matter, not absolutes)
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 60000 --csv --log-file gpurun_out/${R}_launches_bench.csv \
    python bench.py --steps 2 --warmup 1 --no-cpu-baseline > gpurun_out/${R}_bench_under_ncu.log 2>&1
# one eager step: time + DRAM traffic of every launch (for roofline.traffic of the GEMM kernel)
timeout 900 ncu --profile-from-start off --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none \
    --csv --log-file gpurun_out/${R}_launches_step.csv python tools/step_once.py > gpurun_out/${R}_step_once.log 2>&1
# full captures of the dominant kernel on its two heaviest shape families
timeout 300 ncu --set full --clock-control none --import-source on -k regex:gemm_tc -s 3 -c 1 -f -o gpurun_out/${R}_full_gemm_conv320 \
    python tools/gemm_one.py fwd 16 32 32 320 320 3 3 > /dev/null 2>&1
timeout 300 ncu --set full --clock-control none --import-source on -k regex:gemm_tc -s 3 -c 1 -f -o gpurun_out/${R}_full_gemm_wgrad320 \
    python tools/gemm_one.py wgrad 1 1 16384 320 320 1 1 > /dev/null 2>&1
ls -la gpurun_out | tail -12
