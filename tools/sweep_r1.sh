export SWEEP_ALL=1
for s in "wgrad 1 1 16384 320 320 1 1" "wgrad 1 1 16384 320 960 1 1" "wgrad 1 1 4096 640 640 1 1" "wgrad 1 1 4096 640 1920 1 1" "wgrad 1 1 1024 1280 1280 1 1" "wgrad 1 1 16384 320 2560 1 1" "wgrad 1 1 16384 1280 320 1 1" "wgrad 1 16 1024 320 320 3 1" "wgrad 1 16 256 640 640 3 1" "wgrad 16 32 32 320 320 3 3" "wgrad 16 16 16 640 640 3 3" "wgrad 16 8 8 1280 1280 3 3" "wgrad 16 32 32 640 320 3 3" "wgrad 16 16 16 1280 640 3 3"; do
  timeout 200 python tools/gemm_sweep.py $s
done
