NVML=1 python tools/loop_strip.py 4096 0
