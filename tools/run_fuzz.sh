mkdir -p gpurun_out
{
timeout 900 python tools/fuzz_warp.py 2500 101 2>&1 | tail -1
timeout 600 python tools/fuzz_score.py 3000 102 2>&1 | tail -1
timeout 600 python tools/fuzz_ransac.py 1500 103 2>&1 | tail -1
timeout 600 python tools/fuzz_compose.py 300 104 2>&1 | tail -1
timeout 600 python tools/fuzz_generic.py 400 105 2>&1 | tail -1
} | tee gpurun_out/r02_fuzz_campaign.txt
