mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_warp.py tests/test_gpu_full_size.py -m gpu -q -x -p no:cacheprovider 2>&1 | tail -4
timeout 600 python tools/fuzz_warp.py 400 21 2>&1 | tail -4
python tools/check_strip.py 256 2>&1 | grep "variant 0"
python bench.py --workload stitch --steps 5 --warmup 2 --no-cpu 2>/dev/null | python -c "
import json,sys
d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print('stitch', d['value'], d['ms_per_step'], d['ransac_hyp_pts_per_s'])"
