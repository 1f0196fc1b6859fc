mkdir -p gpurun_out
timeout 1200 python -m pytest tests -m gpu -q -x -p no:cacheprovider 2>&1 | tail -6
python tools/prof_ransac.py 1000000 2>&1 | tail -2
python tools/prof_ransac.py 125000 2>&1 | tail -2
