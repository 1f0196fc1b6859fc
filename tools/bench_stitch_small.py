"""C5 step at the per-GPU size of an 8-GPU run (128 pairs) on one GPU.   python tools/bench_stitch_small.py [pairs]"""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import bench
n = int(sys.argv[1]) if len(sys.argv) > 1 else 128
bench.C5["pairs"] = n
sys.argv = ["bench.py", "--workload", "stitch", "--steps", "10", "--warmup", "3", "--no-cpu"]
bench.main()
