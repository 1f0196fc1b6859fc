mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -q -x -p no:cacheprovider 2>&1 | tail -3 | tee gpurun_out/r02_pytest_gpu.log
python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -1
timeout 900 python bench.py > gpurun_out/r02_bench.json 2> gpurun_out/r02_bench.err; tail -3 gpurun_out/r02_bench.err
timeout 600 python bench.py --impl reference --steps 2 --warmup 1 > gpurun_out/r02_bench_ref.json 2>gpurun_out/r02_bench_ref.err
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 700 --csv --log-file gpurun_out/r02_launches_bench.csv python bench.py --steps 2 --warmup 3 --no-cpu > gpurun_out/ncu_bench.log 2>&1
timeout 600 ncu --set full --clock-control none --import-source on -k regex:score_fast_kernel -s 1 -c 1 -o gpurun_out/r02_score_fast -f python bench.py --workload ransac --steps 1 --warmup 3 --no-cpu 2>&1 | tail -1
timeout 300 python tools/bench_app.py > gpurun_out/r02_bench_app.json 2>&1; tail -2 gpurun_out/r02_bench_app.json
timeout 600 python tools/fuzz_score.py 1500 11 2>&1 | tail -1
timeout 600 python tools/fuzz_ransac.py 600 12 2>&1 | tail -1
