mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -q -x -p no:cacheprovider 2>&1 | tail -4 | tee gpurun_out/r02_pytest_gpu.log
python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -2
timeout 900 python bench.py --steps 3 --warmup 3 --no-cpu > gpurun_out/r02_bench_quick.json 2> gpurun_out/r02_bench_quick.err; tail -3 gpurun_out/r02_bench_quick.err
python - <<'PY'
import json
d=json.loads(open('gpurun_out/r02_bench_quick.json').read().strip().splitlines()[-1])
print("aux", json.dumps(d.get("aux"), indent=1)[:3000])
PY
