mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -q -x -p no:cacheprovider 2>&1 | tail -6
timeout 900 python bench.py --steps 5 --warmup 3 > gpurun_out/run8_bench.json 2> gpurun_out/run8_bench.err
python - <<'PY'
import json
d=json.loads(open('gpurun_out/run8_bench.json').read().strip().splitlines()[-1])
print("ransac", d["value"], d["ms_per_step"], "e2e", d["e2e"]["value"], d["e2e"].get("link_frac"), d["roofline"]["frac"])
print("link", d["e2e"]["link"])
w=d["warp"]; print("warp", w["value"], w["roofline"]["frac"], "e2e", w["e2e"]["value"], w["e2e"].get("copy_gbs"), w["e2e"].get("link_frac"))
print("stitch", d["stitch"]["value"], d["stitch"]["ms_per_step"])
print("aux", d.get("aux"))
print("cpu", d["cpu_baseline"]["value"], d["warp"]["cpu_baseline"]["value"])
PY
tail -3 gpurun_out/run8_bench.err
