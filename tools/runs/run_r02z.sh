mkdir -p gpurun_out
(timeout 1500 python -m pytest tests -m gpu -q 2>&1 | tail -5) > gpurun_out/r02z_pytest.log
timeout 900 python bench.py --steps 20 --warmup 5 > gpurun_out/r02z_bench_k20.json 2> gpurun_out/r02z_bench_k20.err
timeout 1500 python bench.py --scaling strong --agents-total 10002432 --scenario GRADUAL,ALL --steps 60 --warmup 5 --no-extras --no-cpu-baseline --max-replays 8 > gpurun_out/r02z_bench_strong_n1.json 2> gpurun_out/r02z_bench_strong_n1.err
cat gpurun_out/r02z_pytest.log
python - <<'PY'
import json
for f in ["r02z_bench_k20","r02z_bench_strong_n1"]:
    try:
        d=json.loads(open("gpurun_out/%s.json"%f).read().strip().split("\n")[-1])
        print(f, "value %.3e (%.4f ms) e2e %.3e (%s ms) h2d %s" % (d["value"], d["ms_per_step"], d["e2e"]["value"], d["e2e"].get("ms_per_step"), d["e2e"]["h2d_bytes_per_step"]), d.get("e2e_lockdown",{}).get("ms_per_step"), d.get("e2e_lockdown_device",{}).get("ms_per_step"))
    except Exception as e: print(f,"ERR",e)
PY
