mkdir -p gpurun_out
(timeout 1500 python -m pytest tests -m gpu -q -x 2>&1 | tail -8) > gpurun_out/r02r_pytest.log
python tools/quick_time.py --communities 89 --days 196 --warm 28 > gpurun_out/r02r_t1m.json 2>/dev/null
python tools/quick_time.py --communities 89 --replicas 16 --days 196 --warm 28 --flags 0 > gpurun_out/r02r_t1m_r16.json 2>/dev/null
python tools/quick_time.py --communities 89 --replicas 4 --days 196 --warm 28 --flags 0 > gpurun_out/r02r_t1m_r4.json 2>/dev/null
python tools/quick_time.py --communities 890 --days 98 --warm 28 > gpurun_out/r02r_t10m.json 2>/dev/null
timeout 900 python bench.py --steps 20 --warmup 5 > gpurun_out/r02r_bench_k20.json 2> gpurun_out/r02r_bench_k20.err
cat gpurun_out/r02r_pytest.log; cut -c1-150 gpurun_out/r02r_t1m.json gpurun_out/r02r_t1m_r16.json gpurun_out/r02r_t1m_r4.json gpurun_out/r02r_t10m.json
python - <<'PY'
import json
d=json.loads(open("gpurun_out/r02r_bench_k20.json").read().strip().split("\n")[-1])
print("bench k20 value %.3e (%.4f ms) e2e %.3e (%.4f ms) lockdown %s / %s rep16 %.3e" % (d["value"], d["ms_per_step"], d["e2e"]["value"], d["e2e"]["ms_per_step"], d.get("e2e_lockdown",{}).get("ms_per_step"), d.get("e2e_lockdown_device",{}).get("ms_per_step"), d.get("batched_replicas",{}).get("value",0)))
PY
tail -3 gpurun_out/r02r_bench_k20.err
