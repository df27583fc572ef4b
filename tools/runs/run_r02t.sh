mkdir -p gpurun_out
(timeout 900 python -m pytest tests/test_gpu_parity.py tests/test_gpu_lockdown.py tests/test_config1.py -q -x 2>&1 | tail -5) > gpurun_out/r02t_pytest.log
timeout 900 python bench.py --steps 20 --warmup 5 > gpurun_out/r02t_bench_k20.json 2> gpurun_out/r02t_bench_k20.err
timeout 1500 python bench.py --scaling strong --agents-total 10002432 --scenario GRADUAL,ALL --steps 60 --warmup 5 --no-extras > gpurun_out/r02t_bench_strong_n1.json 2> gpurun_out/r02t_bench_strong_n1.err
cat gpurun_out/r02t_pytest.log
python - <<'PY'
import json
for f in ["k20","strong_n1"]:
    try:
        d=json.loads(open("gpurun_out/r02t_bench_%s.json"%f).read().strip().split("\n")[-1])
        print(f, "value %.3e (%.4f ms) e2e %.3e (%.4f ms) lockdown %s / %s cpu %s" % (d["value"], d["ms_per_step"], d["e2e"]["value"], d["e2e"]["ms_per_step"], d.get("e2e_lockdown",{}).get("ms_per_step"), d.get("e2e_lockdown_device",{}).get("ms_per_step"), d.get("cpu_baseline",{}).get("value")))
    except Exception as e: print(f,"ERR",e)
PY
tail -3 gpurun_out/r02t_bench_k20.err gpurun_out/r02t_bench_strong_n1.err
