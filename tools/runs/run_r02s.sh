mkdir -p gpurun_out
(timeout 1500 python -m pytest tests/test_gpu_parity.py tests/test_gpu_lockdown.py tests/test_gpu_scale.py tests/test_gpu_checked.py -q -x 2>&1 | tail -8) > gpurun_out/r02s_pytest.log
timeout 900 python bench.py --steps 60 --warmup 5 --scenario GRADUAL,ALL --no-extras --no-cpu-baseline > gpurun_out/r02s_bench_lockdown.json 2> gpurun_out/r02s_bench_lockdown.err
timeout 900 python bench.py --steps 20 --warmup 5 --no-cpu-baseline > gpurun_out/r02s_bench_k20.json 2> gpurun_out/r02s_bench_k20.err
cat gpurun_out/r02s_pytest.log
python - <<'PY'
import json
for f in ["lockdown","k20"]:
    d=json.loads(open("gpurun_out/r02s_bench_%s.json"%f).read().strip().split("\n")[-1])
    print(f, "value %.3e (%.4f ms) e2e %.3e (%.4f ms) h2d %s lockdown %s / %s" % (d["value"], d["ms_per_step"], d["e2e"]["value"], d["e2e"]["ms_per_step"], d["e2e"]["h2d_bytes_per_step"], d.get("e2e_lockdown",{}).get("ms_per_step"), d.get("e2e_lockdown_device",{}).get("ms_per_step")))
PY
tail -3 gpurun_out/r02s_bench_lockdown.err
