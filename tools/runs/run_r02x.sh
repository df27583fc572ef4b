mkdir -p gpurun_out
(timeout 1500 python -m pytest tests -m gpu -q 2>&1 | tail -5) > gpurun_out/r02x_pytest.log
timeout 900 python bench.py --steps 20 --warmup 5 > gpurun_out/r02x_bench_k20.json 2> gpurun_out/r02x_bench_k20.err
timeout 900 python bench.py --steps 20 --warmup 5 --impl reference > gpurun_out/r02x_bench_k20_reference.json 2> gpurun_out/r02x_bench_k20_reference.err
timeout 1200 python bench.py --no-extras > gpurun_out/r02x_bench_default.json 2> gpurun_out/r02x_bench_default.err
timeout 1500 python bench.py --scaling strong --agents-total 10002432 --scenario GRADUAL,ALL --steps 60 --warmup 5 --no-extras --no-cpu-baseline > gpurun_out/r02u_bench_strong_n1.json 2> gpurun_out/r02u_bench_strong_n1.err
python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/r02x_smoke.log 2>&1
cat gpurun_out/r02x_pytest.log gpurun_out/r02x_smoke.log
python - <<'PY'
import json
for f in ["r02x_bench_k20","r02x_bench_k20_reference","r02x_bench_default","r02u_bench_strong_n1"]:
    try:
        d=json.loads(open("gpurun_out/%s.json"%f).read().strip().split("\n")[-1])
        print(f, "value %.3e (%.4f ms) e2e %.3e (%s ms)" % (d["value"], d["ms_per_step"], d["e2e"]["value"], d["e2e"].get("ms_per_step")))
    except Exception as e: print(f,"ERR",e)
PY
