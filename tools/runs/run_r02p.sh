mkdir -p gpurun_out
(timeout 1500 python -m pytest tests -m gpu -q 2>&1 | tail -8) > gpurun_out/r02p_pytest.log
timeout 900 python bench.py --steps 20 --warmup 5 > gpurun_out/r02p_bench_k20.json 2> gpurun_out/r02p_bench_k20.err
python tools/quick_time.py --communities 89 --days 196 --warm 28 > gpurun_out/r02p_t1m.json 2>/dev/null
python _r1/quick_time_r1.py --communities 89 --days 196 --warm 28 > gpurun_out/r02p_t1m_r1.json 2>/dev/null
ncu --metrics gpu__time_duration.sum --clock-control none -k regex:k_days -s 4 -c 20 --csv --log-file gpurun_out/r02p_kdur_new.csv python tools/quick_time.py --communities 89 --days 168 --warm 28 > /dev/null 2>&1
ncu --metrics gpu__time_duration.sum --clock-control none -k regex:k_days -s 4 -c 20 --csv --log-file gpurun_out/r02p_kdur_r1.csv python _r1/quick_time_r1.py --communities 89 --days 168 --warm 28 > /dev/null 2>&1
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/r02p_launches.csv python bench.py --steps 20 --warmup 5 --no-extras --no-cpu-baseline --max-replays 3 > gpurun_out/r02p_ncu_bench.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:k_days -s 9 -c 1 -o gpurun_out/r02p_days_r16 python tools/quick_time.py --communities 89 --replicas 16 --days 70 --warm 28 --flags 0 > gpurun_out/r02p_ncu_r16.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:k_days -s 9 -c 1 -o gpurun_out/r02p_days_r1 python tools/quick_time.py --communities 89 --days 70 --warm 28 > gpurun_out/r02p_ncu_r1.log 2>&1
cat gpurun_out/r02p_pytest.log; cut -c1-150 gpurun_out/r02p_t1m.json gpurun_out/r02p_t1m_r1.json
python - <<'PY'
import csv, json
for f in ["new","r1"]:
    rows=[r for r in csv.reader(open("gpurun_out/r02p_kdur_%s.csv"%f)) if len(r)>5 and r[-1].replace('.','').replace(',','').isdigit()]
    v=[float(r[-1].replace(',','')) for r in rows]
    print(f, len(v), "mean us per launch", sum(v)/max(1,len(v))/ (1000 if max(v)>1e5 else 1))
d=json.loads(open("gpurun_out/r02p_bench_k20.json").read().strip().split("\n")[-1])
print("bench k20 value %.3e e2e %.3e (%.4f ms) lockdown %s / %s" % (d["value"], d["e2e"]["value"], d["e2e"]["ms_per_step"], d.get("e2e_lockdown",{}).get("ms_per_step"), d.get("e2e_lockdown_device",{}).get("ms_per_step")))
PY
tail -3 gpurun_out/r02p_bench_k20.err
