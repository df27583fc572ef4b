mkdir -p gpurun_out
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node 4 --master-addr 127.0.0.1 --master-port 29611"
timeout 600 $TR bench.py --gpus 4 --scaling strong --agents-total 10002432 --steps 60 --warmup 5 --no-extras --max-replays 5 > gpurun_out/r02y_strong_n4_nolock.json 2> gpurun_out/r02y_strong_n4_nolock.err
DYS_NO_STEAL=1 timeout 600 $TR bench.py --gpus 4 --scaling strong --agents-total 10002432 --scenario GRADUAL,ALL --steps 60 --warmup 5 --no-extras --max-replays 5 > gpurun_out/r02y_strong_n4_nosteal.json 2> gpurun_out/r02y_strong_n4_nosteal.err
python - <<'PY'
import json
for f in ["nolock","nosteal"]:
    try:
        d=json.loads(open("gpurun_out/r02y_strong_n4_%s.json"%f).read().strip().split("\n")[-1])
        print(f, "value %.3e (%.4f ms) e2e %.3e (%s ms) ever %s" % (d["value"], d["ms_per_step"], d["e2e"]["value"], d["e2e"].get("ms_per_step"), d["config"]["ever_infected_at_end"]))
    except Exception as e: print(f,"ERR",e)
PY
