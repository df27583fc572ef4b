mkdir -p gpurun_out
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29711"
timeout 600 $TR bench.py --gpus 8 --steps 20 --warmup 5 --no-extras > gpurun_out/r02w_bench_n8_k20.json 2> gpurun_out/r02w_bench_n8_k20.err
timeout 600 $TR bench.py --gpus 8 --steps 120 --warmup 5 --no-extras > gpurun_out/r02w_bench_n8_k120.json 2> gpurun_out/r02w_bench_n8_k120.err
timeout 900 $TR bench.py --gpus 8 --steps 30 --warmup 5 --communities-per-gpu 1110 --no-extras > gpurun_out/r02w_bench_n8_100m.json 2> gpurun_out/r02w_bench_n8_100m.err
python - <<'PY'
import json
for f in ["n8_k20","n8_k120","n8_100m"]:
    try:
        d=json.loads(open("gpurun_out/r02w_bench_%s.json"%f).read().strip().split("\n")[-1])
        print("==",f,"value %.3e"%d["value"],"ms/step %.4f"%d["ms_per_step"],"e2e %.3e"%d["e2e"]["value"], "%.4f"%d["e2e"].get("ms_per_step"), d["config"]["agents_total"], d["config"]["timing"][-60:])
    except Exception as e: print(f, "ERR", e)
PY
for f in n8_k20 n8_k120 n8_100m; do grep -v "^W10\|^\*\*\*\|OMP_NUM\|^$\|Warning: WARNING\|NCCL version" gpurun_out/r02w_bench_$f.err | tail -3; done
