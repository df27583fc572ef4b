mkdir -p gpurun_out
(timeout 900 python -m pytest tests/test_gpu_checked.py tests/test_gpu_sharded.py -q 2>&1 | tail -15) > gpurun_out/r02o_pytest.log
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29611"
timeout 600 $TR bench.py --gpus 2 --steps 20 --warmup 5 --no-extras > gpurun_out/r02o_bench_n2_k20.json 2> gpurun_out/r02o_bench_n2_k20.err
timeout 600 $TR bench.py --gpus 2 --steps 120 --warmup 5 --no-extras > gpurun_out/r02o_bench_n2_k120.json 2> gpurun_out/r02o_bench_n2_k120.err
cat gpurun_out/r02o_pytest.log
python - <<'PY'
import json
for f in ["n2_k20","n2_k120"]:
    try:
        d=json.loads(open("gpurun_out/r02o_bench_%s.json"%f).read().strip().split("\n")[-1])
        print("==",f,"value %.3e"%d["value"],"ms/step %.4f"%d["ms_per_step"],"e2e %.3e"%d["e2e"]["value"], "%.4f"%d["e2e"].get("ms_per_step"), d["config"]["sharded_equals_single_gpu"])
    except Exception as e: print(f, "ERR", e)
PY
grep -v "^W10\|^\*\*\*\|OMP_NUM\|^$\|Warning: WARNING" gpurun_out/r02o_bench_n2_k20.err | tail -5
