mkdir -p gpurun_out
(timeout 600 python -m pytest tests/test_gpu_replicas.py tests/test_gpu_sharded.py tests/test_gpu_checked.py -q -x 2>&1 | tail -4) > gpurun_out/r02v_pytest.log
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node 4 --master-addr 127.0.0.1 --master-port 29711"
timeout 600 $TR bench.py --gpus 4 --steps 20 --warmup 5 --no-extras > gpurun_out/r02v_bench_n4_k20.json 2> gpurun_out/r02v_bench_n4_k20.err
cat gpurun_out/r02v_pytest.log
python - <<'PY'
import json
d=json.loads(open("gpurun_out/r02v_bench_n4_k20.json").read().strip().split("\n")[-1])
print("weak N=4 value %.3e (%.4f ms) e2e %.3e (%.4f ms)" % (d["value"], d["ms_per_step"], d["e2e"]["value"], d["e2e"]["ms_per_step"]))
PY
bash tools/run_strong.sh 4
bash tools/run_strong.sh 2
