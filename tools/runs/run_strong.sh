# usage: bash tools/run_strong.sh N   (strong scaling of the 10 M-agent metro under ['GRADUAL','ALL'], BASELINE.json configs[2])
N=$1
mkdir -p gpurun_out
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29611"
timeout 1500 $TR bench.py --gpus $N --scaling strong --agents-total 10002432 --scenario GRADUAL,ALL --steps 60 --warmup 5 --no-extras > gpurun_out/r02u_bench_strong_n$N.json 2> gpurun_out/r02u_bench_strong_n$N.err
python - <<PY
import json
d=json.loads(open("gpurun_out/r02u_bench_strong_n$N.json").read().strip().split("\n")[-1])
print("strong N=$N value %.3e (%.4f ms) e2e %.3e (%.4f ms) verified %s" % (d["value"], d["ms_per_step"], d["e2e"]["value"], d["e2e"]["ms_per_step"], d["config"]["sharded_equals_single_gpu"]))
PY
grep -v "^W10\|^\*\*\*\|OMP_NUM\|^$\|Warning: WARNING\|NCCL version" gpurun_out/r02u_bench_strong_n$N.err | tail -3
