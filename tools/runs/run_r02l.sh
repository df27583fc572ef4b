mkdir -p gpurun_out
(timeout 900 python -m pytest tests/test_gpu_checked.py tests/test_gpu_sharded.py -q 2>&1 | tail -15) > gpurun_out/r02n_pytest.log
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29611"
timeout 600 $TR bench.py --gpus 2 --steps 20 --warmup 5 > gpurun_out/r02n_bench_n2_k20.json 2> gpurun_out/r02n_bench_n2_k20.err

timeout 600 $TR bench.py --gpus 2 --steps 120 --warmup 5 --no-extras > gpurun_out/r02n_bench_n2_k120.json 2> gpurun_out/r02n_bench_n2_k120.err
timeout 900 $TR bench.py --gpus 2 --steps 40 --warmup 5 --scaling strong --agents-total 2004992 --scenario GRADUAL,ALL > gpurun_out/r02n_bench_n2_strong.json 2> gpurun_out/r02n_bench_n2_strong.err
cat gpurun_out/r02n_pytest.log
for f in n2_k20 n2_k120 n2_strong; do echo "== $f"; tail -c 1200 gpurun_out/r02n_bench_$f.json; grep -v "^W10\|^\*\*\*\|OMP_NUM\|^$" gpurun_out/r02n_bench_$f.err | tail -4; done
python tools/quick_time.py --communities 89 --days 196 --warm 28 > gpurun_out/r02n_t1m.json 2>/dev/null; python _r1/quick_time_r1.py --communities 89 --days 196 --warm 28 > gpurun_out/r02n_t1m_r1.json 2>/dev/null; cat gpurun_out/r02n_t1m.json gpurun_out/r02n_t1m_r1.json | cut -c1-200
