mkdir -p gpurun_out
(timeout 900 python -m pytest tests/test_gpu_checked.py tests/test_gpu_sharded.py -q 2>&1 | tail -15) > gpurun_out/r02l_pytest.log
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29611"
timeout 600 $TR bench.py --gpus 2 --steps 20 --warmup 5 > gpurun_out/r02l_bench_n2_k20.json 2> gpurun_out/r02l_bench_n2_k20.err
timeout 600 $TR bench.py --gpus 2 --steps 20 --warmup 5 --impl reference > gpurun_out/r02l_bench_n2_ref.json 2> gpurun_out/r02l_bench_n2_ref.err
timeout 600 $TR bench.py --gpus 2 --steps 20 --warmup 7 --replicas 8 > gpurun_out/r02l_bench_n2_rep8.json 2> gpurun_out/r02l_bench_n2_rep8.err
timeout 900 $TR bench.py --gpus 2 --steps 40 --warmup 5 --scaling strong --agents-total 2005504 --scenario GRADUAL,ALL > gpurun_out/r02l_bench_n2_strong.json 2> gpurun_out/r02l_bench_n2_strong.err
cat gpurun_out/r02l_pytest.log
for f in n2_k20 n2_ref n2_rep8 n2_strong; do echo "== $f"; tail -c 1200 gpurun_out/r02l_bench_$f.json; grep -v "^W10\|^\*\*\*\|OMP_NUM\|^$" gpurun_out/r02l_bench_$f.err | tail -4; done
