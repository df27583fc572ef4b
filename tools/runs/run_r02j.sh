mkdir -p gpurun_out
(timeout 900 python -m pytest tests/test_gpu_sharded.py -q 2>&1 | tail -25) > gpurun_out/r02j_pytest_sharded.log
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29611"
timeout 600 $TR bench.py --gpus 2 --steps 20 --warmup 5 > gpurun_out/r02j_bench_n2_k20.json 2> gpurun_out/r02j_bench_n2_k20.err
timeout 600 $TR bench.py --gpus 2 --steps 120 --warmup 5 > gpurun_out/r02j_bench_n2_k120.json 2> gpurun_out/r02j_bench_n2_k120.err
timeout 900 compute-sanitizer --tool memcheck --target-processes all $TR tools/sanitize_run.py --peer > gpurun_out/r02j_sanitizer_memcheck_peer.txt 2>&1
cat gpurun_out/r02j_pytest_sharded.log; tail -c 1500 gpurun_out/r02j_bench_n2_k20.json; tail -5 gpurun_out/r02j_bench_n2_k20.err; tail -c 600 gpurun_out/r02j_bench_n2_k120.json; tail -3 gpurun_out/r02j_bench_n2_k120.err; tail -8 gpurun_out/r02j_sanitizer_memcheck_peer.txt
