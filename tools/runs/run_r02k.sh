mkdir -p gpurun_out
timeout 900 python bench.py --steps 20 --warmup 5 > gpurun_out/r02k_bench_k20.json 2> gpurun_out/r02k_bench_k20.err
timeout 900 python bench.py --steps 20 --warmup 5 --impl reference > gpurun_out/r02k_bench_ref_k20.json 2> gpurun_out/r02k_bench_ref_k20.err
timeout 1200 python bench.py > gpurun_out/r02k_bench_default.json 2> gpurun_out/r02k_bench_default.err
tail -c 3000 gpurun_out/r02k_bench_k20.json; tail -5 gpurun_out/r02k_bench_k20.err; tail -c 800 gpurun_out/r02k_bench_ref_k20.json; tail -c 3000 gpurun_out/r02k_bench_default.json; tail -5 gpurun_out/r02k_bench_default.err
