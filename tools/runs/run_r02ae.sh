mkdir -p gpurun_out
(timeout 900 python -m pytest tests -m gpu -x -q 2>&1 | tail -15) > gpurun_out/r02ae_pytest.log
python bench.py --steps 20 --warmup 5 > gpurun_out/r02ae_bench_k20.json 2> gpurun_out/r02ae_bench_k20.err
timeout 150 python tools/network_time.py --ip 2>&1 | tail -1 > gpurun_out/r02ae_network_time.json
timeout 120 ncu --set full --clock-control none --import-source on -k regex:k_sim_edges -c 3 -o gpurun_out/r02ae_sim_edges python tools/network_probe.py > gpurun_out/r02ae_ncu1.log 2>&1
timeout 120 ncu --set full --clock-control none --import-source on -k regex:k_sp_relax -s 30 -c 4 -o gpurun_out/r02ae_sp_relax python tools/network_probe.py > gpurun_out/r02ae_ncu2.log 2>&1
cat gpurun_out/r02ae_pytest.log; tail -c 1500 gpurun_out/r02ae_bench_k20.json; echo; cat gpurun_out/r02ae_network_time.json; tail -n 2 gpurun_out/r02ae_ncu1.log gpurun_out/r02ae_ncu2.log
