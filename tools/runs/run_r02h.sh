mkdir -p gpurun_out
(timeout 1500 python -m pytest tests -m gpu -x -q 2>&1 | tail -25) > gpurun_out/r02h_pytest.log
O=gpurun_out/r02h_ab.txt; : > $O
run() { echo "== $1" >> $O; shift; "$@" 2>>gpurun_out/r02h.err | python -c "import sys,json; d=json.loads(sys.stdin.read()); print(round(d['us_per_day'],2), 'us/day', '%.3e'%d['agent_days_per_s'], d['us_per_day_by_window'][:12])" >> $O; }
for L in libdysturbd libdys_bin; do
  export DYS_LIB=$PWD/dysturbd_b200/$L.so
  run "$L 1M R1" python tools/quick_time.py --communities 89 --days 196 --warm 28
  run "$L 1M R4" python tools/quick_time.py --communities 89 --replicas 4 --days 196 --warm 28 --flags 0
  run "$L 10M" python tools/quick_time.py --communities 890 --days 98 --warm 28
done
unset DYS_LIB
run "r1 1M" python _r1/quick_time_r1.py --communities 89 --days 196 --warm 28
run "libdysturbd 1M R16" python tools/quick_time.py --communities 89 --replicas 16 --days 196 --warm 28 --flags 0
cat gpurun_out/r02h_pytest.log $O; tail -3 gpurun_out/r02h.err
