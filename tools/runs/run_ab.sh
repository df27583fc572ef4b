mkdir -p gpurun_out
O=gpurun_out/r02d_ab.txt; : > $O
run() { echo "== $1" >> $O; shift; shift; "$@" 2>>gpurun_out/r02d.err | python -c "import sys,json; d=json.loads(sys.stdin.read()); print(round(d['us_per_day'],2), 'us/day', '%.3e'%d['agent_days_per_s'], d['us_per_day_by_window'][:12])" >> $O; }
run "r1 1M" x python _r1/quick_time_r1.py --communities 89 --days 196 --warm 28
for L in libdysturbd libdys_s0p0 libdys_s1p0 libdys_s0p1; do
  export DYS_LIB=$PWD/dysturbd_b200/$L.so
  run "$L 1M R1" x python tools/quick_time.py --communities 89 --days 196 --warm 28
  run "$L 1M R4" x python tools/quick_time.py --communities 89 --replicas 4 --days 196 --warm 28 --flags 0
done
unset DYS_LIB
run "cur 1M R16" x python tools/quick_time.py --communities 89 --replicas 16 --days 196 --warm 28 --flags 0
run "cur 1M R16 nosteal" x env DYS_NO_STEAL=1 python tools/quick_time.py --communities 89 --replicas 16 --days 196 --warm 28 --flags 0
run "r1 10M" x python _r1/quick_time_r1.py --communities 890 --days 98 --warm 28
for L in libdysturbd libdys_s0p0; do
  export DYS_LIB=$PWD/dysturbd_b200/$L.so
  run "$L 10M" x python tools/quick_time.py --communities 890 --days 98 --warm 28
done
unset DYS_LIB
cat $O; tail -3 gpurun_out/r02d.err
