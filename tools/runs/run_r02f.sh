mkdir -p gpurun_out
O=gpurun_out/r02f_ab.txt; : > $O
run() { echo "== $1" >> $O; shift; "$@" 2>>gpurun_out/r02f.err | python -c "import sys,json; d=json.loads(sys.stdin.read()); print(round(d['us_per_day'],2), 'us/day', '%.3e'%d['agent_days_per_s'], d['us_per_day_by_window'][:12])" >> $O; }
for L in libdys_n00 libdys_n10 libdys_n11; do
  export DYS_LIB=$PWD/dysturbd_b200/$L.so
  run "$L 1M R1" python tools/quick_time.py --communities 89 --days 196 --warm 28
  run "$L 1M R4 plan+steal" env DYS_PLAN_MAX=100000000 python tools/quick_time.py --communities 89 --replicas 4 --days 196 --warm 28 --flags 0
  run "$L 1M R4 steal only" env DYS_PLAN_MAX=0 python tools/quick_time.py --communities 89 --replicas 4 --days 196 --warm 28 --flags 0
  run "$L 1M R4 plan only" env DYS_NO_STEAL=1 DYS_PLAN_MAX=100000000 python tools/quick_time.py --communities 89 --replicas 4 --days 196 --warm 28 --flags 0
  run "$L 10M steal only" env DYS_PLAN_MAX=0 python tools/quick_time.py --communities 890 --days 98 --warm 28
  run "$L 10M plan only" env DYS_NO_STEAL=1 DYS_PLAN_MAX=100000000 python tools/quick_time.py --communities 890 --days 98 --warm 28
done
unset DYS_LIB
run "r1 1M" python _r1/quick_time_r1.py --communities 89 --days 196 --warm 28
cat $O; tail -3 gpurun_out/r02f.err
