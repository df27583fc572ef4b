mkdir -p gpurun_out
O=gpurun_out/r02g_ab.txt; : > $O
run() { echo "== $1" >> $O; shift; "$@" 2>>gpurun_out/r02g.err | python -c "import sys,json; d=json.loads(sys.stdin.read()); print(round(d['us_per_day'],2), 'us/day', '%.3e'%d['agent_days_per_s'], d['us_per_day_by_window'][:12])" >> $O; }
for L in libdysturbd libdys_b1p0 libdys_b0p0; do
  export DYS_LIB=$PWD/dysturbd_b200/$L.so
  run "$L 1M R1" python tools/quick_time.py --communities 89 --days 196 --warm 28
  run "$L 1M R4" python tools/quick_time.py --communities 89 --replicas 4 --days 196 --warm 28 --flags 0
  run "$L 10M" python tools/quick_time.py --communities 890 --days 98 --warm 28
done
export DYS_LIB=$PWD/dysturbd_b200/libdysturbd.so
run "libdysturbd 1M R16" python tools/quick_time.py --communities 89 --replicas 16 --days 196 --warm 28 --flags 0
run "libdysturbd 1M R64" python tools/quick_time.py --communities 89 --replicas 64 --days 98 --warm 28 --flags 0
DYS_TRACE=/tmp/tr10.txt DYS_NO_STEAL=1 DYS_PLAN_MAX=100000000 python tools/quick_time.py --communities 890 --days 70 --warm 28 > /dev/null 2>&1; python tools/trace_stats.py /tmp/tr10.txt > gpurun_out/r02g_trace_10m_planonly.txt 2>&1
cat $O gpurun_out/r02g_trace_10m_planonly.txt; tail -3 gpurun_out/r02g.err
