mkdir -p gpurun_out
(timeout 1200 python -m pytest tests -m gpu -x -q 2>&1 | tail -25) > gpurun_out/r02e_pytest.log
O=gpurun_out/r02e_ab.txt; : > $O
run() { echo "== $1" >> $O; shift; "$@" 2>>gpurun_out/r02e.err | python -c "import sys,json; d=json.loads(sys.stdin.read()); print(round(d['us_per_day'],2), 'us/day', '%.3e'%d['agent_days_per_s'], d['us_per_day_by_window'][:12])" >> $O; }
run "1M R1" python tools/quick_time.py --communities 89 --days 196 --warm 28
run "1M R4 steal" python tools/quick_time.py --communities 89 --replicas 4 --days 196 --warm 28 --flags 0
run "1M R4 nosteal" env DYS_NO_STEAL=1 python tools/quick_time.py --communities 89 --replicas 4 --days 196 --warm 28 --flags 0
run "1M R16 steal" python tools/quick_time.py --communities 89 --replicas 16 --days 196 --warm 28 --flags 0
run "1M R16 nosteal" env DYS_NO_STEAL=1 python tools/quick_time.py --communities 89 --replicas 16 --days 196 --warm 28 --flags 0
run "10M steal" python tools/quick_time.py --communities 890 --days 98 --warm 28
run "10M nosteal" env DYS_NO_STEAL=1 python tools/quick_time.py --communities 890 --days 98 --warm 28
cat gpurun_out/r02e_pytest.log $O; tail -3 gpurun_out/r02e.err
