cd /root/repo
run() { timeout 120 python bench.py --workload $1 --steps 5 --warmup 3 2>&1 | python -c "import sys,json; d=json.loads(sys.stdin.read().strip().splitlines()[-1]); s=d['stages_ms']; print('$2', round(d['value']), round(d['ms_per_step'],2), d['e2e']['matches_device_path'], s.get('prefilter'), s.get('cost'))"; }
timeout 600 python -m pytest tests -m gpu -x -q 2>&1 | tail -5
run cfg3 cfg3
run cfg2 cfg2
run cfg4 cfg4
timeout 300 python bench.py --workload cfg3 --steps 1 --warmup 3 > gpurun_out/plain3.log 2>&1 && timeout 900 ncu --set full --clock-control none --import-source on -k regex:'k_cost2' -s 2 -c 1 -o gpurun_out/prof_cost2 -f python bench.py --workload cfg3 --steps 1 --warmup 3 > gpurun_out/ncu_c2.log 2>&1
