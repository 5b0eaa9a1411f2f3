cd /root/repo
run() { timeout 120 python bench.py --workload $1 --steps 5 --warmup 3 2>&1 | python -c "import sys,json; d=json.loads(sys.stdin.read().strip().splitlines()[-1]); s=d['stages_ms']; print('$2', round(d['value']), round(d['ms_per_step'],2), d['e2e']['matches_device_path'], s.get('vertical_fwd'), s.get('vertical_wta'))"; }
timeout 600 python -m pytest tests -m gpu -x -q 2>&1 | tail -8
run cfg3 cfg3
run cfg2 cfg2
run cfg4 cfg4
