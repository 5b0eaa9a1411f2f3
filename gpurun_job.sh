cd /root/repo
run() { timeout 120 python bench.py --workload $1 --steps 5 --warmup 3 2>&1 | python -c "import sys,json; d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print('$2', round(d['value']), round(d['ms_per_step'],2), d['e2e']['matches_device_path'], d['stages_ms'])"; }
timeout 300 python -m pytest tests/test_gpu_parity.py -x -q -k "stages_vs_oracle" 2>&1 | tail -15
run cfg2 cfg2
run cfg3 cfg3
timeout 600 python -m pytest tests -m gpu -x -q 2>&1 | tail -15
run cfg4 cfg4
