cd /root/repo
timeout 600 python -m pytest tests/test_gpu_parity.py -x -q 2>&1 | tail -15
run() { timeout 300 python bench.py --workload $1 --steps 3 --warmup 3 2>&1 | python -c "import sys,json; d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print('$2', d['value'], d['ms_per_step'], d['stages_ms'])"; }
run cfg3 cfg3
run cfg2 cfg2
