cd /root/repo
timeout 900 python -m pytest tests/test_gpu_parity.py -x -q 2>&1 | tail -5
for w in cfg3 cfg5 cfg2; do python bench.py --workload $w --steps 5 --warmup 3 2>&1 | python -c "import sys,json; d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print(d['config']['workload'][:5], d['value'], d['ms_per_step'], d['stages_ms'])"; done
