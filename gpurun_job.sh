cd /root/repo
timeout 600 python -m pytest tests -m gpu -x -q 2>&1 | tail -4
for c in cfg3 cfg4 cfg5 cfg2; do
timeout 200 python bench.py --workload $c --steps 8 --warmup 3 2>&1 | python -c "import sys,json; d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print('$c', round(d['value']), round(d['ms_per_step'],2), 'e2e', round(d['e2e']['value']), d['e2e'])"
done
