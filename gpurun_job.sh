cd /root/repo
timeout 900 python -m pytest tests/test_gpu_parity.py -x -q 2>&1 | tail -15
for r in 1 5 9 13; do SGBM_VR=$r python bench.py --workload cfg3 --steps 3 --warmup 3 2>&1 | python -c "import sys,json; d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print('R=$r', d['value'], d['ms_per_step'], d['stages_ms'])"; done
SGBM_VR=9 python bench.py --workload cfg2 --steps 3 --warmup 3 2>&1 | python -c "import sys,json; d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print('cfg2 R=9', d['value'], d['ms_per_step'], d['stages_ms'])"
SGBM_VR=9 python bench.py --workload cfg4 --steps 3 --warmup 3 2>&1 | python -c "import sys,json; d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print('cfg4 R=9', d['value'], d['ms_per_step'], d['stages_ms'])"
