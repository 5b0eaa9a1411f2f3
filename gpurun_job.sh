cd /root/repo
timeout 600 python -m pytest tests/test_gpu_parity.py -x -q 2>&1 | tail -15
run() { timeout 300 python bench.py --workload $1 --steps 3 --warmup 3 2>&1 | python -c "import sys,json; d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print('$2', d['value'], d['ms_per_step'], d['stages_ms'])"; }
run cfg3 "cfg3 default"
SGBM_VR=7 run cfg3 "cfg3 R7"
SGBM_VR=5 run cfg3 "cfg3 R5"
SGBM_VR=5 SGBM_NSTG=4 run cfg3 "cfg3 R5 nstg4"
run cfg5 cfg5
run cfg4 cfg4
run cfg2 cfg2
