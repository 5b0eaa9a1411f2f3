cd /root/repo
timeout 600 python -m pytest tests -m gpu -x -q 2>&1 | tail -5
for wl in cfg3 cfg4 cfg2 cfg5; do
  SGBM_SWEEP_VERBOSE=1 timeout 300 python bench.py --workload $wl --steps 6 --warmup 3 2>gpurun_out/err_$wl.log | python -c "
import sys,json
for l in sys.stdin:
    try: j=json.loads(l)
    except Exception: continue
    print('$wl', round(j['ms_per_step'],3),'ms', round(j['value']),'MDE/s e2e',round(j['e2e']['value']), j['stages_ms'], j['e2e'].get('matches_device_path'))
"
grep -m2 "sweep:" gpurun_out/err_$wl.log
done
