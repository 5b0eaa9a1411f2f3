cd /root/repo
SGBM_SWEEP_TRACE=gpurun_out/trace_cfg2.bin timeout 120 python bench.py --workload cfg2 --steps 1 --warmup 3 2>&1 | grep "sweep trace" | tail -1
SGBM_SWEEP_TRACE=gpurun_out/trace_cfg3.bin timeout 120 python bench.py --workload cfg3 --steps 1 --warmup 3 2>&1 | grep "sweep trace" | tail -2
