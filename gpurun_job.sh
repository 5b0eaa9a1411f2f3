cd /root/repo
timeout 300 python bench.py --workload cfg3 --steps 1 --warmup 3 > gpurun_out/plain.log 2>&1 && timeout 1200 ncu --set full --clock-control none --import-source on -k regex:'k_sweep' -s 2 -c 2 -o gpurun_out/prof_r1e_sweep -f python bench.py --workload cfg3 --steps 1 --warmup 3 > gpurun_out/ncu_fs.log 2>&1
grep -E "Profiling|ERROR" gpurun_out/ncu_fs.log
