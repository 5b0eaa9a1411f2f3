cd /root/repo
for n in 8 4; do
  echo "== SGBM_NREG=$n"
  SGBM_NREG=$n python bench.py --steps 3 --warmup 3 2>&1 | python -c "import sys,json; d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print(d['value'], d['ms_per_step'], d['stages_ms'])"
done
python bench.py --steps 2 --warmup 3 > gpurun_out/plain.log 2>&1 && \
ncu --metrics gpu__time_duration.sum --clock-control none -c 200 --csv --log-file gpurun_out/launches_r1.csv python bench.py --steps 2 --warmup 3 > gpurun_out/ncu1.log 2>&1
python bench.py --steps 1 --warmup 3 > gpurun_out/plain2.log 2>&1 && \
ncu --set full --clock-control none --import-source on -k regex:'k_cost|k_vertical|k_horizontal' -s 8 -c 4 -o gpurun_out/prof_r1 python bench.py --steps 1 --warmup 3 > gpurun_out/ncu2.log 2>&1
ls -la gpurun_out/
