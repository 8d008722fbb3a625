timeout 300 python -m pytest tests/test_gpu_ops.py tests/test_gpu_e2e.py -m gpu -x -q -k "forward_interpolate or warm_start or pipeline" 2>&1 | tail -3
for P in 1024 768 1536 512; do
PFB_STATS_PPB=$P timeout 200 ncu --profile-from-start off --metrics gpu__time_duration.sum --clock-control none -k regex:inorm_stats --csv --log-file gpurun_out/stats_ppb$P.csv python tools/profile_step.py --iters 1 > /dev/null 2>&1
python - <<PY
import csv
rows=list(csv.reader(open('gpurun_out/stats_ppb$P.csv')))
h=[i for i,r in enumerate(rows) if 'Kernel Name' in r][0]
mv=rows[h].index('Metric Value')
v=[float(r[mv].replace(',','')) for r in rows[h+1:] if len(r)>mv]
v=[x/1000 if x>1000 else x for x in v]
print('ppb',$P,'sum us',round(sum(v),1),[round(x,1) for x in v])
PY
done
