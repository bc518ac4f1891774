cd $GRAFT_REPO_ROOT
timeout 900 python -m pytest tests/test_gpu_parity.py -x -q -m gpu -k "motion" 2>&1 | tail -5
for rep in 1 2; do
python bench.py --steps 200 --warmup 10 --quick 2>/dev/null | python -c "import json,sys; d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print('bench: mean', d['ms_per_step'], 'median', d.get('step_ms_median'), 'e2e ms', d['e2e']['ms_per_step'], 'launches', d['gpu_launches'])"
done
ncu --metrics gpu__time_duration.sum --clock-control none -c 1100 --csv --log-file gpurun_out/motion_new.csv python bench.py --steps 200 --warmup 10 --quick > /dev/null 2>&1
python - <<PY
import csv
rows=[r for r in csv.reader(open('gpurun_out/motion_new.csv')) if len(r)>5]
for name in ('k_motion(','k_motion_retry','k_likelihood_g1','k_tail'):
    t=[float(r[-1].replace(',','')) for r in rows if name in r[4]]
    t=[x/1000 for x in t]
    print(name, len(t), 'mean %.1f'%(sum(t)/max(1,len(t))))
    print(' '.join('%.0f'%x for x in t[:240:4]))
PY
ncu --set full --clock-control none --import-source on -k regex:"k_motion_retry" -s 50 -c 1 -f -o gpurun_out/prof_retry python bench.py --steps 60 --warmup 10 --quick > gpurun_out/ncu_retry.log 2>&1
ls -la gpurun_out/prof_retry.ncu-rep
