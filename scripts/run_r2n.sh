cd $GRAFT_REPO_ROOT
ncu --metrics gpu__time_duration.sum --clock-control none -c 800 --csv --log-file gpurun_out/motion_new.csv python bench.py --steps 200 --warmup 10 --quick > /dev/null 2>&1
python - <<PY
import csv
rows=[r for r in csv.reader(open('gpurun_out/motion_new.csv')) if len(r)>5]
for name in ('k_motion','k_likelihood_g1','k_tail'):
    t=[float(r[-1].replace(',','')) for r in rows if name in r[4]]
    t=[x/1000 for x in t]
    print(name, len(t), 'mean %.1f'%(sum(t)/max(1,len(t))))
    print(' '.join('%.0f'%x for x in t[:240:4]))
PY
