cd $GRAFT_REPO_ROOT
MCL_LIK_SKEW=1 timeout 900 python -m pytest tests/test_gpu_parity.py -x -q -m gpu -k "likelihood or fused or lockstep or filter" 2>&1 | tail -3
for sk in 0 1 0 1; do
MCL_LIK_SKEW=$sk python bench.py --steps 200 --warmup 10 --quick 2>/dev/null | python -c "import json,sys; d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print('skew $sk bench: mean', d['ms_per_step'], 'median', d.get('step_ms_median'), 'lik ms', d['roofline']['launch_ms'])"
done
for sk in 0 1; do
MCL_LIK_SKEW=$sk ncu --metrics gpu__time_duration.sum,l1tex__data_pipe_lsu_wavefronts_mem_shared.sum --clock-control none -k regex:"k_likelihood_g1" -c 210 --csv --log-file gpurun_out/lik_skew$sk.csv python bench.py --steps 200 --warmup 10 --quick > /dev/null 2>&1
python - <<PY
import csv
rows=[r for r in csv.reader(open('gpurun_out/lik_skew$sk.csv')) if len(r)>5]
t=[float(r[-1].replace(',',''))/1000 for r in rows if 'gpu__time_duration' in r[-3]]
w=[float(r[-1].replace(',',''))/1e6 for r in rows if 'wavefronts' in r[-3]]
print('skew $sk time us:', ' '.join('%.0f'%x for x in t[::6]))
print('skew $sk Mwavefronts:', ' '.join('%.0f'%x for x in w[::6]))
PY
done
