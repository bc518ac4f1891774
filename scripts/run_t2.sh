cd $GRAFT_REPO_ROOT
timeout 600 python -m pytest tests/test_gpu_parity.py -x -q -m gpu -k "tiled or large_map" 2>&1 | tail -2
ncu --metrics gpu__time_duration.sum --clock-control none -k regex:"k_likelihood_tiled" -c 16 --csv --log-file gpurun_out/tiled_x.csv python scripts/config5.py 6250000 > gpurun_out/tiled_x.log 2>&1
python - <<PY
import csv
rows=[r for r in csv.DictReader(l for l in open('gpurun_out/tiled_x.csv') if not l.startswith('=='))]
v=[float(r['Metric Value'].replace(',',''))/1e3 for r in rows]
print("n=%d mean %.1f us min %.1f max %.1f"%(len(v), sum(v)/len(v), min(v), max(v)))
PY
python scripts/config5.py 6250000 2>&1 | tail -1
