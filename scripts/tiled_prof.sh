cd $GRAFT_REPO_ROOT
for old in 0 1; do
MCL_TILED_OLD=$old ncu --metrics gpu__time_duration.sum --clock-control none -k regex:"k_likelihood_tiled" -c 16 --csv --log-file gpurun_out/tiled_$old.csv python scripts/config5.py 6250000 > /dev/null 2>&1
python - <<PY
import csv
rows=[r for r in csv.DictReader(l for l in open('gpurun_out/tiled_$old.csv') if not l.startswith('=='))]
v=[float(r['Metric Value'].replace(',',''))/1e3 for r in rows]
print("old=$old n=%d mean %.1f us min %.1f max %.1f"%(len(v), sum(v)/len(v), min(v), max(v)))
PY
done
