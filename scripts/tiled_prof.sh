# per-launch time of the tiled likelihood kernel (6.25 M particles on the 4096^2 map) for several work-item sizes
cd $GRAFT_REPO_ROOT
for piece in 7168 3584 1792; do
MCL_TILED_PIECE=$piece ncu --metrics gpu__time_duration.sum --clock-control none -k regex:"k_likelihood_tiled" -c 16 --csv --log-file gpurun_out/tiled_$piece.csv python scripts/config5.py 6250000 > gpurun_out/tiled_$piece.log 2>&1
python - <<PY
import csv
rows=[r for r in csv.DictReader(l for l in open('gpurun_out/tiled_$piece.csv') if not l.startswith('=='))]
v=[float(r['Metric Value'].replace(',',''))/1e3 for r in rows]
print("piece=$piece n=%d mean %.1f us min %.1f max %.1f"%(len(v), sum(v)/len(v), min(v), max(v)))
PY
tail -2 gpurun_out/tiled_$piece.log
done
timeout 600 python -m pytest tests/test_gpu_parity.py -x -q -m gpu -k "tiled or large_map or config5 or 4096" 2>&1 | tail -3
