cd $GRAFT_REPO_ROOT
timeout 900 python -m pytest tests/test_gpu_parity.py -x -q -m gpu -k "tail or fused or resampl or lockstep" 2>&1 | tail -3
timeout 600 python scripts/tail_span.py reference 1000000 210 2>&1 | awk 'NR<4 || NR%3==0' | cut -c1-260
for rep in 1 2; do
python bench.py --steps 200 --warmup 10 --quick 2>/dev/null | python -c "import json,sys; d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print('bench: mean', d['ms_per_step'], 'median', d.get('step_ms_median'), 'e2e ms', d['e2e']['ms_per_step'])"
done
