cd $GRAFT_REPO_ROOT
timeout 1800 python -m pytest tests -x -q -m gpu 2>&1 | tail -6
python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -2
python bench.py --steps 200 --warmup 10 --quick 2>/dev/null | python -c "import json,sys; d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print('bench', d['ms_per_step'], d['step_ms_median'], d['step_ms_min'], d['e2e']['ms_per_step'])"
