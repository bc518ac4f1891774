cd $GRAFT_REPO_ROOT
timeout 900 python -m pytest tests/test_gpu_parity.py tests/test_ros_adapter.py -x -q -m gpu -k "localizer or lockstep or filter or ros or step or converge" 2>&1 | tail -3
for rep in 1 2 3; do
python bench.py --steps 200 --warmup 10 --quick 2>/dev/null | python -c "import json,sys; d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print('bench: mean', d['ms_per_step'], 'median', d.get('step_ms_median'), 'e2e ms', d['e2e']['ms_per_step'])"
done
