cd $GRAFT_REPO_ROOT
timeout 1500 python -m pytest tests -x -q -m gpu 2>&1 | tail -15 > gpurun_out/r2d_tests.log
cat gpurun_out/r2d_tests.log
timeout 900 python bench.py --steps 100 --warmup 10 > gpurun_out/r2d_bench.json 2> gpurun_out/r2d_bench.err
tail -3 gpurun_out/r2d_bench.err
python - <<'PY'
import json
d=json.loads(open('gpurun_out/r2d_bench.json').read().strip().splitlines()[-1])
for k in ('value','ms_per_step','step_ms_median','step_ms_min','step_ms_other_resampling','tail_kernel_error','gpu_launches','extras','cpu_baseline'):
    print(k, d.get(k))
print('e2e', d['e2e']); print('roofline frac', d['roofline']['frac'], d['roofline']['traffic'], d['roofline']['launch_ms']); print('gather', d['gather_roofline'].get('frac'))
PY
timeout 300 python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/r2d_ref.json 2>gpurun_out/r2d_ref.err; cat gpurun_out/r2d_ref.json | cut -c1-600
