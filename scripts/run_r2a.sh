set -x
cd $GRAFT_REPO_ROOT
timeout 900 python -m pytest tests/test_gpu_parity.py -x -q -m gpu -k "tail_kernel or fused_step or fused_reference or diverge" 2>&1 | tail -40 > gpurun_out/r2a_tests.log
cat gpurun_out/r2a_tests.log
for mode in reference fixed; do
  timeout 300 python bench.py --steps 60 --warmup 5 --quick --resample $mode > gpurun_out/r2a_bench_$mode.json 2> gpurun_out/r2a_bench_$mode.err
  MCL_NO_TAIL=1 timeout 300 python bench.py --steps 60 --warmup 5 --quick --resample $mode > gpurun_out/r2a_bench_${mode}_notail.json 2> gpurun_out/r2a_bench_${mode}_notail.err
done
python - <<'PY'
import json,glob
for f in sorted(glob.glob('gpurun_out/r2a_bench_*.json')):
    try:
        d=json.loads(open(f).read().strip().splitlines()[-1])
        print(f, d['ms_per_step'], d['step_ms_median'], d['step_ms_min'], d['e2e']['ms_per_step'], d['gpu_launches'])
    except Exception as e:
        print(f, 'ERR', e)
PY
