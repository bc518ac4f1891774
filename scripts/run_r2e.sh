cd $GRAFT_REPO_ROOT
timeout 900 python -m pytest tests/test_gpu_parity.py -x -q -m gpu -k "sharded_two_gpus" 2>&1 | tail -4
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29611 bench.py --gpus 2 --steps 100 --warmup 10 > gpurun_out/r2e_bench2.json 2> gpurun_out/r2e_bench2.err
echo "rc $?"; tail -3 gpurun_out/r2e_bench2.err
python - <<'PY'
import json
d=json.loads(open('gpurun_out/r2e_bench2.json').read().strip().splitlines()[-1])
for k in ('value','ms_per_step','step_ms_median','gpu_launches','extras','parity'): print(k, d.get(k))
print('e2e', d['e2e'])
PY
