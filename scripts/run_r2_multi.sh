# final multi-GPU measurements: bash scripts/run_r2_multi.sh N  (under gpurun --gpus N)
cd $GRAFT_REPO_ROOT
N=$1
timeout 1500 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29671 bench.py --gpus $N --steps 200 --warmup 10 > gpurun_out/r2_final_n$N.json 2> gpurun_out/r2_final_n$N.err
echo "bench n=$N rc $?"
python - <<PY
import json
d=json.loads(open('gpurun_out/r2_final_n$N.json').read().strip().splitlines()[-1])
print({k:d.get(k) for k in ('value','ms_per_step','step_ms_median','step_ms_other_resampling','gpu_launches')})
print('e2e', d['e2e']['ms_per_step'], d['e2e']['value'])
print('parity', d['parity'])
print(json.dumps(d['extras'])[:1200])
PY
