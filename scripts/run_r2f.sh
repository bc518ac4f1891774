cd $GRAFT_REPO_ROOT
for n in 20000 320000; do
MCL_EXCHANGE=native timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29621 scripts/dist_check.py $n 2>&1 | grep -E "step|DIST_CHECK|Error|error" | tail -14
done
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29611 bench.py --gpus 2 --steps 100 --warmup 10 --no-extras > gpurun_out/r2f_bench2.json 2> gpurun_out/r2f_bench2.err
echo "rc $?"; tail -3 gpurun_out/r2f_bench2.err
python - <<'PY'
import json
d=json.loads(open('gpurun_out/r2f_bench2.json').read().strip().splitlines()[-1])
for k in ('value','ms_per_step','step_ms_median','step_ms_min','gpu_launches','parity'): print(k, d.get(k))
print('e2e', d['e2e'])
PY
MCL_NO_TAIL=1 timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29612 bench.py --gpus 2 --steps 100 --warmup 10 --no-extras --no-parity 2>/dev/null | python -c "import json,sys; d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print('notail', d['ms_per_step'], d['step_ms_median'], d['e2e']['ms_per_step'], d['gpu_launches'])"
