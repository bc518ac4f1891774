cd $GRAFT_REPO_ROOT
for n in 20000 320000 1000000; do
MCL_EXCHANGE=native MCL_RESAMPLE=reference timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29621 scripts/dist_check.py $n 2>&1 | grep -E "step|DIST_CHECK|Error|error|rror" | tail -8
done
for mode in fixed reference; do
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29612 bench.py --gpus 2 --steps 100 --warmup 10 --no-extras --no-parity --resample $mode 2>/dev/null | python -c "import json,sys; d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print('$mode', d['ms_per_step'], d['step_ms_median'], d['e2e']['ms_per_step'], d['gpu_launches'])"
done
