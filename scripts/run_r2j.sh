cd $GRAFT_REPO_ROOT
timeout 600 python -m pytest tests/test_gpu_parity.py -x -q -m gpu -k "chain" 2>&1 | tail -2
for nt in 0 1; do
MCL_NO_TAIL=$nt python bench.py --steps 20 --warmup 3 --quick --mh-iters 32 2>/dev/null | python -c "import json,sys; d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print('1gpu notail=$nt chainx32', d['ms_per_step'], d['gpu_launches'])"
done
MCL_EXCHANGE=native timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29621 scripts/dist_check.py 320000 2>&1 | grep -E "chain|DIST_CHECK|rror" | tail -6
for nt in 0 1; do
MCL_NO_TAIL=$nt timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29612 bench.py --gpus 2 --steps 20 --warmup 3 --quick --no-parity --mh-iters 32 2>/dev/null | python -c "import json,sys; d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print('2gpu notail=$nt chainx32', d['ms_per_step'], d['gpu_launches'])"
done
