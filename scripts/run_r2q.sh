cd $GRAFT_REPO_ROOT
timeout 900 python -m pytest tests/test_gpu_parity.py -x -q -m gpu -k "sharded_two_gpus" 2>&1 | tail -4
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29655 bench.py --gpus 2 --steps 200 --warmup 10 > gpurun_out/r2q_bench2.json 2> gpurun_out/r2q_bench2.err
echo "bench2 rc $?"; tail -c 1500 gpurun_out/r2q_bench2.json | head -c 1500
