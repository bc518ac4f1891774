set -x
python -m pytest tests -m gpu -x -q > gpurun_out/t_r1g.log 2>&1; echo "tests rc=$?" >> gpurun_out/t_r1g.log
python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29611 bench.py --gpus 2 --steps 100 --warmup 10 --quick > gpurun_out/bench_g_n2.json 2> gpurun_out/bench_g_n2.err
MCL_NO_FUSE=1 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29612 bench.py --gpus 2 --steps 100 --warmup 10 --quick > gpurun_out/bench_g_n2_nofuse.json 2> gpurun_out/bench_g_n2_nofuse.err
python bench.py --steps 100 --warmup 10 --quick > gpurun_out/bench_g_n1.json 2> gpurun_out/bench_g_n1.err
echo done
