set -x
python -m pytest tests -m gpu -x -q > gpurun_out/t_verify.log 2>&1; echo "tests rc=$?" >> gpurun_out/t_verify.log
python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/smoke_verify.log 2>&1; echo "smoke rc=$?" >> gpurun_out/smoke_verify.log
python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29611 bench.py --gpus 2 --steps 100 --warmup 10 --quick > gpurun_out/bench_verify_n2.json 2> gpurun_out/bench_verify_n2.err
echo done
