set -x
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1"
$TR --master-port 29701 bench.py --gpus 8 --steps 100 --warmup 10 --no-cpu > gpurun_out/bench_final_n8.json 2> gpurun_out/bench_final_n8.err
$TR --master-port 29702 bench.py --gpus 8 --steps 100 --warmup 10 --particles 1250000 --quick > gpurun_out/bench_final_n8_10M.json 2> gpurun_out/bench_final_n8_10M.err
$TR --master-port 29703 scripts/config5.py 6250000 > gpurun_out/config5_n8.log 2>&1
python -m torch.distributed.run --nnodes=1 --nproc-per-node 4 --master-addr 127.0.0.1 --master-port 29705 bench.py --gpus 4 --steps 100 --warmup 10 --quick > gpurun_out/bench_final_n4.json 2> gpurun_out/bench_final_n4.err
echo done
