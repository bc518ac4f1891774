set -x
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1"
$TR --master-port 29701 bench.py --gpus 8 --steps 100 --warmup 10 > gpurun_out/bench_r1f_n8.json 2> gpurun_out/bench_r1f_n8.err
$TR --master-port 29702 bench.py --gpus 8 --steps 100 --warmup 10 --particles 1250000 > gpurun_out/bench_r1f_n8_10M.json 2> gpurun_out/bench_r1f_n8_10M.err
$TR --master-port 29703 scripts/config5.py 6250000 > gpurun_out/config5_n8.log 2>&1
$TR --master-port 29704 scripts/dist_check.py 320000 > gpurun_out/dist_check_n8.log 2>&1
echo done
