set -x
python -m pytest tests -m gpu -x -q > gpurun_out/t_r1k.log 2>&1; echo "tests rc=$?" >> gpurun_out/t_r1k.log
python scripts/debug_house.py > gpurun_out/house.log 2>&1
python bench.py --steps 100 --warmup 10 --quick > gpurun_out/bench_k_n1.json 2> gpurun_out/bench_k_n1.err
echo done
