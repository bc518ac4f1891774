set -x
python -m pytest tests -m gpu -x -q > gpurun_out/t_r1n.log 2>&1; echo "tests rc=$?" >> gpurun_out/t_r1n.log
python scripts/debug_house.py > gpurun_out/house.log 2>&1
python bench.py --steps 50 --warmup 10 --quick > gpurun_out/bench_n_n1.json 2> gpurun_out/bench_n_n1.err
echo done
