set -x
python -m pytest tests -m gpu -x -q > gpurun_out/t_r1j.log 2>&1; echo "tests rc=$?" >> gpurun_out/t_r1j.log
python bench.py --steps 200 --warmup 10 --particles 1000 --quick > gpurun_out/bench_j_1k.json 2> gpurun_out/bench_j_1k.err
python bench.py --steps 200 --warmup 10 --particles 10000 --quick > gpurun_out/bench_j_10k.json 2> gpurun_out/bench_j_10k.err
python bench.py --steps 200 --warmup 10 --particles 100000 --quick > gpurun_out/bench_j_100k.json 2> gpurun_out/bench_j_100k.err
python scripts/debug_house.py > gpurun_out/house.log 2>&1
python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/smoke_j.log 2>&1
echo done
