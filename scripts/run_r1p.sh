set -x
python -m pytest tests -m gpu -x -q > gpurun_out/t_r1p.log 2>&1; echo "tests rc=$?" >> gpurun_out/t_r1p.log
python scripts/debug_motion2.py > gpurun_out/motion2.log 2>&1
python bench.py --steps 200 --warmup 10 --quick > gpurun_out/bench_p_n1.json 2> gpurun_out/bench_p_n1.err
echo done
