set -x
python -m pytest tests -m gpu -x -q > gpurun_out/t_r1h.log 2>&1; echo "tests rc=$?" >> gpurun_out/t_r1h.log
python bench.py --steps 200 --warmup 10 --quick > gpurun_out/bench_h_n1.json 2> gpurun_out/bench_h_n1.err
ncu --metrics gpu__time_duration.sum --clock-control none -c 300 --csv --log-file gpurun_out/launches_r1h.csv python bench.py --steps 3 --warmup 10 --quick > gpurun_out/ncu13.log 2>&1
echo done
