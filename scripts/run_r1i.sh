set -x
python -m pytest tests -m gpu -x -q > gpurun_out/t_r1i.log 2>&1; echo "tests rc=$?" >> gpurun_out/t_r1i.log
python bench.py --steps 200 --warmup 10 --quick > gpurun_out/bench_i_dup.json 2> gpurun_out/bench_i_dup.err
MCL_LIK_NODUP=1 python bench.py --steps 200 --warmup 10 --quick > gpurun_out/bench_i_nodup.json 2> gpurun_out/bench_i_nodup.err
python scripts/profile_house.py > gpurun_out/house.log 2>&1
echo done
