set -x
python bench.py > gpurun_out/bench_final_n1.json 2> gpurun_out/bench_final_n1.err
MCL_LIK_VARIANT=5 python bench.py --steps 100 --warmup 10 --quick > gpurun_out/bench_o_v5.json 2> gpurun_out/bench_o_v5.err
MCL_LIK_VARIANT=6 python bench.py --steps 100 --warmup 10 --quick > gpurun_out/bench_o_v6.json 2> gpurun_out/bench_o_v6.err
python bench.py --steps 100 --warmup 10 --quick > gpurun_out/bench_o_v0.json 2> gpurun_out/bench_o_v0.err
echo done
