set -x
python -m pytest tests -m gpu -x -q > gpurun_out/t_r1m.log 2>&1; echo "tests rc=$?" >> gpurun_out/t_r1m.log
python bench.py --steps 20 --warmup 3 --mh-iters 32 --quick > gpurun_out/bench_m_mh32.json 2> gpurun_out/bench_m_mh32.err
MCL_NO_FUSE=1 python bench.py --steps 20 --warmup 3 --mh-iters 32 --quick > gpurun_out/bench_m_mh32_nofuse.json 2> gpurun_out/bench_m_mh32_nofuse.err
python bench.py --steps 20 --warmup 5 --quick > gpurun_out/bench_m_n1.json 2> gpurun_out/bench_m_n1.err
echo done
