set -x
python -m pytest tests -m gpu -x -q > gpurun_out/t_r1f.log 2>&1; echo "tests rc=$?" >> gpurun_out/t_r1f.log
for v in 0 1 2 3; do
  MCL_LIK_VARIANT=$v python bench.py --steps 100 --warmup 10 --quick > gpurun_out/bench_f_v$v.json 2> gpurun_out/bench_f_v$v.err
done
MCL_NO_FUSE=1 python bench.py --steps 100 --warmup 10 --quick > gpurun_out/bench_f_nofuse.json 2> gpurun_out/bench_f_nofuse.err
echo done
