set -x
python -m pytest tests -m gpu -x -q > gpurun_out/t_r1l.log 2>&1; echo "tests rc=$?" >> gpurun_out/t_r1l.log
python scripts/config5.py 6250000 > gpurun_out/config5_n1.log 2>&1
MCL_NO_TILED=1 python scripts/config5.py 6250000 > gpurun_out/config5_n1_notile.log 2>&1
python scripts/debug_motion2.py > gpurun_out/motion2.log 2>&1
echo done
