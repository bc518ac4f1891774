cd $GRAFT_REPO_ROOT
export CUDA_LAUNCH_BLOCKING=1
for n in 1000 5000; do
for stop in 31 33 34 35 36 37; do
    echo "== n $n stop $stop mode 0 (reference)"; MCL_TAIL_STOP=$stop timeout 120 python scripts/debug_tail.py 0 $n 2>&1 | tail -1
done
echo "== n $n stop 41 mode 1 (fixed)"; MCL_TAIL_STOP=41 timeout 120 python scripts/debug_tail.py 1 $n 2>&1 | tail -1
done
