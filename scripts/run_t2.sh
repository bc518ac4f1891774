cd $GRAFT_REPO_ROOT
timeout 900 python -m pytest tests/test_gpu_parity.py -x -q -m gpu -k "tail or fused or resampl or lockstep or tiled or large_map" 2>&1 | tail -2
for i in 1 2; do python scripts/config5.py 6250000 2>&1 | tail -1; done
timeout 600 python scripts/tail_prof.py fixed 6250000 2>&1 | grep -E "kernel span|S5"
