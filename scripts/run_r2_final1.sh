# final 1-GPU measurements of round 2: tests, smoke, both bench arms
cd $GRAFT_REPO_ROOT
timeout 1500 python -m pytest tests -x -q -m gpu 2>&1 | tail -3
python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -2
python bench.py > gpurun_out/r2_final_n1.json 2> gpurun_out/r2_final_n1.err; echo "bench rc $?"
python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/r2_final_ref.json 2> gpurun_out/r2_final_ref.err; echo "ref rc $?"
tail -c 600 gpurun_out/r2_final_ref.json
