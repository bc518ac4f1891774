set -x
python -m pytest tests -m gpu -x -q > gpurun_out/t_final.log 2>&1; echo "tests rc=$?" >> gpurun_out/t_final.log
python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/smoke_final.log 2>&1
python bench.py > gpurun_out/bench_final_n1.json 2> gpurun_out/bench_final_n1.err
python bench.py --steps 3 --warmup 10 --quick > gpurun_out/plain.log 2>&1 && ncu --metrics gpu__time_duration.sum --clock-control none -c 300 --csv --log-file gpurun_out/launches_r1g.csv python bench.py --steps 3 --warmup 10 --quick > gpurun_out/ncu16.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:"k_likelihood_g1|k_fz|k_motion" -s 36 -c 6 -o gpurun_out/prof_r1g python bench.py --steps 3 --warmup 10 --quick > gpurun_out/ncu17.log 2>&1
echo done
