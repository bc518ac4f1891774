# round-2 ncu evidence (1 GPU): launch list of a short bench run + --set full of the three kernels of a step
cd $GRAFT_REPO_ROOT
python bench.py --steps 3 --warmup 10 --quick > gpurun_out/plain_r2.log 2>&1 && \
ncu --metrics gpu__time_duration.sum --clock-control none -c 300 --csv --log-file gpurun_out/launches_r2a.csv python bench.py --steps 3 --warmup 10 --quick > gpurun_out/ncu_r2_1.log 2>&1
echo "launch list rc $?"
ncu --set full --clock-control none --import-source on -k regex:"k_likelihood_g1|k_tail|k_motion" -s 30 -c 6 -o gpurun_out/prof_r2a python bench.py --steps 3 --warmup 10 --quick > gpurun_out/ncu_r2_2.log 2>&1
echo "full rc $?"; tail -3 gpurun_out/ncu_r2_2.log; ls -la gpurun_out/prof_r2a.ncu-rep
