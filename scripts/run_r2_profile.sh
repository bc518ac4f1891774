# round-2 ncu evidence (1 GPU): launch list of a short bench run + --set full of the four kernels of a step
cd $GRAFT_REPO_ROOT
TAG=${1:-r2d}
python bench.py --steps 30 --warmup 10 --quick > gpurun_out/plain_$TAG.log 2>&1 && \
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/launches_$TAG.csv python bench.py --steps 30 --warmup 10 --quick > gpurun_out/ncu_${TAG}_1.log 2>&1
echo "launch list rc $?"
ncu --set full --clock-control none --import-source on -k regex:"k_likelihood_g1|k_tail|k_motion" -s 120 -c 8 -f -o gpurun_out/prof_$TAG python bench.py --steps 30 --warmup 10 --quick > gpurun_out/ncu_${TAG}_2.log 2>&1
echo "full rc $?"; tail -3 gpurun_out/ncu_${TAG}_2.log; ls -la gpurun_out/prof_$TAG.ncu-rep
