set -x
python -m pytest tests -m gpu -x -q > gpurun_out/t_verify.log 2>&1; echo "tests rc=$?" >> gpurun_out/t_verify.log
python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/smoke_verify.log 2>&1; echo "smoke rc=$?" >> gpurun_out/smoke_verify.log
echo done
