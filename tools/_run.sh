set -x
python -m pytest tests -m gpu -q > gpurun_out/r02t_pytest_default.log 2>&1; echo "rc=$?" >> gpurun_out/r02t_pytest_default.log
python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/r02t_smoke.log 2>&1; echo "rc=$?" >> gpurun_out/r02t_smoke.log
