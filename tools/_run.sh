set -x
python -m pytest tests -m gpu -q -x > gpurun_out/r02n_pytest_default.log 2>&1; echo "rc=$?" >> gpurun_out/r02n_pytest_default.log
python bench.py --no-cpu > gpurun_out/r02n_bench.json 2> gpurun_out/r02n_bench.err
