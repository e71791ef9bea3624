set -x
python -m pytest tests -m gpu -q > gpurun_out/r02k_pytest_default.log 2>&1; echo "rc=$?" >> gpurun_out/r02k_pytest_default.log
( time python bench.py ) > gpurun_out/r02k_bench.json 2> gpurun_out/r02k_bench.err
( time python bench.py --impl reference --steps 3 --warmup 1 ) > gpurun_out/r02k_bench_ref.json 2> gpurun_out/r02k_bench_ref.err
