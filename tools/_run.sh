set -x
RBVFIT_B200_STREAM=1 python -m pytest tests -m gpu -q -x > gpurun_out/r02v_pytest_stream.log 2>&1; echo "rc=$?" >> gpurun_out/r02v_pytest_stream.log
python -m pytest tests -m gpu -q -x > gpurun_out/r02v_pytest_default.log 2>&1; echo "rc=$?" >> gpurun_out/r02v_pytest_default.log
{
python tools/profile_step.py --walkers 2048
python tools/profile_step.py --walkers 8192
python tools/profile_sightlines.py 256
python tools/profile_sightlines.py 1024
python tools/profile_step.py --workload C2 --walkers 1024
python tools/profile_step.py --workload C5a_L4 --walkers 8192
} > gpurun_out/r02v_perf.log 2>&1
