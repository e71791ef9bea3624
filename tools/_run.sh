set -x
RBVFIT_B200_STREAM=1 python -m pytest tests -m gpu -q -x > gpurun_out/r02s_pytest_stream.log 2>&1; echo "rc=$?" >> gpurun_out/r02s_pytest_stream.log
python -m pytest tests -m gpu -q -x > gpurun_out/r02s_pytest_default.log 2>&1; echo "rc=$?" >> gpurun_out/r02s_pytest_default.log
{
python tools/profile_step.py --walkers 2048
python tools/profile_step.py --walkers 8192
python tools/profile_sightlines.py 256
python tools/profile_step.py --workload C2 --walkers 1024
python tools/profile_step.py --workload C5a_L4 --walkers 8192
C5A_W=2048 python tools/check_farfield.py
} > gpurun_out/r02s_perf.log 2>&1
python tools/profile_step.py --walkers 2048 --steps 3 > gpurun_out/r02s_plain.log 2>&1 &&
ncu --set full --clock-control none --import-source on -k regex:voigt_stream -s 1 -c 1 -f -o gpurun_out/prof_r02s_stream python tools/profile_step.py --walkers 2048 --steps 3 > gpurun_out/r02s_ncu.log 2>&1
