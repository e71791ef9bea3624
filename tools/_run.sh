set -x
RBVFIT_B200_STREAM=1 python -m pytest tests -m gpu -q -x > gpurun_out/r02c_pytest_stream.log 2>&1; echo "rc=$?" >> gpurun_out/r02c_pytest_stream.log
python -m pytest tests -m gpu -q -x > gpurun_out/r02c_pytest_default.log 2>&1; echo "rc=$?" >> gpurun_out/r02c_pytest_default.log
{
for s in 0 4 8 16; do echo "SEGS=$s"; RBVFIT_B200_STREAM_SEGS=$s python tools/profile_step.py --walkers 2048; done
echo "u3"; RBVFIT_B200_LIB=build/variants/lib_u3.so RBVFIT_B200_STREAM_SEGS=16 python tools/profile_step.py --walkers 2048
RBVFIT_B200_STREAM=0 python tools/profile_step.py --walkers 8192
python tools/profile_step.py --walkers 8192
RBVFIT_B200_STREAM_SEGS=16 python tools/profile_step.py --walkers 8192
RBVFIT_B200_STREAM=0 python tools/profile_sightlines.py 256
python tools/profile_sightlines.py 256
RBVFIT_B200_STREAM=0 python tools/profile_step.py --workload C2 --walkers 1024
RBVFIT_B200_STREAM=1 python tools/profile_step.py --workload C2 --walkers 1024
RBVFIT_B200_STREAM=0 python tools/profile_step.py --workload C3 --walkers 1024
RBVFIT_B200_STREAM=1 python tools/profile_step.py --workload C3 --walkers 1024
RBVFIT_B200_STREAM=0 python tools/profile_step.py --workload C5a_L4 --walkers 2048
RBVFIT_B200_STREAM=1 python tools/profile_step.py --workload C5a_L4 --walkers 2048
} > gpurun_out/r02c_perf.log 2>&1
export RBVFIT_B200_STREAM_SEGS=16
python tools/profile_step.py --walkers 2048 --steps 3 > gpurun_out/r02c_plain.log 2>&1 &&
ncu --set full --clock-control none --import-source on -k regex:voigt_stream -s 1 -c 1 -f -o gpurun_out/prof_r02c_stream python tools/profile_step.py --walkers 2048 --steps 3 > gpurun_out/r02c_ncu.log 2>&1
