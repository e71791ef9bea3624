set -x
RBVFIT_B200_STREAM=1 python -m pytest tests -m gpu -q -x > gpurun_out/r02q_pytest_stream.log 2>&1; echo "rc=$?" >> gpurun_out/r02q_pytest_stream.log
python -m pytest tests -m gpu -q -x > gpurun_out/r02q_pytest_default.log 2>&1; echo "rc=$?" >> gpurun_out/r02q_pytest_default.log
{
for w in 200 300 500 700 1024 2048 8192; do
python tools/profile_step.py --walkers $w
done
for w in 300 600 1024; do
python tools/profile_step.py --workload C2 --walkers $w
done
python tools/profile_sightlines.py 32
python tools/profile_sightlines.py 64
python tools/profile_sightlines.py 128
python tools/profile_sightlines.py 256
} > gpurun_out/r02q_perf.log 2>&1
