set -x
python bench.py > gpurun_out/r02_final_bench_n1.json 2> gpurun_out/r02_final_bench_n1.err
python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/r02_final_bench_ref.json 2> gpurun_out/r02_final_bench_ref.err
python bench.py --steps 3 --warmup 3 --no-cpu --no-extras > gpurun_out/r02_plain.log 2>&1 &&
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/r02_launch_list_bench.csv python bench.py --steps 3 --warmup 3 --no-cpu --no-extras > gpurun_out/r02_ncu_launches.log 2>&1
