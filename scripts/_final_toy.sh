set -x
mkdir -p gpurun_out
timeout 300 python bench.py > gpurun_out/r1c_bench.json 2> gpurun_out/r1c_bench.err || exit 1
timeout 300 python bench.py --impl reference > gpurun_out/r1c_bench_reference_arm.json 2> gpurun_out/r1c_ref.err
timeout 300 python bench.py --steps 2 --warmup 3 --no-cpu-baseline > gpurun_out/r1c_small.json 2>&1 || exit 1
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/r1c_bench_launches.csv python bench.py --steps 2 --warmup 3 --no-cpu-baseline > gpurun_out/r1c_ncu_launch.log 2>&1
timeout 900 bash scripts/ncu_capture.sh r1c_toy "toy_fused_kernel" 2 python bench.py --steps 2 --warmup 3 --no-cpu-baseline | tail -3
