set -u
bash tools/ncu_capture.sh verify 21 k_run verify_affine r02h 2
bash tools/ncu_capture.sh verify_vargen 20 k_run verify_vargen_affine r02h 2
bash tools/ncu_capture.sh verify_double 20 k_run verify_double_affine r02h 2
bash tools/ncu_capture.sh sign 20 'k_fixed_batch<\(int\)3' sign r02h 5
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/launches_r02h.csv python bench.py --steps 2 --warmup 3 --no-extras > gpurun_out/ncu_bench_r02h.log 2>&1
ls -la gpurun_out | tail -20
