set -u
bash tools/ncu_capture.sh verify 21 "k_run|k_curve_p" verify_affine r02j 2
bash tools/ncu_capture.sh verify_vargen 20 "k_run|k_curve_p" verify_vargen_affine r02j 2
bash tools/ncu_capture.sh verify_double 20 "k_run|k_curve_p" verify_double_affine r02j 2
bash tools/ncu_capture.sh sign 20 'k_fixed_batch<\(int\)3' sign r02j 5
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/launches_r02j.csv python bench.py --steps 2 --warmup 3 --no-extras > gpurun_out/ncu_bench_r02j.log 2>&1
ls -la gpurun_out | tail -20
