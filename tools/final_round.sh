set -u
python bench.py > gpurun_out/bench_r02j.json 2> gpurun_out/bench_r02j.err
python bench.py --impl reference > gpurun_out/bench_r02j_ref.json 2>> gpurun_out/bench_r02j.err
bash tools/ncu_round.sh > gpurun_out/ncu_round_r02j.log 2>&1
python -m pytest tests -m gpu -q > gpurun_out/gputest_r02j.log 2>&1
SB200_ARK=plain python -m pytest tests -m gpu -q > gpurun_out/gputest_r02j_plain.log 2>&1
tail -n 2 gpurun_out/gputest_r02j.log gpurun_out/gputest_r02j_plain.log; tail -c 300 gpurun_out/bench_r02j.json
