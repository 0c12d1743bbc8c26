set -u
python bench.py > gpurun_out/bench_r02i.json 2> gpurun_out/bench_r02i.err
python bench.py --impl reference > gpurun_out/bench_r02i_ref.json 2>> gpurun_out/bench_r02i.err
bash tools/ncu_round.sh > gpurun_out/ncu_round_r02i.log 2>&1
python -m pytest tests -m gpu -q > gpurun_out/gputest_r02i.log 2>&1
SB200_ARK=plain python -m pytest tests -m gpu -q > gpurun_out/gputest_r02i_plain.log 2>&1
tail -n 2 gpurun_out/gputest_r02i.log gpurun_out/gputest_r02i_plain.log; tail -c 300 gpurun_out/bench_r02i.json
