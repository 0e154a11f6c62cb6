# one box: GPU tests, the default bench line, the launch list of the same bench
# command and a full ncu capture of one step's hot kernels (each ncu pass only
# after the same command has exited 0 without ncu)
set -x
python -m pytest tests -m gpu -x -q 2>&1 | tail -4 > gpurun_out/gpu_tests.txt
python bench.py > gpurun_out/bench_1gpu.json 2> gpurun_out/bench_1gpu.err || exit 1
python tools/step_timeline.py 2>&1 | grep -v Warn | tail -13 > gpurun_out/step_timeline.txt
python tools/eval_bench.py > gpurun_out/eval_bench.json 2>gpurun_out/eval_bench.err
python bench.py --steps 2 --warmup 3 --no-cpu --no-extra > gpurun_out/b_short.log 2>&1 || exit 1
ncu --metrics gpu__time_duration.sum --clock-control none -c 60 --csv --log-file gpurun_out/step_launches.csv python bench.py --steps 2 --warmup 3 --no-cpu --no-extra > gpurun_out/ncu_launches.log 2>&1
python tools/profile_step.py > gpurun_out/ps.log 2>&1 || exit 1
ncu --set full --import-source on --clock-control none -k regex:'col_kernel|cons_rows|pyramid' -s 9 -c 9 -o gpurun_out/prof_r02c -f python tools/profile_step.py > gpurun_out/ncu_full.log 2>&1
tail -2 gpurun_out/ncu_full.log
