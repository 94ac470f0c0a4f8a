set -x
timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/r02i_pytest_gpu.log 2>&1; tail -3 gpurun_out/r02i_pytest_gpu.log
timeout 500 python bench.py > gpurun_out/r02i_bench.json 2> gpurun_out/r02i_bench.err; tail -c 300 gpurun_out/r02i_bench.err
B200S_NO_GRAPH=1 timeout 500 ncu --profile-from-start off --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/r02i_launches_solve.csv python tools/prof_solve_only_ncu.py 100 3 > gpurun_out/ncu1.log 2>&1
B200S_NO_GRAPH=1 timeout 500 ncu --profile-from-start off --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/r02i_launches_solve_mode0.csv python tools/prof_solve_only_ncu.py 100 0 > gpurun_out/ncu0.log 2>&1
B200S_NO_GRAPH=1 timeout 600 ncu --profile-from-start off -k regex:persist --set full --clock-control none --import-source on -o gpurun_out/r02i_ncu_persist python tools/prof_solve_only_ncu.py 100 3 > gpurun_out/ncu2.log 2>&1; tail -2 gpurun_out/ncu2.log
ls -la gpurun_out | tail -8
