#!/bin/bash
# Round-end snapshots on one B200 (run under gpurun from the repo root): GPU test log, N=1 bench line, reference arm,
# ncu launch lists of one 100^3 solve (persistent sweeps / launch per step) and of the KLU bench command.
TAG=${1:-r02j}
set -x
timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/${TAG}_pytest_gpu.log 2>&1; tail -3 gpurun_out/${TAG}_pytest_gpu.log
timeout 600 python bench.py > gpurun_out/${TAG}_bench.json 2> gpurun_out/${TAG}_bench.err; tail -c 300 gpurun_out/${TAG}_bench.err
timeout 600 python bench.py --impl reference > gpurun_out/${TAG}_bench_reference.json 2> gpurun_out/${TAG}_bench_reference.err
B200S_NO_GRAPH=1 timeout 500 ncu --profile-from-start off --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/${TAG}_launches_solve.csv python tools/prof_solve_only_ncu.py 100 3 > gpurun_out/ncu1.log 2>&1
B200S_PERSIST_DBG=1 B200S_NO_GRAPH=1 timeout 300 python tools/prof_solve_persist_ncu.py 100 > gpurun_out/${TAG}_persist_phase_clocks.txt 2>&1
ls -la gpurun_out | tail -8
