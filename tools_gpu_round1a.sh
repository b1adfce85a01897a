#!/bin/bash
# first measurement pass: tests, smoke, bench, ncu launch list + full captures
mkdir -p gpurun_out
python -m pytest tests -m gpu -x -q 2>&1 | tail -5
python __graft_entry__.py --smoke 2>&1 | tail -8
python bench.py --steps 200 --warmup 20 > gpurun_out/bench_ape.json 2> gpurun_out/bench_ape.err; tail -c 3000 gpurun_out/bench_ape.json; tail -5 gpurun_out/bench_ape.err
python bench.py --workload dense_b32 --steps 5 --warmup 3 --no-cpu-baseline > gpurun_out/bench_dense.json 2> gpurun_out/bench_dense.err; tail -c 2500 gpurun_out/bench_dense.json; tail -5 gpurun_out/bench_dense.err
python bench.py --impl reference --steps 5 --warmup 3 > gpurun_out/bench_ref.json 2>&1; tail -c 1500 gpurun_out/bench_ref.json
ncu --metrics gpu__time_duration.sum --clock-control none -c 60 --csv --log-file gpurun_out/launches_ape.csv python bench.py --steps 5 --warmup 3 --no-cpu-baseline --no-dense > gpurun_out/ncu_ape.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:kdot_tiled -c 1 -o gpurun_out/prof_tiled python bench.py --workload dense_b32 --steps 1 --warmup 3 --no-cpu-baseline > gpurun_out/ncu_tiled.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:kdot_small -s 3 -c 1 -o gpurun_out/prof_small python bench.py --steps 2 --warmup 3 --no-cpu-baseline --no-dense > gpurun_out/ncu_small.log 2>&1
ls -la gpurun_out
