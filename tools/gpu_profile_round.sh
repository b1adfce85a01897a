#!/bin/bash
# Run on the GPU box (gpurun): tests, smoke, bench (both arms), ncu launch list + full captures -> gpurun_out/
R=${1:-r02}
mkdir -p gpurun_out
python -m pytest tests -m gpu -x -q 2>&1 | tail -3 | tee gpurun_out/${R}_pytest_gpu.txt
python __graft_entry__.py --smoke 2>&1 | grep -v Warning | tail -6 | tee gpurun_out/${R}_smoke.txt
python tools/accuracy_report.py --out gpurun_out/${R}_accuracy.json > gpurun_out/${R}_accuracy.txt 2>&1
python tools/accuracy_stream_scan.py > gpurun_out/${R}_accuracy_stream16.txt 2>&1
python tools/time_kd_pose_loss.py 64 > /dev/null 2>&1; cp gpurun_out/kd_pose_loss_timing.json gpurun_out/${R}_kd_pose_loss_timing.json
python bench.py --impl reference --steps 10 --warmup 3 > gpurun_out/${R}_bench_reference.json 2>/dev/null
python bench.py > gpurun_out/${R}_bench_ape_b64.json 2> gpurun_out/${R}_bench.err
python bench.py --workload dense_b32 --steps 10 --warmup 3 --no-cpu-baseline > gpurun_out/${R}_bench_dense_b32.json 2>> gpurun_out/${R}_bench.err
tail -c 600 gpurun_out/${R}_bench_ape_b64.json; echo; tail -c 400 gpurun_out/${R}_bench_dense_b32.json; echo
# launch list of the same bench command (per-launch device times are cold-cache / serialised: compare SHARES)
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/${R}_launches_ape_b64.csv \
    python bench.py --steps 5 --warmup 3 --no-cpu-baseline --no-b0 > /dev/null 2>&1
ncu --set full --clock-control none --import-source on -k regex:kdot_small_fast -s 3 -c 1 -o gpurun_out/${R}_prof_small_fast \
    python bench.py --steps 2 --warmup 3 --no-cpu-baseline --no-dense --no-b0 > /dev/null 2>&1
ncu --set full --clock-control none --import-source on -k regex:kdot_stream -c 1 -o gpurun_out/${R}_prof_stream \
    python bench.py --workload dense_b32 --steps 1 --warmup 3 --no-cpu-baseline > /dev/null 2>&1
ncu --set full --clock-control none --import-source on -k regex:kdot_stream -c 1 -o gpurun_out/${R}_prof_stream_zebra \
    python bench.py --workload zebra_b8 --steps 1 --warmup 3 --no-cpu-baseline --no-dense > /dev/null 2>&1
ncu --set full --clock-control none --import-source on -k regex:kdot_tiled -s 3 -c 1 -o gpurun_out/${R}_prof_tiled \
    python bench.py --workload multi_b64 --steps 2 --warmup 3 --no-cpu-baseline --no-dense > /dev/null 2>&1
python bench.py --workload multi_b64 --steps 50 --warmup 5 --no-cpu-baseline --no-dense > gpurun_out/${R}_bench_multi_b64.json 2>> gpurun_out/${R}_bench.err
python bench.py --workload zebra_b8 --steps 5 --warmup 3 --no-cpu-baseline --no-dense > gpurun_out/${R}_bench_zebra_b8.json 2>> gpurun_out/${R}_bench.err
ncu --metrics gpu__time_duration.sum --clock-control none -c 100 --csv --log-file gpurun_out/${R}_launches_dense_b32.csv \
    python bench.py --workload dense_b32 --steps 3 --warmup 3 --no-cpu-baseline > /dev/null 2>&1
ncu --set full --clock-control none -k regex:kdot_select -c 1 -o gpurun_out/${R}_prof_select \
    python -m pytest tests/test_postprocess_gpu.py -m gpu -q -k bit_exact > /dev/null 2>&1
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active --format=csv > gpurun_out/${R}_nvidia_smi.csv
ls -la gpurun_out | tail -12
