for lib in libkdot.so libkdot_mb2.so; do for n in 64 256 1024; do KDOT_LIB=$PWD/kd_6d_pose_adlp_b200/lib/$lib python bench.py --steps 30 --warmup 5 --images $n 2>/dev/null | tail -1 | python -c "import json,sys; d=json.loads(sys.stdin.read()); print('$lib', $n, round(d['ms_per_step']*1000,2),'us', round(d['value']/1e6,3),'M img/s', 'e2e', round(d['e2e']['value']/1e6,3))"; done; done
python tools/accuracy_report.py 2>&1
for w in dense_b32 zebra_b8 multi_b64; do python bench.py --steps 5 --warmup 3 --workload $w 2>/dev/null | tail -1 | python -c "import json,sys; d=json.loads(sys.stdin.read()); print('$w', round(d['ms_per_step'],3),'ms', round(d['value'],1),'img/s', d['roofline']['frac'])"; done
python -m pytest tests -m gpu -x -q 2>&1 | tail -5
