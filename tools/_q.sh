python tools/accuracy_report.py 2>&1 | cut -c1-60 | tail -4
python tools/_acc2.py
for w in dense_b32 zebra_b8; do python bench.py --steps 10 --warmup 3 --workload $w --no-dense --no-cpu-baseline 2>/dev/null | tail -1 | python -c "import json,sys; d=json.loads(sys.stdin.read()); print('$w', round(d['ms_per_step'],4),'ms', round(d['value'],1),'img/s', round(d['roofline']['frac'],4))"; done
python -m pytest tests -m gpu -q 2>&1 | tail -12
