python tools/accuracy_report.py 2>&1 | cut -c1-60 | sed -n 5,7p
for w in multi_b64; do python bench.py --steps 20 --warmup 3 --workload $w --no-dense --no-cpu-baseline 2>/dev/null | tail -1 | python -c "import json,sys; d=json.loads(sys.stdin.read()); print('$w', round(d['ms_per_step'],4),'ms', round(d['value'],1),'img/s', round(d['roofline']['frac'],4))"; done
python -m pytest tests -m gpu -x -q 2>&1 | tail -3
