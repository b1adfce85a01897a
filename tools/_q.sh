python tools/accuracy_report.py 2>&1 | cut -c1-60 | head -4
for n in 64 256 1024; do python bench.py --steps 30 --warmup 5 --images $n --no-dense --no-cpu-baseline 2>/dev/null | tail -1 | python -c "import json,sys; d=json.loads(sys.stdin.read()); print($n, round(d['ms_per_step']*1000,2),'us', round(d['value']/1e6,3),'M img/s', 'e2e', round(d['e2e']['value']/1e6,3))"; done
python -m pytest tests -m gpu -x -q 2>&1 | tail -3
