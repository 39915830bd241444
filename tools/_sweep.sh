python -m pytest tests -m gpu -x -q 2>&1 | tail -2
python bench.py --no-cpu-baseline | python -c "import sys,json; j=json.loads(sys.stdin.read()); print(j['ms_per_step'], j['e2e']['ms_per_step'], j['roofline']['achieved'], j['config']['uncertified_rows_per_step'], j['clocks'])"
python tools/tc_debug.py 2 2168 118000 512 8 --time 2>&1 | tail -1
