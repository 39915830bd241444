python tools/tc_debug.py 2 3000 70000 768 3 2>&1 | tail -6
python tools/tc_debug.py 2 1000 9000 640 1 2>&1 | tail -6
run() { echo "== $*"; env "$@" python tools/tc_debug.py 2 50000 400000 768 1 --time 2>&1 | tail -1; }
run A=1
run LEMON_TC_KRES=12
run LEMON_TC_KRES=10
run LEMON_TC_KRES=8
run LEMON_TC_KRES=6
python tools/tc_debug.py 2 118000 118000 512 1 --time 2>&1 | tail -1
python -m pytest tests/test_gpu_parity.py -x -q -k "tc or 768 or fuzz" 2>&1 | tail -3
