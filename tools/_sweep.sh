python tools/tc_debug.py 2 4000 20000 512 1 2>&1 | tail -6
run() { echo "== $*"; env "$@" python tools/tc_debug.py 2 118000 118000 512 1 --time 2>&1 | tail -1; }
run A=1
run LEMON_TC_DEBUG=2
run LEMON_TC_PPT=6
python tools/tc_debug.py 2 50000 400000 768 1 --time 2>&1 | tail -1
