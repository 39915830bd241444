run() { echo "== $*"; env "$@" python tools/tc_debug.py 2 118000 118000 512 1 --time 2>&1 | tail -1; }
run A=1
run LEMON_TC_DEBUG=2
run LEMON_TC_SOFT=128
run LEMON_TC_SOFT=208
run LEMON_TC_PPT=6
run LEMON_TC_PPT=1
run LEMON_TC_BOOT=16
run LEMON_TC_BOOT=0
run LEMON_TC_SOFT=144 LEMON_TC_PPT=6
