#!/bin/bash
# K1 experiments (needs the -DLEMON_TC_EXPERIMENT build): times one K1 launch shape under the tuning knobs.
#   tools/k1_variants.sh NQ M D  ->  one line per configuration
export LEMON_B200_LIB=lemon_b200/build_exp/liblemon_b200_exp.so
NQ=$1; M=$2; D=$3
for cfg in "0 0" "4 0" "8 0" "0 1" "0 2" "0 3" "8 1" "8 3" "4 1"; do
  set -- $cfg
  echo -n "stagger=$1 variant=$2 : "
  LEMON_TC_STAGGER=$1 LEMON_TC_VARIANT=$2 python tools/tc_debug.py 2 $NQ $M $D 1 --time 2>&1 | grep -E "^time" || echo FAILED
done
