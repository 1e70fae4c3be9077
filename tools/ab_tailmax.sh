#!/bin/bash
# k_tail hand-over threshold (RL_TAIL_MAX, read per render) on one rank's eighth of config 2 and on the full frame: tools/ab_tailmax.sh "<values>" [lib]   (development aid)
for t in $1; do echo -n "RL_TAIL_MAX=$t "; RL_TAIL_MAX=$t RL_B200_LIB=${2:-rustlight_b200/librl_b200.so} python tools/ab_rank.py 8 128 2>&1 | tail -1; done
