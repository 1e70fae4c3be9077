#!/bin/bash
# Tessellated Cornell box (tree kernels) through each variant in build/variants: tools/ab_tess.sh "<variants>" "<N list>" [spp]   (development aid)
for v in $1; do for n in $2; do echo -n "$v "; RL_B200_LIB=build/variants/librl_$v.so timeout 300 python tools/tess_cbox.py $n ${3:-16} 2>&1 | grep "tess$n"; done; done
