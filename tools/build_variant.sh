#!/bin/bash
# Build a kernel variant into build/variants/librl_<name>.so: tools/build_variant.sh <name> [-DMACRO=value ...]   (development aid, see tools/ab_variants.py)
set -e
cd "$(dirname "$0")/.."
name=$1; shift
mkdir -p build/variants
nvcc -gencode arch=compute_100a,code=sm_100a -std=c++17 -O3 -lineinfo -fmad=false -Xcompiler -fPIC,-ffp-contract=off -shared \
  -Iinclude -Irustlight_b200/csrc "$@" rustlight_b200/csrc/rl_b200.cu -o build/variants/librl_$name.so -ldl
cuobjdump -res-usage build/variants/librl_$name.so 2>/dev/null | grep -A1 -E "k_trace_shadow_flat|k_shadow_flatE" | grep REG | awk -v n=$name '{print n, $1, $2}'
