#!/bin/bash
# Developer helper: build an A/B variant of the library next to the product one.
#   scripts/build_variant.sh NAME [-DFLAG ...]   ->  amuse_b200/lib/libamuse_b200_NAME.so
# Use it with  AMUSE_B200_LIB=amuse_b200/lib/libamuse_b200_NAME.so python ...
set -e
cd "$(dirname "$0")/.."
name=$1; shift
srcs=$(python -c "import __graft_entry__ as g; print(' '.join(str(g.CSRC / s) for s in g.SOURCES))")
/usr/local/cuda/bin/nvcc -gencode arch=compute_100a,code=sm_100a -lineinfo -O3 -std=c++17 -Xcompiler -fPIC -shared "$@" \
  -o amuse_b200/lib/libamuse_b200_$name.so $srcs 2>&1 | grep -E "error|warning: v" || true
ls -la amuse_b200/lib/libamuse_b200_$name.so
