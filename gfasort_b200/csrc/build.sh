#!/usr/bin/env bash
# Builds gfasort_b200/libgfasort_cuda.so for B200 (sm_100a).  Usage: build.sh [extra nvcc flags]
set -euo pipefail
here="$(cd "$(dirname "${BASH_SOURCE[0]}")" && pwd)"
out="$here/../libgfasort_cuda.so"
nvcc -gencode arch=compute_100a,code=sm_100a -O3 -lineinfo -std=c++17 \
     -Xcompiler -fPIC,-ffp-contract=off,-pthread -shared \
     -o "$out" "$here/gfs_lib.cu" "$here/gfs_p2p.cu" "$here/gfs_multi.cu" "$here/gfs_synth.cpp" "$here/gfs_host_graph.cpp" "$here/gfs_io.cpp" "$@"
echo "built $out"
