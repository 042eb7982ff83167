#!/usr/bin/env bash
# Builds tests/cpp/test_reference_api: the reference's own tests restated against the C++ host layer
# (include/gfasort.hpp) over gfasort_b200/libgfasort_cuda.so.
set -euo pipefail
here="$(cd "$(dirname "${BASH_SOURCE[0]}")" && pwd)"
root="$here/../.."
${CXX_HOST:-g++} -std=c++17 -O1 -Wall -Wextra -I"$root/include" "$here/test_reference_api.cpp" \
    -L"$root/gfasort_b200" -l:libgfasort_cuda.so -Wl,-rpath,'$ORIGIN/../../gfasort_b200' \
    -o "$here/test_reference_api"
echo "built $here/test_reference_api"
