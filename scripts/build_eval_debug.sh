#!/bin/bash
# Builds diagnostic variants of the library into chaorec_b200/dbg/ (run here, before gpurun: the .so files are
# git-ignored but travel with the snapshot).  Arguments: name:flags ..., e.g.  floor3:-DCHAOREC_EV_DEBUG=3
cd "$(dirname "$0")/.."
mkdir -p chaorec_b200/dbg
for spec in "$@"; do
  name=${spec%%:*}; flags=${spec#*:}
  nvcc -gencode arch=compute_100a,code=sm_100a -lineinfo -O3 -std=c++17 -Xcompiler -fPIC -shared ${flags//,/ } \
       -o chaorec_b200/dbg/libchaorec_b200_$name.so chaorec_b200/csrc/*.cu 2>&1 | grep -E "error" &
done
wait
ls -la chaorec_b200/dbg
