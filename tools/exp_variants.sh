#!/bin/bash
# build libvcb variants that differ in the H=3 round-2 stream kernel (timing experiments only) into tools/exp/
set -e
cd "$(dirname "$0")/.."
mkdir -p tools/exp
F="-gencode arch=compute_100a,code=sm_100a -lineinfo -O3 -std=c++17 -Xcompiler -fPIC -I include -I velocycle_b200/csrc"
for v in "$@"; do
  flags=""
  IFS='+' read -ra parts <<< "$v"
  for p in "${parts[@]}"; do [ "$p" != "BASE" ] && flags="$flags -DVCB_EXP_$p"; done
  nvcc $F $flags -DVCB_INST_H=3 -c -o /tmp/exp_$v.o velocycle_b200/csrc/vcb_stream2_inst.cu &
done
wait
for v in "$@"; do
  objs=$(ls build/obj/*.o | grep -v vcb_stream2_h3.o)
  nvcc -gencode arch=compute_100a,code=sm_100a -shared -o tools/exp/libvcb_$v.so $objs /tmp/exp_$v.o
done
ls -la tools/exp
