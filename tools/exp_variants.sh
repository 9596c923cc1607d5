#!/bin/bash
# exp_variants.sh <2|3> VARIANT...: build libvcb variants that differ in the H=3 streaming kernel of generation 2
# (vcb_stream2.cuh) or 3 (vcb_stream3.cuh) into tools/exp/ (timing experiments only).  VARIANT = flags joined by '+'
# (each becomes -DVCB_EXP_<flag>); BASE = no flag.
set -e
cd "$(dirname "$0")/.."
GEN=$1; shift
mkdir -p tools/exp
F="-gencode arch=compute_100a,code=sm_100a -lineinfo -O3 -std=c++17 -Xcompiler -fPIC -I include -I velocycle_b200/csrc"
nvcc $F -DVCB_STREAM_GEN=$GEN -c -o /tmp/exp_vcb_g$GEN.o velocycle_b200/csrc/vcb.cu &
for v in "$@"; do
  flags=""
  IFS='+' read -ra parts <<< "$v"
  for p in "${parts[@]}"; do [ "$p" != "BASE" ] && flags="$flags -DVCB_EXP_$p"; done
  nvcc $F $flags -DVCB_INST_H=3 -c -o /tmp/exp_s${GEN}_$v.o velocycle_b200/csrc/vcb_stream${GEN}_inst.cu &
done
wait
for v in "$@"; do
  objs=$(ls build/obj/*.o | grep -v "vcb_stream${GEN}_h3.o\|/vcb.o")
  nvcc -gencode arch=compute_100a,code=sm_100a -shared -o tools/exp/libvcb_s${GEN}_$v.so $objs /tmp/exp_vcb_g$GEN.o /tmp/exp_s${GEN}_$v.o
done
ls tools/exp
