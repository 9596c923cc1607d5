#!/bin/bash
# Rebuild only the objects named on the command line (default: vcb vcb_umma) and relink libvcb.so.
set -e
cd "$(dirname "$0")/.."
F="-gencode arch=compute_100a,code=sm_100a -lineinfo -O3 -std=c++17 -Xcompiler -fPIC -I include -I velocycle_b200/csrc"
objs=${@:-vcb vcb_umma}
for o in $objs; do
  case $o in
    vcb_stream2_h*) h=${o#vcb_stream2_h}; nvcc $F -DVCB_INST_H=$h -c -o build/obj/$o.o velocycle_b200/csrc/vcb_stream2_inst.cu & ;;
    vcb_stream_h*) h=${o#vcb_stream_h}; nvcc $F -DVCB_INST_H=$h -c -o build/obj/$o.o velocycle_b200/csrc/vcb_stream_inst.cu & ;;
    *) nvcc $F -c -o build/obj/$o.o velocycle_b200/csrc/$o.cu & ;;
  esac
done
wait
nvcc -gencode arch=compute_100a,code=sm_100a -shared -o velocycle_b200/libvcb.so build/obj/*.o
touch velocycle_b200/libvcb.so
echo "linked"
