#!/bin/bash
# compile the H=3 instantiations of the round-2 streaming kernel, print registers/spills, dump the velocity+grad kernel's SASS
cd "$(dirname "$0")/.."
nvcc -gencode arch=compute_100a,code=sm_100a -lineinfo -O3 -std=c++17 -Xcompiler -fPIC -I include -I velocycle_b200/csrc -DVCB_INST_H=3 $EXTRA -Xptxas -v -c -o /tmp/s2_h3.o velocycle_b200/csrc/vcb_stream2_inst.cu 2>&1 | grep -A2 "ILi3ELb1ELb1" | grep -v "Compiling\|Function prop"
mkdir -p /tmp/s2 && tools/sass_loop.sh /tmp/s2_h3.o '_ZN3vcb2s218vcb_stream2_kernelILi3ELb1ELb1EEEvNS0_6ParamsE' /tmp/s2/k.sass
