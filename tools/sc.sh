#!/bin/bash
# sc.sh [extra nvcc flags...]: compile the H=3 round-2 streaming kernel to /tmp and print the static hot-region histogram
cd "$(dirname "$0")/.."
F="-gencode arch=compute_100a,code=sm_100a -lineinfo -O3 -std=c++17 -Xcompiler -fPIC -I include -I velocycle_b200/csrc"
nvcc $F "$@" -DVCB_INST_H=3 -Xptxas -v -c -o /tmp/s/sc.o velocycle_b200/csrc/vcb_stream2_inst.cu 2>&1 | grep -A1 "ILi3ELb1ELb1E" | grep -E "registers|spill" | head -2
python tools/static_count.py /tmp/s/sc.o
