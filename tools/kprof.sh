#!/bin/bash
# kprof.sh [lib.so ...]: per library, time the streaming kernel (exp_time.py) and take one ncu --set full capture of it at
# 200k x 2k; then print the time, the per-stage opcode histogram (weighted by execution count) and the stall breakdown.
cd "$(dirname "$0")/.."
LIBS=${@:-velocycle_b200/libvcb.so}
CMD=""
for LIB in $LIBS; do
  TAG=$(basename $LIB .so)
  CMD="$CMD python tools/exp_time.py $LIB 2>&1 | grep 'stream kernel'; VCB_LIB=$LIB ncu --set full --clock-control none --import-source on -k regex:vcb_stream[23]_kernel -s 2 -c 1 -f -o gpurun_out/kprof_$TAG python tools/prof_one.py 200000 2000 > gpurun_out/kprof_$TAG.log 2>&1; tail -1 gpurun_out/kprof_$TAG.log;"
done
/usr/local/graft/bin/gpurun --timeout 1200 -- "$CMD" 2>&1 | grep -E "stream kernel|status=|rror"
for LIB in $LIBS; do
  TAG=$(basename $LIB .so)
  echo "== $TAG"
  python tools/ncu_hist.py gpurun_out/kprof_$TAG.ncu-rep
done
