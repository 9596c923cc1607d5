"""Timing of the fused step at one size with the tcgen05 kernel.  The VCB_UMMA_DEBUG / VCB_UMMA_TRACE experiments of DESIGN.md
section 4c need a library built with -DVCB_UMMA_INSTRUMENT (both vcb.cu and vcb_umma.cu); the product build ignores them."""
import os, sys
os.environ.setdefault("VCB_STREAM_KERNEL", "umma")
sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))
from quick_bench import run
run(int(sys.argv[1]) if len(sys.argv) > 1 else 400_000, 2000, True)
