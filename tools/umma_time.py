"""Timing of the fused step at one size with the tcgen05 kernel (VCB_UMMA_DEBUG experiments from the environment)."""
import os, sys
os.environ.setdefault("VCB_STREAM_KERNEL", "umma")
sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))
from quick_bench import run
run(int(sys.argv[1]) if len(sys.argv) > 1 else 400_000, 2000, True)
