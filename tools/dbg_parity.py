"""Debug: per-output errors of the fused kernel against the fp64 oracle on one small shape."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from velocycle_b200 import _lib
if os.environ.get("VCB_LIB"): _lib.LIB_PATH = os.path.abspath(os.environ["VCB_LIB"])
from velocycle_b200.synthetic import make_synthetic
sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests"))
import test_kernel_parity as T
Nc, Ng, H, Hw, Nb, Nx = [int(x) for x in sys.argv[1:7]] if len(sys.argv) > 6 else (257, 203, 3, 1, 2, 2)
velocity = (sys.argv[7] != "phase") if len(sys.argv) > 7 else True
d = make_synthetic(Nc, Ng, H=H, Hw=Hw, Nb=Nb, Nx=Nx, seed=3, device="cuda", sorted_batches=True)
out, ref, p = T._run(d, velocity)
for k, v in ref.items():
    if k in ("total", "omega", "_ref32") or k not in out: continue
    got = out[k].double().cpu().reshape(v.shape)
    err = (got - v).abs()
    print(f"{k:12s} max|ref| {float(v.abs().max()):.4e}  max err {float(err.max()):.3e}  rel {float(err.max() / (v.abs().max() + 1e-30)):.3e}  argmax {int(err.reshape(-1).argmax())}  got {float(got.reshape(-1)[err.reshape(-1).argmax()]):.6e} ref {float(v.reshape(-1)[err.reshape(-1).argmax()]):.6e}")
if "lp_S" in out:
    e = (out["lp_S"].double().cpu() - ref["lp_S"]).reshape(-1)
    print("lp_S err first 16 genes:", [f"{float(x):.2e}" for x in e[:16]])
    print("lp_S ref first 16 genes:", [f"{float(x):.2e}" for x in ref["lp_S"].reshape(-1)[:16]])
