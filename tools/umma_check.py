"""Parity report of the tcgen05 streaming kernel (VCB_FLAG_TCGEN05) against the fp64 oracle."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests"))
import torch
from velocycle_b200.synthetic import make_synthetic
import test_kernel_parity as T

shapes = [(11, 7, 2, 1, 1, 2), (1849, 76, 1, 0, 1, 1), (257, 203, 3, 1, 1, 2), (600, 1918, 1, 1, 1, 1), (5000, 2000, 3, 1, 1, 1),
          (20003, 2000, 3, 1, 1, 1)]
if len(sys.argv) > 1 and sys.argv[1] == "small":
    shapes = shapes[:3]
worst = 0.0
for (Nc, Ng, H, Hw, Nb, Nx) in shapes:
    d = make_synthetic(Nc, Ng, H=H, Hw=Hw, Nb=Nb, Nx=Nx, seed=3, device="cuda", sorted_batches=True)
    out, ref, _ = T._run(d, True, tcgen05=True)
    ref32 = ref.pop("_ref32")
    line = []
    for k, v in ref.items():
        if k in ("total", "omega") or k not in out:
            continue
        got = out[k].double().cpu().reshape(v.shape)
        err = float((got - v).abs().max() / (v.abs().max() + 1e-30)) if torch.isfinite(got).all() else float("nan")
        e32 = float((ref32[k].double().reshape(v.shape) - v).abs().max() / (v.abs().max() + 1e-30)) if k in ref32 else 0.0
        line.append(f"{k}={err:.1e}({e32:.0e})")
        if not (err <= max(1e-4, 2 * e32)):
            worst = max(worst, err if err == err else 1e9)
    print(f"Nc={Nc} Ng={Ng} H={H}: " + " ".join(line), flush=True)
print("WORST VIOLATION", worst, flush=True)
if len(sys.argv) > 1 and sys.argv[1] == "small":
    sys.exit(0)
import quick_bench  # noqa
sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))
from quick_bench import run
run(100_000, 2000, True)
run(1_000_000, 2000, True)
