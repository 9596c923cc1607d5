"""Scratch: per-tensor normwise error of the fused CUDA path against the fp64 oracle for a few shapes."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), 'tests'))
from test_kernel_parity import _run, SHAPES
from velocycle_b200.synthetic import make_synthetic

shapes = SHAPES + [(4096, 512, 3, 1, 16, 2), (20000, 2000, 3, 1, 1, 1)]
for shape in shapes:
    for velocity in (False, True):
        for sorted_b in ((True, False) if shape[4] > 1 else (True,)):
            Nc, Ng, H, Hw, Nb, Nx = shape
            d = make_synthetic(Nc, Ng, H=H, Hw=Hw, Nb=Nb, Nx=Nx, seed=3, device="cuda", sorted_batches=sorted_b)
            try:
                out, ref, _ = _run(d, velocity)
            except Exception as e:  # noqa: BLE001
                print(shape, velocity, sorted_b, "EXC", repr(e)[:200], flush=True)
                continue
            ref32 = ref.pop("_ref32")
            errs = []
            for k, v in ref.items():
                if k in ("total", "omega") or k not in out:
                    continue
                got = out[k].double().cpu().reshape(v.shape)
                err = float((got - v).abs().max() / (v.abs().max() + 1e-30))
                e32 = float((ref32[k].double().reshape(v.shape) - v).abs().max() / (v.abs().max() + 1e-30))
                flag = "" if err <= max(1e-4, 2 * e32) else " <<<<<"
                errs.append(f"{k}={err:.1e}({e32:.0e}){flag}")
            print(shape, "velo" if velocity else "phase", "sorted" if sorted_b else "unsorted", " ".join(errs), flush=True)
