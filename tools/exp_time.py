"""Time the streaming kernel of library variants (tools/exp_variants.sh): exp_time.py lib1.so lib2.so ...  (one process per lib)."""
import os, subprocess, sys
if len(sys.argv) > 2 or (len(sys.argv) == 2 and not sys.argv[1].endswith(".so")):
    libs = sys.argv[1:]
    for l in libs:
        subprocess.run([sys.executable, __file__, l])
    sys.exit(0)
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from velocycle_b200 import _lib
_lib.LIB_PATH = os.path.abspath(sys.argv[1])
from velocycle_b200.fused import PackedCounts, fused_elbo_grad
from velocycle_b200.synthetic import make_synthetic
import bench
Nc, Ng = 400_000, 2000
d = make_synthetic(Nc, Ng, H=3, Hw=1, seed=0, device="cuda", stats=False)
counts = PackedCounts(d.S, d.U, d.Ng, d.batch_id, d.cond_id)
ev = bench.CudaEvents(); counts.profile_events = ev.handles()
gamma = torch.exp(d.loggamma)
args = (counts, d.phi, d.cf, d.nu, d.dnu, d.shape_inv, d.logbeta, gamma, d.nu_omega)
def t(legacy):
    ts = []
    for i in range(10):
        fused_elbo_grad(*args, grad=True, legacy_stream=legacy)
        torch.cuda.synchronize()
        if i >= 3: ts.append(ev.elapsed_ms())
    return sum(ts) / len(ts)
new, ref = t(False), t(True)
print(f"{os.path.basename(sys.argv[1]):32s} stream kernel {new:.3f} ms at {Nc}x{Ng}  -> {new*2.5:.2f} ms per 1M cells | v16 in the same process {ref*2.5:.2f} ms | ratio {new/ref:.3f}", flush=True)
