"""Run a few fused steps of one configuration (for ncu)."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from velocycle_b200 import _lib
if os.environ.get("VCB_LIB"): _lib.LIB_PATH = os.path.abspath(os.environ["VCB_LIB"])
from velocycle_b200.fused import PackedCounts, fused_elbo_grad
from velocycle_b200.synthetic import make_synthetic
Nc = int(sys.argv[1]) if len(sys.argv) > 1 else 100_000
Ng = int(sys.argv[2]) if len(sys.argv) > 2 else 2000
velocity = (sys.argv[3] != "phase") if len(sys.argv) > 3 else True
d = make_synthetic(Nc, Ng, H=3, Hw=1, seed=0, device="cuda", stats=False)
counts = PackedCounts(d.S, d.U if velocity else None, d.Ng, d.batch_id, d.cond_id)
gamma = torch.exp(d.loggamma)
args = (counts, d.phi, d.cf, d.nu, d.dnu, d.shape_inv) + ((d.logbeta, gamma, d.nu_omega) if velocity else ())
for _ in range(4):
    fused_elbo_grad(*args, grad=True)
torch.cuda.synchronize()
