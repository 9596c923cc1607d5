"""Cross-check library variants against each other: exp_check.py lib1.so lib2.so ... runs the same three problems (velocity,
velocity with 5 shuffled batches, phase) through every library in its own process and prints the largest normwise difference
of every output against the first library."""
import os, subprocess, sys, tempfile
if len(sys.argv) > 2 and sys.argv[1] != "--one":
    outs = []
    for l in sys.argv[1:]:
        f = tempfile.mktemp(suffix=".pt")
        subprocess.run([sys.executable, __file__, "--one", l, f], check=True)
        outs.append(f)
    import torch
    ref = torch.load(outs[0])
    for l, f in zip(sys.argv[2:], outs[1:]):
        o = torch.load(f)
        worst = max(((o[k].double() - ref[k].double()).norm() / (ref[k].double().norm() + 1e-300)).item() for k in ref)
        wk = max(ref, key=lambda k: ((o[k].double() - ref[k].double()).norm() / (ref[k].double().norm() + 1e-300)).item())
        print(f"{os.path.basename(l):32s} vs {os.path.basename(sys.argv[1])}: worst normwise diff {worst:.2e} ({wk})", flush=True)
    sys.exit(0)
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from velocycle_b200 import _lib
_lib.LIB_PATH = os.path.abspath(sys.argv[2])
from velocycle_b200.fused import PackedCounts, fused_elbo_grad
from velocycle_b200.synthetic import make_synthetic
res = {}
for tag, Nc, Ng, Nb, velo in (("v", 30001, 2000, 1, True), ("vb", 20011, 1000, 5, True), ("p", 30001, 2000, 1, False)):
    d = make_synthetic(Nc, Ng, H=3, Hw=1, Nb=Nb, seed=1, device="cuda", stats=False)
    bid = d.batch_id
    if Nb > 1:
        bid = bid[torch.randperm(Nc, device="cuda", generator=torch.Generator("cuda").manual_seed(0))]
    counts = PackedCounts(d.S, d.U if velo else None, d.Ng, bid, d.cond_id)
    args = (counts, d.phi, d.cf, d.nu * 1.1, d.dnu, d.shape_inv) + ((d.logbeta, torch.exp(d.loggamma) + 0.3, d.nu_omega) if velo else ())
    out = fused_elbo_grad(*args, grad=True)
    torch.cuda.synchronize()
    for k, v in out.items():
        if torch.is_tensor(v) and not k.startswith("_"): res[f"{tag}.{k}"] = v.detach().cpu()  # ("_workspace" is scratch)
torch.save(res, sys.argv[3])
