"""Scratch timing of the fused step (CUDA events); bench.py is the contract version."""
import sys, os, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from velocycle_b200.fused import PackedCounts, fused_elbo_grad
from velocycle_b200.synthetic import make_synthetic

def run(Nc, Ng, velocity=True, H=3, iters=5, inline=False, Nb=1, Nx=1, legacy=False):
    d = make_synthetic(Nc, Ng, H=H, Hw=1, Nb=Nb, Nx=Nx, seed=0, device="cuda", stats=True)
    t0 = time.time()
    counts = PackedCounts(d.S, d.U if velocity else None, d.Ng, d.batch_id, d.cond_id, spectrum=not inline)
    torch.cuda.synchronize(); t_spec = time.time() - t0
    gamma = torch.exp(d.loggamma)
    args = (counts, d.phi, d.cf, d.nu, d.dnu, d.shape_inv) + ((d.logbeta, gamma, d.nu_omega) if velocity else ())
    for _ in range(3):
        fused_elbo_grad(*args, grad=True, inline_lgamma=inline, legacy_stream=legacy)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(iters):
        fused_elbo_grad(*args, grad=True, inline_lgamma=inline, legacy_stream=legacy)
    e1.record(); torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / iters
    nb = (2 if velocity else 1) * 4 * Nc * Ng
    print(f"Nc={Nc} Ng={Ng} velo={velocity} H={H} Nb={Nb} legacy={legacy}: {ms:.3f} ms/step  {nb/ms/1e6:.1f} GB/s "
          f"({nb/ms/1e6/6518.6*100:.1f}% of measured HBM)  zeros S/U={d.zero_frac_S:.2f}/{d.zero_frac_U:.2f} spectrum {t_spec*1e3:.0f} ms "
          f"maxk={counts.spec_S.max_count if counts.spec_S else -1}", flush=True)

def kernels(Nc, Ng, velocity=True, H=3):
    """Per-kernel device times of one fused step (torch.profiler / CUPTI)."""
    from torch.profiler import profile, ProfilerActivity
    d = make_synthetic(Nc, Ng, H=H, Hw=1, seed=0, device="cuda", stats=False)
    counts = PackedCounts(d.S, d.U if velocity else None, d.Ng, d.batch_id, d.cond_id)
    gamma = torch.exp(d.loggamma)
    args = (counts, d.phi, d.cf, d.nu, d.dnu, d.shape_inv) + ((d.logbeta, gamma, d.nu_omega) if velocity else ())
    for _ in range(2):
        fused_elbo_grad(*args, grad=True)
    torch.cuda.synchronize()
    with profile(activities=[ProfilerActivity.CUDA]) as prof:
        for _ in range(3):
            fused_elbo_grad(*args, grad=True)
        torch.cuda.synchronize()
    for e in sorted(prof.key_averages(), key=lambda e: -e.device_time_total)[:8]:
        print(f"   {e.key[:70]:70s} n={e.count:3d} avg={e.device_time_total / e.count / 1e3:8.3f} ms", flush=True)


if __name__ == "__main__":
    if len(sys.argv) > 1 and sys.argv[1] == "kernels":
        kernels(int(sys.argv[2]), int(sys.argv[3]))
        sys.exit(0)
    for legacy in (False, True):
        run(400_000, 2000, True, legacy=legacy)
        run(400_000, 2000, False, legacy=legacy)
        run(100_000, 5000, True, Nb=16, Nx=2, legacy=legacy)
        run(100_000, 2000, True, legacy=legacy)
        run(3000, 218, True, H=1, legacy=legacy)
