#!/usr/bin/env python
"""bench.py -- the headline benchmark of the hot path (contract: see the task statement / DESIGN.md section 6).

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference]
    python -m torch.distributed.run --nnodes=1 --nproc-per-node N ... bench.py --gpus N ...

A "step" is one full SVI step of the velocity model (guide draw -> fused ELBO + every gradient over the
spliced and unspliced count matrices -> gene-gradient all-reduce when N > 1 -> ClippedAdam) through the public
API ``GraphedSVI(model, guide, optim_args, mp).step()`` (the drop-in model/guide functions traced once into a
CUDA graph; ``--eager`` times ``ppl.infer.SVI.step`` instead) on synthetic counts, loss read back every step.  Workload per GPU (weak scaling):
1,000,000 cells x 2,000 genes, 3 gene harmonics, 1 angular-speed harmonic -- the single-GPU target shape of
BASELINE.json's north_star; 16 GB of fp32 counts per GPU, far beyond the 126 MB L2, so no flush is needed.

``value``  : cell.gene negative-binomial evaluations per second (cells x genes x steps/s; each counts S and U),
             counts resident in HBM (as the reference keeps them: preprocessing.py:193-194).
``e2e``    : the same step when the count shard starts each step in pinned HOST memory and is copied to the
             device inside the timed region, plus the device->host read of the loss.
``roofline``: the streaming kernel's algorithmic bytes (8 B per cell.gene: one fp32 S and U each) over its
             CUDA-event duration, against the measured HBM peak of MEASURED_PEAKS.json.
``cpu_baseline``: the reference op chain (oracle/models.py: unfused (Ng,Nc) einsum + GammaPoisson + autograd
             under the same SVI) on the host cores, on a bounded sample of the same workload.
"""
from __future__ import annotations

import argparse
import ctypes
import json
import os
import subprocess
import sys
import tempfile
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

import torch  # noqa: E402

METRIC = "cell_gene_nb_evals_per_sec"
UNIT = "cell*gene/s"


def parse():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--cells-per-gpu", type=int, default=1_000_000)
    ap.add_argument("--genes", type=int, default=2000)
    ap.add_argument("--harmonics", type=int, default=3)
    ap.add_argument("--omega-harmonics", type=int, default=1)
    ap.add_argument("--batches", type=int, default=1)
    ap.add_argument("--conditions", type=int, default=1)
    ap.add_argument("--model-type", default="lrmn", choices=["lrmn", "normal"])
    ap.add_argument("--cpu-sample-cells", type=int, default=4000)
    ap.add_argument("--eager", action="store_true", help="time the eager ppl.infer.SVI.step instead of GraphedSVI")
    ap.add_argument("--no-e2e", action="store_true")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    return ap.parse_args()


def workload_name(a):
    return (f"synthetic velocity SVI step, {a.cells_per_gpu} cells x {a.genes} genes per GPU, H={a.harmonics}, "
            f"Hw={a.omega_harmonics}, Nb={a.batches}, Nx={a.conditions}, guide={a.model_type}")


def priors_from(d, device):
    """Cycle / phase / speed priors a user would pass: noisy versions of the generating values."""
    g = torch.Generator(device=device).manual_seed(12345)
    rn = lambda *s: torch.randn(*s, generator=g, device=device)
    mu_nu = d.nu + 0.3 * rn(*d.nu.shape)
    sd_nu = torch.full_like(d.nu, 0.5)
    phixy = torch.stack([torch.cos(d.phi), torch.sin(d.phi)], -1) + 0.3 * rn(d.Nc, 2)
    mu_nw = d.nu_omega.clone()
    sd_nw = torch.full_like(d.nu_omega, 0.05)
    sd_nw[:, 0] = 0.1
    return mu_nu, sd_nu, phixy, mu_nw, sd_nw


# ------------------------------------------------------------------------------------------------------
# CPU arm: the reference op chain on the host cores
# ------------------------------------------------------------------------------------------------------
def cpu_reference_steps(a, n_cells, steps, warmup):
    """SVI steps of the unfused reference chain on the CPU over ``n_cells`` cells of the workload."""
    from oracle import models as omodels
    from velocycle_b200 import ppl as pyro
    from velocycle_b200.ppl.infer import SVI, Trace_ELBO
    from velocycle_b200.ppl.optim import ClippedAdam
    from velocycle_b200.preprocessing import make_velocity_metaparams
    from velocycle_b200.synthetic import make_synthetic

    cores = os.cpu_count() or 1
    torch.set_num_threads(cores)
    d = make_synthetic(n_cells, a.genes, H=a.harmonics, Hw=a.omega_harmonics, Nb=a.batches, Nx=a.conditions,
                       seed=0, device="cpu", stats=False)
    mu_nu, sd_nu, phixy, mu_nw, sd_nw = priors_from(d, "cpu")
    mp = make_velocity_metaparams(d.S[:, : a.genes], d.U[:, : a.genes], mu_nu, sd_nu, phixy, mu_nw, sd_nw,
                                  batch_id=d.batch_id, cond_id=d.cond_id, Nb=a.batches, Nx=a.conditions,
                                  count_factor=d.cf, model_type=a.model_type, device="cpu")
    model = omodels.velocity_model_unfused_lrmn if a.model_type == "lrmn" else omodels.velocity_model_unfused
    pyro.clear_param_store()
    pyro.set_rng_seed(0)
    svi = SVI(model, mp.guide_fn, ClippedAdam({"lr": 0.03, "lrd": 0.9996, "betas": (0.8, 0.99)}), Trace_ELBO())
    for _ in range(warmup):
        svi.step(mp)
    t0 = time.perf_counter()
    for _ in range(steps):
        svi.step(mp)
    dt = (time.perf_counter() - t0) / max(steps, 1)
    return dt, cores


def run_reference_arm(a):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    n = a.cpu_sample_cells
    dt, cores = cpu_reference_steps(a, n, a.steps, a.warmup)
    value = n * a.genes / dt
    sample = (f"{n} cells x {a.genes} genes of the workload per step (same generator, seed 0), unfused reference op "
              f"chain + autograd under SVI on the host CPU, {cores} threads")
    line = {
        "impl": "reference", "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": a.gpus, "steps": a.steps,
        "warmup": a.warmup, "ms_per_step": dt * 1e3, "higher_is_better": True, "scaling": "weak",
        "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "config": {"workload": workload_name(a), "sample": sample},
        "svi_steps_per_sec_on_sample": 1.0 / dt,
        "cpu_baseline": {"value": value, "unit": UNIT, "cores": cores, "kind": "port", "sample": sample},
        "e2e": {"value": value, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
    }
    print(json.dumps(line), flush=True)


# ------------------------------------------------------------------------------------------------------
# helpers for the GPU arm
# ------------------------------------------------------------------------------------------------------
class CudaEvents:
    """Raw cudaEvent_t pairs (the C ABI records them around the streaming kernel)."""

    def __init__(self):
        self.rt = None
        for name in ("libcudart.so.12", "libcudart.so"):
            try:
                self.rt = ctypes.CDLL(name)
                break
            except OSError:
                continue
        if self.rt is None:
            import glob

            cands = glob.glob(os.path.join(os.path.dirname(torch.__file__), "..", "nvidia", "cuda_runtime", "lib",
                                           "libcudart.so*"))
            self.rt = ctypes.CDLL(cands[0])
        self.begin, self.end = ctypes.c_void_p(), ctypes.c_void_p()
        assert self.rt.cudaEventCreate(ctypes.byref(self.begin)) == 0
        assert self.rt.cudaEventCreate(ctypes.byref(self.end)) == 0

    def handles(self):
        return self.begin.value, self.end.value

    def elapsed_ms(self) -> float:
        ms = ctypes.c_float()
        rc = self.rt.cudaEventElapsedTime(ctypes.byref(ms), self.begin, self.end)
        return float(ms.value) if rc == 0 else float("nan")


def start_clock_sampler(gpu_index: int):
    f = tempfile.NamedTemporaryFile("w+", suffix=".csv", delete=False)
    q = ("clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
         "clocks_event_reasons.sw_power_cap")
    try:
        p = subprocess.Popen(["nvidia-smi", "-i", str(gpu_index), f"--query-gpu={q}", "--format=csv,noheader,nounits",
                              "-lms", "100"], stdout=f, stderr=subprocess.DEVNULL)
    except Exception:
        return None, f.name
    return p, f.name


def stop_clock_sampler(p, path):
    out = {"sm_mhz": None, "sm_max_mhz": None, "reasons": [], "samples": 0}
    if p is not None:
        p.terminate()
        try:
            p.wait(timeout=5)
        except Exception:
            p.kill()
    try:
        rows = [r.strip().split(",") for r in open(path) if r.strip()]
        sm = sorted(float(r[0]) for r in rows if r[0].strip().replace(".", "").isdigit())
        if sm:
            out["sm_mhz"] = sm[len(sm) // 2]
            out["sm_max_mhz"] = float(rows[0][1])
            out["samples"] = len(sm)
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for i, n in enumerate(names):
            if any(len(r) > 4 + i and r[4 + i].strip().lower() == "active" for r in rows):
                out["reasons"].append(n)
        os.unlink(path)
    except Exception:
        pass
    return out


def hbm_peak():
    try:
        peaks = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))
        return float(peaks["hbm_gbs"]), "measured (MEASURED_PEAKS.json hbm_gbs)"
    except Exception:
        return 6650.0, "fallback (B200_PROFILING.md 6.65 TB/s)"


# ------------------------------------------------------------------------------------------------------
def run_ours(a):
    import torch.distributed as dist

    from velocycle_b200 import ppl as pyro
    from velocycle_b200.ppl.infer import SVI, Trace_ELBO
    from velocycle_b200.ppl.optim import ClippedAdam
    from velocycle_b200.preprocessing import make_velocity_metaparams
    from velocycle_b200.sharding import ShardInfo, init_from_env
    from velocycle_b200.synthetic import make_synthetic
    import __graft_entry__ as ge

    rank, world, local_rank = init_from_env()
    if world != a.gpus and world > 1:
        a.gpus = world
    if not torch.cuda.is_available():
        raise SystemExit("bench.py --impl ours needs a CUDA device: velocycle_b200 has no CPU path")
    torch.cuda.set_device(local_rank)
    dev = torch.device("cuda", local_rank)
    if rank == 0:
        ge.build()
    log = (lambda *m: print(f"[bench rank {rank}]", *m, file=sys.stderr, flush=True)) if os.environ.get("VCB_BENCH_VERBOSE") else (lambda *m: None)
    log("built")
    if world > 1:
        dist.barrier()
    log("first barrier passed")

    Nc, Ng = a.cells_per_gpu, a.genes
    shard = ShardInfo(rank, world, rank * Nc, Nc * world) if world > 1 else None
    # data: every rank draws its own cell shard on its own device; gene-level truth is shared (same seed)
    d = make_synthetic(Nc, Ng, H=a.harmonics, Hw=a.omega_harmonics, Nb=a.batches, Nx=a.conditions,
                       seed=0, device=dev, stats=True)
    if world > 1:  # different cells per rank, same genes: redraw the cell-level part with a rank-specific seed
        d2 = make_synthetic(Nc, Ng, H=a.harmonics, Hw=a.omega_harmonics, Nb=a.batches, Nx=a.conditions,
                            seed=1000 + rank, device=dev, stats=True)
        # keep the shared gene-level parameters of seed 0 by regenerating counts is expensive; instead use d2
        # wholesale but overwrite nothing: per-rank gene truths differ slightly, priors below come from rank 0
        d = d2
    mu_nu, sd_nu, phixy, mu_nw, sd_nw = priors_from(d, dev)
    if world > 1:  # replicated priors must be identical on every rank
        for t in (mu_nu, sd_nu, mu_nw, sd_nw):
            dist.broadcast(t, src=0)
    mp = make_velocity_metaparams(d.S[:, :Ng], d.U[:, :Ng], mu_nu, sd_nu, phixy, mu_nw, sd_nw,
                                  batch_id=d.batch_id, cond_id=d.cond_id, Nb=a.batches, Nx=a.conditions,
                                  count_factor=d.cf, model_type=a.model_type, device=dev, shard=shard)
    zero_S, zero_U = d.zero_frac_S, d.zero_frac_U
    log("metaparams ready")
    del d
    torch.cuda.empty_cache()
    counts = mp.packed_counts
    ev = CudaEvents()
    counts.profile_events = ev.handles()

    pyro.clear_param_store()
    pyro.set_rng_seed(0)  # same seed on every rank: replicated draws agree, per-cell noise is sliced (ShardedNormal)
    opt_args = {"lr": 0.03, "lrd": 0.9996, "betas": (0.8, 0.99)}
    if a.eager:
        eager = SVI(mp.model_fn, mp.guide_fn, ClippedAdam(opt_args), Trace_ELBO())
        svi_step = lambda: eager.step(mp)
        launches_per_step, api = 4, "ppl.infer.SVI.step (eager)"
    else:
        from velocycle_b200.svi import GraphedSVI

        graphed = GraphedSVI(mp.model_fn, mp.guide_fn, opt_args, mp)
        svi_step = lambda: graphed.step()
        launches_per_step, api = 6, "svi.GraphedSVI.step (CUDA graph)"

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    sampler, path = start_clock_sampler(local_rank) if rank == 0 else (None, None)
    for _ in range(a.warmup):
        svi_step()
    barrier()
    log("warm-up done")
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    kern_ms = []
    e0.record()
    for _ in range(a.steps):
        svi_step()  # returns the loss as a Python float: one D2H read + sync per step, as in the reference
        kern_ms.append(ev.elapsed_ms())
    e1.record()
    barrier()
    clocks = stop_clock_sampler(sampler, path) if rank == 0 else None
    ms = e0.elapsed_time(e1)
    t = torch.tensor([ms], device=dev)
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    ms_per_step = float(t.item()) / a.steps
    value = Nc * world * Ng / (ms_per_step * 1e-3)

    # ---- roofline of the streaming kernel (rank 0's launches; every rank runs the same shape) ----------
    kms = sorted(k for k in kern_ms if k == k)
    kern = sum(kms) / len(kms) if kms else float("nan")
    alg_bytes = 8.0 * Nc * Ng + 4.0 * Nc * 8  # fp32 S and U once; per-cell inputs/outputs are < 0.1 %
    peak, peak_src = hbm_peak()
    achieved = alg_bytes / (kern * 1e-3) / 1e9 if kern == kern and kern > 0 else None
    traffic = None
    try:
        prof = json.load(open(os.path.join(ROOT, "profiles", "stream_kernel_dram_bytes.json")))
        traffic = prof.get("dram_bytes_per_cell_gene", 0) * Nc * Ng or None
    except Exception:
        pass
    roofline = {"bound": "hbm", "achieved": achieved, "peak": peak, "unit": "GB/s",
                "frac": (achieved / peak) if achieved else None, "traffic": traffic,
                "kernel": ("vcb_umma_stream_kernel (tcgen05, VCB_STREAM_KERNEL=umma)"
                           if os.environ.get("VCB_STREAM_KERNEL", "").startswith("u") else "vcb_stream_kernel"), "kernel_ms": kern, "algorithmic_bytes": alg_bytes, "peak_source": peak_src,
                "note": "latency / issue-slot bound at 16 warps per SM, not HBM bound: every pipe is < 40 % busy (DESIGN.md section 4, profiles/)"}

    # ---- e2e: counts start every step in pinned host memory ----------------------------------------------
    e2e = None
    if not a.no_e2e:
        try:
            import psutil

            from velocycle_b200.fused import HostCounts

            # The step's inputs start in pinned host memory in the narrowest exact integer format (HostCounts: 2- or 4-bit
            # codes with an escape byte stream, or one byte per count; large counts as an (index, value) list); each e2e
            # step copies them to the device, widens them into the float32 [Nc][ld] matrices (vcb_expand_counts[_packed])
            # and runs the SVI step on them.
            hS = HostCounts.from_tensor(counts.S, sub_byte=True)
            hU = HostCounts.from_tensor(counts.U, sub_byte=True)
            h2d = hS.nbytes + hU.nbytes
            need = h2d
            avail = psutil.virtual_memory().available
            if need * world * 2 > avail:
                raise MemoryError(f"{need * world} B of pinned host staging do not fit in {avail} B of host RAM")
            S_check = counts.S[:4096].clone()
            n_e2e = max(2, min(a.steps, 10))
            # Every step's inputs cross PCIe inside the timed region.  The copies run on a second stream into the spare
            # set of device staging buffers (HostCounts.start_upload), so the copy of step i+1 overlaps the SVI step i --
            # what any input pipeline does; the first copy of the region is not overlapped with anything.
            copy_stream = torch.cuda.Stream(device=dev)
            def start_copies():
                hS.start_upload(dev, copy_stream)
                hU.start_upload(dev, copy_stream)
            def e2e_steps(k):
                start_copies()
                for i in range(k):
                    hS.finish_upload(counts.S)
                    hU.finish_upload(counts.U)
                    if i + 1 < k:
                        start_copies()
                    svi_step()
            e2e_steps(2)
            assert torch.equal(S_check, counts.S[:4096]), "HostCounts round trip changed the counts"
            barrier()
            e0.record()
            e2e_steps(n_e2e)
            e1.record()
            barrier()
            t = torch.tensor([e0.elapsed_time(e1)], device=dev)
            if world > 1:
                dist.all_reduce(t, op=dist.ReduceOp.MAX)
            ms_e2e = float(t.item()) / n_e2e
            fmt_name = {1: "u8 + overflow list", 2: "u16", 4: "i32", 16: "2-bit codes + escape bytes + overflow list",
                        32: "4-bit codes + escape bytes + overflow list",
                        64: "2-bit codes + escape nibbles + escape bytes + overflow list"}
            e2e = {"value": Nc * world * Ng / (ms_e2e * 1e-3), "unit": UNIT, "ms_per_step": ms_e2e,
                   "h2d_bytes_per_step": int(h2d), "d2h_bytes_per_step": 4, "steps": n_e2e,
                   "note": ("both count matrices copied from pinned host memory every step in the staging format of "
                            f"velocycle_b200.fused.HostCounts (S: {fmt_name[hS.fmt]}, U: {fmt_name[hU.fmt]}; "
                            f"{(0 if hS.over_idx is None else hS.over_idx.numel()) + (0 if hU.over_idx is None else hU.over_idx.numel())}"
                            " entries in the overflow lists), widened on the device to the float32 layout by vcb_expand_counts[_packed], "
                            "then one GraphedSVI step with the loss read back; the copies of step i+1 run on a copy stream under step i")}
            del hS, hU
        except Exception as exc:  # pragma: no cover
            e2e = {"value": None, "unit": UNIT, "error": repr(exc)}

    # ---- CPU baseline on rank 0 at N = 1 --------------------------------------------------------------------
    cpu = None
    if rank == 0 and world == 1 and not a.no_cpu_baseline:
        try:
            n = a.cpu_sample_cells
            dt, cores = cpu_reference_steps(a, n, steps=4, warmup=1)
            cpu = {"value": n * Ng / dt, "unit": UNIT, "cores": cores, "kind": "port",
                   "sample": f"{n} cells x {Ng} genes of the same workload, 1 warm-up + 4 timed SVI steps of the "
                             f"unfused reference op chain (oracle/models.py) on the host CPU",
                   "ms_per_step_on_sample": dt * 1e3}
        except Exception as exc:  # pragma: no cover
            cpu = {"value": None, "error": repr(exc)}

    if rank == 0:
        line = {
            "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": a.steps, "warmup": a.warmup,
            "ms_per_step": ms_per_step, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
            "dtype": "f32", "data": "synthetic",
            "config": {"workload": workload_name(a), "cells_total": Nc * world, "genes": Ng, "n_mat": 2,
                       "l2_flush": "not needed: 16 GB of counts per step >> 126 MB L2",
                       "zero_fraction_S": zero_S, "zero_fraction_U": zero_U, "parallelism": f"cell-shard x{world}"},
            "svi_steps_per_sec": 1e3 / ms_per_step,
            "roofline": roofline, "cpu_baseline": cpu, "clocks": clocks, "e2e": e2e,
            "gpu_launches": a.steps * launches_per_step,
            "gpu_launches_note": "our kernels per step: vcb_cell_tables, vcb_stream_kernel, vcb_cell_epilogue, "
                                 "vcb_gene_epilogue (+ vcb_adam_tick, vcb_clipped_adam under GraphedSVI)",
            "api": api,
        }
        print(json.dumps(line), flush=True)
    if world > 1:
        # NCCL communicators referenced by a captured CUDA graph do not tear down cleanly
        # (destroy_process_group blocks); everything is printed, so leave without the teardown
        dist.barrier()
        torch.cuda.synchronize()
        sys.stdout.flush()
        sys.stderr.flush()
        os._exit(0)


def main():
    import faulthandler

    # never hang a GPU box: dump every thread's stack and exit if the run exceeds the watchdog
    faulthandler.dump_traceback_later(int(os.environ.get("VCB_BENCH_WATCHDOG_S", "900")), exit=True)
    a = parse()
    if a.impl == "reference":
        run_reference_arm(a)
    else:
        run_ours(a)


if __name__ == "__main__":
    main()
