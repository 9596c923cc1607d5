#!/usr/bin/env python
"""bench.py -- the headline benchmark of the hot path (contract: the task statement / DESIGN.md section 6).

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference]
    python -m torch.distributed.run --nnodes=1 --nproc-per-node N ... bench.py --gpus N ...

A "step" is one full SVI step of the velocity model (guide draws -> fused ELBO + every gradient over the spliced and
unspliced count matrices -> gene-gradient all-reduce when N > 1 -> ClippedAdam) through the step function the reference's
public entry point loops over: ``VelocityFitModel.fit`` obtains it from ``svi.stepper_for`` (a ``GraphedSVI.step``: the
whole step replayed as one CUDA graph, loss read back every step like ``pyro.infer.SVI.step``).

Workloads (BASELINE.json):
  N = 1 : 1,000,000 cells x 2,000 genes, H = 3, Hw = 1 -- the single-GPU target shape of north_star (16 GB of fp32 counts,
          far beyond the 126 MB L2: no flush needed).
  N > 1 : config C4, 2,000,000 cells x 2,000 genes in total, cell-sharded (2M / N cells per GPU): STRONG scaling; the weak
          figure (1M cells per GPU) rides along as ``weak``.  Before timing, 6 steps of the golden case ``case_multi`` are
          run sharded and unsharded and compared (``shard_parity``).
``value``       : cell.gene negative-binomial evaluations per second (cells x genes x steps/s; S and U count as one), counts
                  resident in HBM (as the reference keeps them: preprocessing.py:193-194).
``e2e``         : the same step when the count matrices start each step in pinned HOST memory and are copied to the device
                  inside the timed region, plus the device->host read of the loss.
``roofline``    : the streaming kernel's algorithmic bytes (8 B per cell.gene: one fp32 S and U each) over its CUDA-event
                  duration, against the measured HBM peak of MEASURED_PEAKS.json.
``breakdown_ms``: CUDA-event times of the phases of one (un-graphed) step: draws / sample / likelihood / allreduce /
                  backward / adam.
``configs``     : (N = 1) the other named configurations: C3 100k x 2k, the per-GPU shard of C5 (62.5k x 5k, 16 batches, 2
                  conditions), a tutorial-sized 2,000 x 218 H = 1 two-sample fit, and the PHASE model at 1M x 2k with its
                  own roofline (4 B per cell.gene).
``cpu_baseline``: the reference op chain (oracle/models.py: unfused (Ng,Nc) einsums + GammaPoisson + autograd under the same
                  SVI) on the host cores on a bounded sample (20,000 x 2,000), extrapolated linearly
                  per cell.gene; ``gpu_unfused`` = the same unfused chain on this GPU at 100k x 2k.
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

_RESULT_FD = 1
METRIC = "cell_gene_nb_evals_per_sec"
UNIT = "cell*gene/s"
OPT_ARGS = {"lr": 0.03, "lrd": 0.9996, "betas": (0.8, 0.99)}


def parse():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--scaling", default="auto", choices=["auto", "strong", "weak"],
                    help="N > 1: strong = --total-cells split over the GPUs (default), weak = --cells-per-gpu on each")
    ap.add_argument("--cells-per-gpu", type=int, default=1_000_000)
    ap.add_argument("--total-cells", type=int, default=2_000_000, help="config C4: the strong-scaling problem")
    ap.add_argument("--genes", type=int, default=2000)
    ap.add_argument("--harmonics", type=int, default=3)
    ap.add_argument("--omega-harmonics", type=int, default=1)
    ap.add_argument("--batches", type=int, default=1)
    ap.add_argument("--conditions", type=int, default=1)
    ap.add_argument("--model-type", default="lrmn", choices=["lrmn", "normal"])
    ap.add_argument("--cpu-sample-cells", type=int, default=20000)
    ap.add_argument("--traced", action="store_true", help="GraphedSVI(fast=False): the step traced through the effect handlers")
    ap.add_argument("--no-e2e", action="store_true")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-configs", action="store_true")
    ap.add_argument("--profiler-range", action="store_true",
                    help="cudaProfilerStart/Stop around the timed steps of the headline workload (ncu --profile-from-start off)")
    return ap.parse_args()


def workload_name(cells, genes, a, model="velocity", Nb=None, Nx=None, H=None):
    return (f"synthetic {model} SVI step, {cells} cells x {genes} genes, H={a.harmonics if H is None else H}, "
            f"Hw={a.omega_harmonics}, Nb={a.batches if Nb is None else Nb}, Nx={a.conditions if Nx is None else Nx}"
            + (f", guide={a.model_type}" if model == "velocity" else ""))


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
# reference op chain (oracle/models.py) under the same SVI: the CPU arm and the "unfused ATen on this GPU" comparator
# ------------------------------------------------------------------------------------------------------
def reference_chain_steps(a, n_cells, steps, warmup, device="cpu"):
    """Seconds per SVI step of the unfused reference chain over ``n_cells`` cells of the workload on ``device``."""
    from oracle import models as omodels
    from velocycle_b200 import ppl as pyro
    from velocycle_b200.ppl.infer import SVI, Trace_ELBO
    from velocycle_b200.ppl.optim import ClippedAdam
    from velocycle_b200.preprocessing import make_velocity_metaparams
    from velocycle_b200.synthetic import make_synthetic

    cores = os.cpu_count() or 1
    if device == "cpu":
        torch.set_num_threads(cores)
    d = make_synthetic(n_cells, a.genes, H=a.harmonics, Hw=a.omega_harmonics, Nb=a.batches, Nx=a.conditions,
                       seed=0, device=device, stats=False)
    mu_nu, sd_nu, phixy, mu_nw, sd_nw = priors_from(d, device)
    mp = make_velocity_metaparams(d.S[:, : a.genes], d.U[:, : a.genes], mu_nu, sd_nu, phixy, mu_nw, sd_nw,
                                  batch_id=d.batch_id, cond_id=d.cond_id, Nb=a.batches, Nx=a.conditions,
                                  count_factor=d.cf, model_type=a.model_type, device=device)
    model = omodels.velocity_model_unfused_lrmn if a.model_type == "lrmn" else omodels.velocity_model_unfused
    pyro.clear_param_store()
    pyro.set_rng_seed(0)
    svi = SVI(model, mp.guide_fn, ClippedAdam(dict(OPT_ARGS)), Trace_ELBO())
    sync = (lambda: torch.cuda.synchronize()) if device != "cpu" else (lambda: None)
    for _ in range(warmup):
        svi.step(mp)
    sync()
    t0 = time.perf_counter()
    for _ in range(steps):
        svi.step(mp)
    sync()
    dt = (time.perf_counter() - t0) / max(steps, 1)
    pyro.clear_param_store()
    return dt, cores


def run_reference_arm(a):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    n = a.cpu_sample_cells
    dt, cores = reference_chain_steps(a, n, a.steps, a.warmup)
    value = n * a.genes / dt
    cells = a.cells_per_gpu if a.gpus == 1 else a.total_cells
    sample = (f"{n} cells x {a.genes} genes of the workload per step (same generator, seed 0), unfused reference op "
              f"chain + autograd under SVI on the host CPU, {cores} threads; cell.gene/s is size-independent, so the value "
              "stands for the full workload (extrapolated linearly, BASELINE.md section 3)")
    line = {
        "impl": "reference", "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": a.gpus, "steps": a.steps,
        "warmup": a.warmup, "ms_per_step": dt * 1e3, "higher_is_better": True,
        "scaling": "weak" if a.gpus == 1 else "strong", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "config": {"workload": workload_name(cells, a.genes, a), "sample": sample, "extrapolated": True},
        "svi_steps_per_sec_on_sample": 1.0 / dt,
        "cpu_baseline": {"value": value, "unit": UNIT, "cores": cores, "kind": "port", "sample": sample},
        "e2e": {"value": value, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
    }
    emit(line)


def emit(line) -> None:
    os.write(_RESULT_FD, (json.dumps(line) + "\n").encode())


# ------------------------------------------------------------------------------------------------------
# helpers for the GPU arm
# ------------------------------------------------------------------------------------------------------
class CudaEvents:
    """Raw cudaEvent_t pair (the C ABI records it around the streaming kernel, as event-record nodes inside a graph)."""

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


def dram_traffic_per_cell_gene():
    try:
        prof = json.load(open(os.path.join(ROOT, "profiles", "stream_kernel_dram_bytes.json")))
        return float(prof.get("dram_bytes_per_cell_gene", 0)) or None
    except Exception:
        return None


OUR_KERNELS_FAST = ("vcb_svi_cell_sample, vcb_svi_gene_sample, vcb_cell_tables, vcb_stream2, vcb_cell_epilogue, "
                    "vcb_gene_epilogue, vcb_svi_cell_backward, vcb_svi_gene_backward, vcb_svi_finalize, vcb_adam_tick, "
                    "vcb_clipped_adam")


class Workload:
    """One model on one synthetic dataset on this rank: data, metaparameters, the fit driver and its step function."""

    def __init__(self, a, Nc, Ng, dev, rank=0, world=1, model="velocity", Nb=None, Nx=None, H=None, data=None):
        from velocycle_b200 import ppl as pyro
        from velocycle_b200.phase_inference_model import PhaseFitModel
        from velocycle_b200.ppl.optim import ClippedAdam
        from velocycle_b200.preprocessing import make_phase_metaparams, make_velocity_metaparams
        from velocycle_b200.sharding import ShardInfo
        from velocycle_b200.svi import stepper_for
        from velocycle_b200.synthetic import make_synthetic
        from velocycle_b200.velocity_inference_model import VelocityFitModel
        import torch.distributed as dist

        Nb = a.batches if Nb is None else Nb
        Nx = a.conditions if Nx is None else Nx
        H = a.harmonics if H is None else H
        self.Nc, self.Ng, self.world, self.model, self.dev = Nc, Ng, world, model, dev
        shard = ShardInfo(rank, world, rank * Nc, Nc * world) if world > 1 else None
        # every rank draws its own cell shard on its own device (seed = base + rank); priors of the replicated gene-level
        # parameters come from rank 0
        d = data if data is not None else make_synthetic(Nc, Ng, H=H, Hw=a.omega_harmonics, Nb=Nb, Nx=Nx,
                                                         seed=(1000 + rank) if world > 1 else 0, device=dev, stats=True)
        mu_nu, sd_nu, phixy, mu_nw, sd_nw = priors_from(d, dev)
        if world > 1:
            for t in (mu_nu, sd_nu, mu_nw, sd_nw):
                dist.broadcast(t, src=0)
        if model == "velocity":
            self.mp = make_velocity_metaparams(d.S[:, :Ng], d.U[:, :Ng], mu_nu, sd_nu, phixy, mu_nw, sd_nw,
                                               batch_id=d.batch_id, cond_id=d.cond_id, Nb=Nb, Nx=Nx, count_factor=d.cf,
                                               model_type=a.model_type, device=dev, shard=shard)
            self.driver = VelocityFitModel(self.mp, get_posterior=False)
        else:
            self.mp = make_phase_metaparams(d.S[:, :Ng], None, mu_nu, sd_nu, phixy, batch_id=d.batch_id, Nb=Nb,
                                            count_factor=d.cf, device=dev, shard=shard)
            self.driver = PhaseFitModel(self.mp, get_posterior=False)
        self.zero_S, self.zero_U = getattr(d, "zero_frac_S", None), getattr(d, "zero_frac_U", None)
        self.counts = self.mp.packed_counts
        self.ev = CudaEvents()
        self.counts.profile_events = self.ev.handles()
        pyro.clear_param_store()
        pyro.set_rng_seed(0)  # same seed on every rank: replicated draws agree, per-cell noise is the rank's slice
        self.optimizer = ClippedAdam(dict(OPT_ARGS))
        if a.traced:
            from velocycle_b200.svi import GraphedSVI

            self.gsvi = GraphedSVI(self.driver.model, self.driver.guide, dict(OPT_ARGS), self.mp, fast=False)
            self.step = self.gsvi.step
        else:
            self.step = stepper_for(self.driver, self.driver.model, self.driver.guide, self.optimizer, None, self.mp)
            self.gsvi = self.step.__self__
        self.n_mat = 2 if model == "velocity" else 1

    def barrier(self):
        import torch.distributed as dist

        if self.world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def time_steps(self, steps, warmup, fn=None, profiler_range=False):
        """ms per step (CUDA events, MAX over ranks) and the streaming kernel's mean event duration on this rank."""
        import torch.distributed as dist

        fn = fn or self.step
        for _ in range(warmup):
            fn()
        self.barrier()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        kern = []
        if profiler_range:
            torch.cuda.cudart().cudaProfilerStart()
        e0.record()
        for _ in range(steps):
            fn()  # returns the loss as a Python float: one D2H read + sync per step, as pyro.infer.SVI.step does
            kern.append(self.ev.elapsed_ms())
        e1.record()
        self.barrier()
        if profiler_range:
            torch.cuda.cudart().cudaProfilerStop()
        t = torch.tensor([e0.elapsed_time(e1)], device=self.dev)
        if self.world > 1:
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
        kern = [k for k in kern if k == k]
        return float(t.item()) / steps, (sum(kern) / len(kern) if kern else float("nan"))

    def roofline(self, kern_ms):
        bytes_per = 4.0 * self.n_mat
        alg = bytes_per * self.Nc * self.Ng + 4.0 * self.Nc * 8  # fp32 counts once; per-cell inputs / outputs are < 0.1 %
        peak, src = hbm_peak()
        achieved = alg / (kern_ms * 1e-3) / 1e9 if kern_ms == kern_ms and kern_ms > 0 else None
        tr = dram_traffic_per_cell_gene()
        return {"bound": "hbm", "achieved": achieved, "peak": peak, "unit": "GB/s",
                "frac": (achieved / peak) if achieved else None,
                "traffic": (tr * self.Nc * self.Ng * self.n_mat / 2.0) if tr else None,
                "kernel": "vcb_stream2_kernel", "kernel_ms": kern_ms, "algorithmic_bytes": alg, "peak_source": src,
                "bytes_per_cell_gene": bytes_per}

    def breakdown(self, reps=5):
        """CUDA-event times of the phases of an un-graphed fused step (mean of ``reps``), ms."""
        g = self.gsvi
        if getattr(g, "_fast", None) is None:
            return None
        names = ["draws", "sample", "likelihood", "allreduce", "backward", "adam"]
        acc = {n: 0.0 for n in names}
        for r in range(reps + 1):
            evs = [torch.cuda.Event(enable_timing=True)]
            evs[0].record()

            def mark(_name):
                e = torch.cuda.Event(enable_timing=True)
                e.record()
                evs.append(e)

            g._fast.body(mark)
            g._adam()
            mark("adam")
            torch.cuda.synchronize()
            if r:  # the first repetition warms the un-graphed path
                for i, n in enumerate(names):
                    acc[n] += evs[i].elapsed_time(evs[i + 1]) / reps
        acc["stream_kernel"] = self.ev.elapsed_ms()
        return acc

    def posterior_seconds(self, num_samples=100, n_per_bin=50):
        """Wall time of the posterior extraction the fit drivers do after the loop (velocity_inference_model.py:201-224):
        ``num_samples`` guide draws replayed through the model in bins of ``n_per_bin``, the reference's return sites, moved
        to the CPU.  The count sites are skipped (nobody asks for them), so no count byte is streamed."""
        rs = self.driver._return_sites() if self.model == "velocity" else ["ν", "ϕxy", "ϕ", "ζ", "shape_inv"]
        self.driver.sample_posterior(num_samples=2, rs=rs)
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        nbytes = 0
        for _ in range(max(1, num_samples // n_per_bin)):
            out = self.driver.sample_posterior(num_samples=n_per_bin, rs=rs)
            nbytes += sum(v.numel() * v.element_size() for v in out.values())
        torch.cuda.synchronize()
        return time.perf_counter() - t0, nbytes

    def posterior_device_seconds(self, num_samples=500, n_per_bin=50):
        """The same extraction kept on the device (``fastposterior.batched_posterior``; what the driver computes before its
        ``.cpu()``): seconds for ``num_samples`` draws of the reference's return sites."""
        from velocycle_b200.faststep import model_code
        from velocycle_b200.fastposterior import batched_posterior

        found = model_code(self.driver.model, self.driver.guide, self.mp)
        if found is None:
            return None
        rs = self.driver._return_sites() if self.model == "velocity" else ["ν", "ϕxy", "ϕ", "ζ", "shape_inv"]
        batched_posterior(self.mp, found[0], found[1], 2, rs, counts=self.counts)
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        nbytes = 0
        for _ in range(max(1, num_samples // n_per_bin)):
            out = batched_posterior(self.mp, found[0], found[1], n_per_bin, rs, counts=self.counts)
            nbytes += sum(v.numel() * v.element_size() for v in out.values())
            del out
        torch.cuda.synchronize()
        return {"seconds": time.perf_counter() - t0, "draws": num_samples, "bytes_produced": nbytes, "count_bytes_streamed": 0}

    def close(self):
        from velocycle_b200 import ppl as pyro

        self.driver._steppers = {}
        self.gsvi = self.step = self.driver = self.mp = self.counts = None
        pyro.clear_param_store()
        torch.cuda.empty_cache()


def conditioned_step_ms(a, w, dev, steps=20):
    """ms per step of the velocity fit conditioned the way the tutorials condition it (sites phixy, nu, shape_inv, Delta-nu
    at fixed values, hidden from the guide) on the data of workload ``w``: the fused step with four sites held."""
    from velocycle_b200 import ppl as pyro
    from velocycle_b200.ppl.optim import ClippedAdam
    from velocycle_b200.svi import stepper_for
    from velocycle_b200.velocity_inference_model import VelocityFitModel

    mp = w.mp
    cond = {"ϕxy": mp.φxy_prior.detach().clone(), "ν": mp.μνg.detach().clone(),
            "shape_inv": torch.full((mp.Ng, 1), 0.5, device=dev), "Δν": torch.zeros((mp.Nb, 1, 1, mp.Ng, 1), device=dev)}
    driver = VelocityFitModel(mp, condition_on=cond, get_posterior=False)
    pyro.clear_param_store()
    pyro.set_rng_seed(0)
    step = stepper_for(driver, driver.model, driver.guide, ClippedAdam(dict(OPT_ARGS)), None, mp)
    for _ in range(3):
        step()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(steps):
        step()
    e1.record()
    torch.cuda.synchronize()
    fast = getattr(step.__self__, "_fast", None) is not None
    driver._steppers = {}
    return {"ms_per_step": e0.elapsed_time(e1) / steps, "fused_step": fast}


def shard_parity(dev, rank, world):
    """6 steps of the golden case `case_multi` (velocity, LRMN guide, 4 batches, 2 conditions) cell-sharded over the ranks
    against the same fit unsharded (run redundantly on every rank): parameters must agree (5e-3 absolute, the tolerance of
    tests/test_sharding_gpu.py -- ClippedAdam's normalised update amplifies last-bit gradient differences)."""
    import numpy as np
    import torch.distributed as dist

    from velocycle_b200 import ppl as pyro
    from velocycle_b200.preprocessing import make_velocity_metaparams
    from velocycle_b200.sharding import ShardInfo, shard_cells
    from velocycle_b200.svi import GraphedSVI

    z = np.load(os.path.join(ROOT, "tests", "golden", "case_multi.npz"), allow_pickle=False)
    inp = {k[3:]: torch.as_tensor(z[k].astype(np.int64) if z[k].dtype == np.uint16 else z[k]) for k in z.files
           if k.startswith("in/")}
    Nc = inp["S"].shape[0]

    def fit(a, b, shard):
        sl = slice(a, b)
        mp = make_velocity_metaparams(inp["S"][sl], inp["U"][sl], inp["mu_nu"], inp["sd_nu"], inp["phixy_prior"][sl],
                                      inp["mu_nw"], inp["sd_nw"], batch_id=inp["batch_id"][sl], cond_id=inp["cond_id"][sl],
                                      Nb=int(inp["Nb"]), Nx=int(inp["Nx"]), count_factor=inp["cf"][sl], model_type="lrmn",
                                      device=dev, shard=shard)
        pyro.clear_param_store()
        pyro.set_rng_seed(2024)
        g = GraphedSVI(mp.model_fn, mp.guide_fn, {"lr": 0.03, "lrd": 0.999, "betas": (0.8, 0.99)}, mp, use_graph=False)
        losses = [g.step() for _ in range(6)]
        return losses, {k: v.detach().clone() for k, v in pyro.get_param_store().named_parameters()}

    a, b = shard_cells(Nc, rank, world)
    l_sh, p_sh = fit(a, b, ShardInfo.make(Nc))
    l_one, p_one = fit(0, Nc, None)
    worst = 0.0
    for name, ref in p_one.items():
        got = p_sh[name]
        if name == "ϕxy_locs":
            ref = ref[a:b]
        finite = torch.isfinite(ref)
        if finite.any():
            worst = max(worst, float((got[finite] - ref[finite]).abs().max()))
    loss_rel = max(abs(x - y) / abs(y) for x, y in zip(l_sh, l_one))
    t = torch.tensor([worst, loss_rel], device=dev)
    dist.all_reduce(t, op=dist.ReduceOp.MAX)
    pyro.clear_param_store()
    return {"max_abs": float(t[0]), "loss_max_rel": float(t[1]), "ok": bool(t[0] <= 5e-3 and t[1] <= 1e-4), "steps": 6,
            "case": "tests/golden/case_multi.npz, velocity LRMN", "tolerance": "parameters 5e-3 abs, losses 1e-4 rel"}


def e2e_run(w, a, n_steps):
    """The step with both count matrices starting in pinned host memory (HostCounts staging formats), copies inside the timed
    region (the copy of step i+1 runs on a copy stream under step i), widened on the device, loss read back."""
    import psutil
    import torch.distributed as dist

    from velocycle_b200.fused import HostCounts

    counts, dev = w.counts, w.dev
    hS = HostCounts.from_tensor(counts.S, sub_byte=True)
    hU = HostCounts.from_tensor(counts.U, sub_byte=True)
    h2d = hS.nbytes + hU.nbytes
    if h2d * w.world * 2 > psutil.virtual_memory().available:
        raise MemoryError("pinned host staging does not fit in host RAM")
    S_check = counts.S[:4096].clone()
    copy_stream = torch.cuda.Stream(device=dev)

    def start_copies():
        hS.start_upload(dev, copy_stream)
        hU.start_upload(dev, copy_stream)

    def steps(k):
        start_copies()
        for i in range(k):
            hS.finish_upload(counts.S)
            hU.finish_upload(counts.U)
            if i + 1 < k:
                start_copies()
            w.step()

    steps(2)
    assert torch.equal(S_check, counts.S[:4096]), "HostCounts round trip changed the counts"
    w.barrier()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    steps(n_steps)
    e1.record()
    w.barrier()
    t = torch.tensor([e0.elapsed_time(e1)], device=dev)
    if w.world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    ms = float(t.item()) / n_steps
    fmt = {1: "u8 + overflow list", 2: "u16", 4: "i32", 16: "2-bit codes + escape bytes", 32: "4-bit codes + escape bytes",
           64: "2-bit codes + escape nibbles + escape bytes"}
    return {"value": w.Nc * w.world * w.Ng / (ms * 1e-3), "unit": UNIT, "ms_per_step": ms, "h2d_bytes_per_step": int(h2d),
            "d2h_bytes_per_step": 4, "steps": n_steps,
            "note": (f"S and U copied from pinned host memory every step (HostCounts: S {fmt[hS.fmt]}, U {fmt[hU.fmt]}), widened on "
                     "the device by vcb_expand_counts*, then one step with the loss read back; PCIe-bound")}


# ------------------------------------------------------------------------------------------------------
def run_ours(a):
    import torch.distributed as dist

    from velocycle_b200.sharding import init_from_env
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
    if world > 1:
        dist.barrier()
    scaling = a.scaling if a.scaling != "auto" else ("strong" if world > 1 else "weak")
    Ng = a.genes
    Nc = a.cells_per_gpu if (world == 1 or scaling == "weak") else a.total_cells // world

    parity = shard_parity(dev, rank, world) if world > 1 else None

    # ---- the headline workload ---------------------------------------------------------------------------------------
    w = Workload(a, Nc, Ng, dev, rank, world)
    sampler, path = start_clock_sampler(local_rank) if rank == 0 else (None, None)
    ms_per_step, kern = w.time_steps(a.steps, a.warmup, profiler_range=a.profiler_range)
    clocks = stop_clock_sampler(sampler, path) if rank == 0 else None
    value = Nc * world * Ng / (ms_per_step * 1e-3)
    roofline = w.roofline(kern)
    roofline["note"] = ("latency / pipe-contention bound at 16 warps per SM, not HBM bound: no pipe above 50 % busy (DESIGN.md "
                        "section 4, profiles/r02_stream_kernel_ncu_summary.txt)")
    breakdown = w.breakdown()
    zero_S, zero_U = w.zero_S, w.zero_U
    e2e = None
    if not a.no_e2e:
        try:
            e2e = e2e_run(w, a, max(2, min(a.steps, 10)))
        except Exception as exc:  # pragma: no cover
            e2e = {"value": None, "unit": UNIT, "error": repr(exc)}
    fast = getattr(w.gsvi, "_fast", None) is not None
    posterior = None
    if world == 1:
        try:
            posterior = w.posterior_device_seconds(500, 50)
        except Exception as exc:  # pragma: no cover
            posterior = {"error": repr(exc)}
    w.close()

    # ---- N > 1: the weak-scaling figure next to the strong one ---------------------------------------------------------
    weak = None
    if world > 1 and scaling == "strong" and a.cells_per_gpu != Nc:
        ww = Workload(a, a.cells_per_gpu, Ng, dev, rank, world)
        ms_w, kern_w = ww.time_steps(max(5, a.steps // 2), a.warmup)
        weak = {"scaling": "weak", "cells_per_gpu": a.cells_per_gpu, "ms_per_step": ms_w, "kernel_ms": kern_w,
                "value": a.cells_per_gpu * world * Ng / (ms_w * 1e-3), "unit": UNIT, "breakdown_ms": ww.breakdown()}
        ww.close()

    # ---- N = 1: the other named configurations, the unfused chain on this GPU, the CPU arm --------------------------------
    configs, cpu = None, None
    if world == 1 and not a.no_configs:
        configs = {}
        specs = [
            ("C3_100k_x_2k", dict(Nc=100_000, Ng=2000, model="velocity")),
            ("C5_shard_62500_x_5k_Nb16_Nx2", dict(Nc=62_500, Ng=5000, model="velocity", Nb=16, Nx=2)),
            ("tutorial_2000_x_218_H1_Nb2_Nx2", dict(Nc=2000, Ng=218, model="velocity", Nb=2, Nx=2, H=1)),
            ("phase_1M_x_2k", dict(Nc=a.cells_per_gpu, Ng=2000, model="phase")),
        ]
        for name, kw in specs:
            try:
                c = Workload(a, kw["Nc"], kw["Ng"], dev, model=kw["model"], Nb=kw.get("Nb"), Nx=kw.get("Nx"), H=kw.get("H"))
                ms_c, kern_c = c.time_steps(a.steps, a.warmup)
                rl = c.roofline(kern_c)
                if kw["model"] == "phase":
                    rl["note"] = "phase model: 4 B and 3 MUFU per cell.gene (S only): the same issue-slot bound, half the bytes"
                configs[name] = {"workload": workload_name(kw["Nc"], kw["Ng"], a, kw["model"], kw.get("Nb"), kw.get("Nx"), kw.get("H")),
                                 "ms_per_step": ms_c, "svi_steps_per_sec": 1e3 / ms_c,
                                 "value": kw["Nc"] * kw["Ng"] / (ms_c * 1e-3), "unit": UNIT, "kernel_ms": kern_c,
                                 "roofline_frac": rl["frac"], "roofline": rl, "breakdown_ms": c.breakdown()}
                if name.startswith("C3"):
                    # the same data with the tutorial's conditioning (velocity stage on the phase stage's phixy, nu, shape_inv,
                    # Delta-nu) and the posterior extraction that follows a fit
                    sec, nbytes = c.posterior_seconds(100, 50)
                    configs[name]["posterior_100_draws"] = {"seconds": sec, "bytes_to_host": nbytes, "count_bytes_streamed": 0}
                    configs[name]["conditioned_ms_per_step"] = conditioned_step_ms(a, c, dev)
                c.close()
            except Exception as exc:  # pragma: no cover
                configs[name] = {"error": repr(exc)}
    if rank == 0 and world == 1 and not a.no_cpu_baseline:
        try:
            n = a.cpu_sample_cells
            dt, cores = reference_chain_steps(a, n, steps=4, warmup=1)
            cpu = {"value": n * Ng / dt, "unit": UNIT, "cores": cores, "kind": "port", "extrapolated": True,
                   "sample": f"{n} cells x {Ng} genes of the same workload, 1 warm-up + 4 timed SVI steps of the unfused "
                             "reference op chain (oracle/models.py) on the host CPU; linear in cells, so the value stands for "
                             "the full workload",
                   "ms_per_step_on_sample": dt * 1e3}
            try:  # the same unfused ATen chain on this GPU at 100k x 2k
                n_gpu = 100_000
                dtg, _ = reference_chain_steps(a, n_gpu, steps=3, warmup=1, device=dev)
                cpu["gpu_unfused"] = {"value": n_gpu * Ng / dtg, "unit": UNIT, "ms_per_step": dtg * 1e3,
                                      "sample": f"{n_gpu} cells x {Ng} genes, the same unfused chain (eager SVI) on this B200"}
            except Exception as exc:  # pragma: no cover
                cpu["gpu_unfused"] = {"error": repr(exc)}
        except Exception as exc:  # pragma: no cover
            cpu = {"value": None, "error": repr(exc)}

    if rank == 0:
        line = {
            "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": a.steps, "warmup": a.warmup,
            "ms_per_step": ms_per_step, "higher_is_better": True, "scaling": scaling, "vs_baseline": None,
            "dtype": "f32", "data": "synthetic",
            "config": {"workload": workload_name(Nc * world, Ng, a), "cells_total": Nc * world, "cells_per_gpu": Nc,
                       "genes": Ng, "n_mat": 2,
                       "l2_flush": f"not needed: {8e-9 * Nc * Ng:.1f} GB of counts per GPU and step >> 126 MB L2",
                       "zero_fraction_S": zero_S, "zero_fraction_U": zero_U, "parallelism": f"cell-shard x{world}"},
            "svi_steps_per_sec": 1e3 / ms_per_step,
            "roofline": roofline, "breakdown_ms": breakdown, "cpu_baseline": cpu, "clocks": clocks, "e2e": e2e,
            "gpu_launches": a.steps * (11 if fast else 6),
            "gpu_launches_note": ("our kernels per step: " + OUR_KERNELS_FAST) if fast else
                                 "our kernels per step: vcb_cell_tables, vcb_stream2, vcb_cell_epilogue, vcb_gene_epilogue, "
                                 "vcb_adam_tick, vcb_clipped_adam (the rest of the traced step is torch)",
            "api": "the step function of VelocityFitModel.fit (svi.stepper_for -> GraphedSVI.step, one CUDA graph per step"
                   + (", fused step)" if fast else ", traced step)"),
        }
        if posterior is not None:
            line["posterior_500_draws_on_device"] = posterior
        if weak is not None:
            line["weak"] = weak
        if parity is not None:
            line["shard_parity"] = parity
        if configs is not None:
            line["configs"] = configs
        emit(line)
    if world > 1:
        # NCCL communicators referenced by a captured CUDA graph do not tear down cleanly (destroy_process_group blocks);
        # everything is printed, so leave without the teardown
        dist.barrier()
        torch.cuda.synchronize()
        sys.stdout.flush()
        sys.stderr.flush()
        os._exit(0)


def main():
    import faulthandler

    # stdout carries exactly ONE line, the JSON result: everything else that writes to file descriptor 1 (NCCL's version
    # banner, library chatter) is sent to stderr
    global _RESULT_FD
    sys.stdout.flush()
    _RESULT_FD = os.dup(1)
    os.dup2(2, 1)

    # never hang a GPU box: dump every thread's stack and exit if the run exceeds the watchdog
    faulthandler.dump_traceback_later(int(os.environ.get("VCB_BENCH_WATCHDOG_S", "1200")), exit=True)
    a = parse()
    if a.impl == "reference":
        run_reference_arm(a)
    else:
        run_ours(a)


if __name__ == "__main__":
    main()
