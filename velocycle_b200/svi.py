"""GraphedSVI: the whole SVI step as one CUDA graph.

``pyro.infer.SVI.step`` (driven by the reference at ``phase_inference_model.py:168-169`` /
``velocity_inference_model.py:118-120``) costs, besides the likelihood, a few hundred tiny launches and about a
dozen host syncs per step: guide sampling, prior log-probs, ``.item()`` per site, autograd bookkeeping and one
``ClippedAdam`` object per parameter tensor.  At 100k cells x 2k genes that host-bound tail is ten times longer
than the fused likelihood kernel.  ``GraphedSVI`` keeps the model / guide functions and their semantics but

* moves every (unconstrained) parameter into ONE flat fp32 buffer (gradients, Adam moments likewise),
* runs the step as a dozen launches when model and guide are the package's own (``faststep.FusedStep``: torch draws in
  the guide's order, ``vcb_svi_sample``, the likelihood, the all-reduce under cell sharding, ``vcb_svi_backward``);
  any other (conditioned, user-written) pair is traced once through the effect handlers while a CUDA graph is being
  captured -- the graph then holds exactly the kernels of one Trace_ELBO step -- and
* ends with the fused multi-tensor ``vcb_clipped_adam`` kernel (device-side step counter, so replays advance it).

``step()`` replays the graph and reads back the loss (one 4-byte D2H copy).  Distribution argument validation is
switched off inside the graph (it would need host syncs); run a few eager ``ppl.infer.SVI`` steps first if you
want Pyro-style validation of a new model.
"""
from __future__ import annotations

import math
from typing import Callable, Dict, Optional

import torch

from . import _lib
from .ppl import backend

__all__ = ["GraphedSVI", "stepper_for", "agree_across_ranks"]


def stepper_for(owner, model: Callable, guide: Callable, optimizer, loss, mp):
    """The step function the fit drivers loop over (``PhaseFitModel.fit`` / ``VelocityFitModel.fit``,
    ``phase_inference_model.py:162-169``, ``velocity_inference_model.py:111-120``): ``GraphedSVI.step`` when the data live on
    a CUDA device, the optimizer is this package's ``ClippedAdam`` and the loss the default single-particle ``Trace_ELBO``
    -- the fused step for the package's own model / guide, a captured trace for conditioned ones; anything else (custom
    losses or optimizers, real Pyro as the backend) gets the eager ``SVI.step``.  The stepper is cached on ``owner`` per
    optimizer object, so a second ``fit`` with the same optimizer continues its Adam state and learning-rate decay like
    Pyro's per-parameter optimizers do."""
    from .ppl import infer as shim_infer, optim as shim_optim

    pyro, _, _, infer, _ = backend.get()
    dev = torch.device(mp.device)
    args = getattr(optimizer, "pt_optim_args", None)
    graphable = (
        dev.type == "cuda" and infer is shim_infer and isinstance(optimizer, shim_optim.PyroOptim)
        and optimizer.pt_optim_constructor is shim_optim.TorchClippedAdam and isinstance(args, dict)
        and (loss is None or (isinstance(loss, shim_infer.Trace_ELBO) and loss.num_particles == 1))
        and float(args.get("weight_decay", 0.0)) == 0.0
    )
    if not graphable:
        loss = infer.Trace_ELBO(num_particles=1) if loss is None else loss
        svi = infer.SVI(model, guide, optimizer, loss)
        return lambda: svi.step(mp)
    cache = getattr(owner, "_steppers", None)
    if cache is None:
        cache = owner._steppers = {}
    key = (id(optimizer), id(model), id(guide))
    if key not in cache:
        cache.clear()  # one live graph per driver: the flat buffers of an older one are re-homed by the new one
        cache[key] = (GraphedSVI(model, guide, dict(args), mp), optimizer)  # (the optimizer is kept alive: its id is the key)
    g = cache[key][0]
    return g.step


def agree_across_ranks(flag: bool, mp) -> bool:
    """Under cell sharding every rank must take the same early-exit decision (a rank that leaves the loop alone hangs the
    others in the step's all-reduce): rank 0's decides."""
    import torch.distributed as dist

    shard = getattr(mp, "shard", None)
    if shard is None or shard.world <= 1 or not dist.is_initialized():
        return flag
    t = torch.tensor([1 if flag else 0], device=torch.device(mp.device) if torch.device(mp.device).type == "cuda" else "cpu")
    dist.broadcast(t, src=dist.get_global_rank(shard.group, 0) if shard.group is not None else 0, group=shard.group)
    return bool(int(t.item()))


class GraphedSVI:
    def __init__(self, model: Callable, guide: Callable, optim_args: Dict, mp, use_graph: bool = True,
                 warmup_iters: int = 3, fast: bool = True, peer_allreduce: bool = True):
        """``fast``: use the fused step (``faststep.FusedStep``: a dozen launches) when ``model`` / ``guide`` are the package's
        own unconditioned functions; otherwise -- and always with ``fast=False`` -- the step is traced through the effect
        handlers like ``pyro.infer.SVI`` does."""
        self.model, self.guide, self.mp = model, guide, mp
        self._want_fast = fast
        self._fast = None
        self.peer_allreduce = peer_allreduce  # fused step under cell sharding: vcb_allreduce_sum instead of NCCL
        self.lr0 = float(optim_args.get("lr", 1e-3))
        self.lrd = float(optim_args.get("lrd", 1.0))
        self.betas = tuple(optim_args.get("betas", (0.9, 0.999)))
        self.eps = float(optim_args.get("eps", 1e-8))
        self.clip = float(optim_args.get("clip_norm", 10.0))
        if float(optim_args.get("weight_decay", 0.0)) != 0.0:
            raise NotImplementedError("weight_decay is not supported by the fused ClippedAdam")
        self.device = torch.device(mp.device)
        if self.device.type != "cuda":
            raise _lib.VcbError("GraphedSVI needs a CUDA device: velocycle_b200 has no CPU path")
        self._lib = _lib.load()
        self._graph: Optional[torch.cuda.CUDAGraph] = None
        self._use_graph = use_graph
        self._warmup_iters = warmup_iters
        self._built = False
        self.steps_done = 0

    # ------------------------------------------------------------------------------------------------------
    def _flatten_params(self) -> None:
        """Create the parameters (one eager guide+model pass) and re-home them into one flat buffer."""
        pyro, _, poutine, _, _ = backend.get()
        with torch.no_grad():
            gt = poutine.trace(self.guide).get_trace(self.mp)
            poutine.trace(poutine.replay(self.model, trace=gt)).get_trace(self.mp)
        store = pyro.get_param_store()
        names = list(store.keys())
        sizes = [store.get_unconstrained(n).numel() for n in names]
        pad = lambda n: (n + 3) // 4 * 4  # every tensor starts on a 16-byte boundary (vector access in the fused kernels)
        total = sum(pad(sz) for sz in sizes)
        dev = self.device
        self.flat_param = torch.zeros(total, dtype=torch.float32, device=dev)
        self.flat_grad = torch.zeros(total, dtype=torch.float32, device=dev)
        self.exp_avg = torch.zeros(total, dtype=torch.float32, device=dev)
        self.exp_avg_sq = torch.zeros(total, dtype=torch.float32, device=dev)
        self.step_dev = torch.zeros(1, dtype=torch.int64, device=dev)
        self.loss_buf = torch.zeros((), dtype=torch.float32, device=dev)
        off = 0
        self.param_slices = {}
        for n, sz in zip(names, sizes):
            old = store.get_unconstrained(n)
            view = self.flat_param[off: off + sz].view(old.shape)
            view.copy_(old.detach())
            view.requires_grad_(True)
            view.grad = self.flat_grad[off: off + sz].view(old.shape)
            store.set_unconstrained(n, view, store.get_constraint(n))
            self.param_slices[n] = (off, sz)
            off += pad(sz)

    def _adam(self) -> None:
        rc = self._lib.vcb_clipped_adam(
            self.flat_param.data_ptr(), self.flat_grad.data_ptr(), self.exp_avg.data_ptr(),
            self.exp_avg_sq.data_ptr(), self.flat_param.numel(), self.step_dev.data_ptr(),
            self.lr0, self.lrd, self.betas[0], self.betas[1], self.eps, self.clip,
            torch.cuda.current_stream(self.device).cuda_stream,
        )
        _lib.check(rc, "vcb_clipped_adam")

    def _body(self) -> None:
        """One Trace_ELBO(num_particles=1) step with every value kept on the device."""
        if self._fast is not None:
            self._fast.body()  # draws, vcb_svi_sample, likelihood, [all-reduce], vcb_svi_backward: loss_buf and flat_grad
            self._adam()
            return
        _, _, poutine, _, _ = backend.get()
        self.flat_grad.zero_()
        guide_trace = poutine.trace(self.guide).get_trace(self.mp)
        model_trace = poutine.trace(poutine.replay(self.model, trace=guide_trace)).get_trace(self.mp)
        model_trace.compute_log_prob()
        guide_trace.compute_log_prob()
        surrogate = 0.0
        for site in model_trace.nodes.values():
            if site["type"] == "sample":
                surrogate = surrogate + site["log_prob_sum"]
        for site in guide_trace.nodes.values():
            if site["type"] == "sample":
                surrogate = surrogate - site["log_prob_sum"]
        loss = -surrogate
        loss.backward()
        shard = getattr(self.mp, "shard", None)
        if shard is not None and shard.world > 1:
            # the likelihood sites are global already (all-reduced inside the fused op) and the gene-level terms replicated;
            # the per-cell site is the rank's own: report the global ELBO, identical on every rank
            import torch.distributed as dist

            local = 0.0
            for tr, sign in ((model_trace, 1.0), (guide_trace, -1.0)):
                site = tr.nodes.get("ϕxy")
                if site is not None and site["type"] == "sample":
                    local = local + sign * site["log_prob_sum"].detach()
            if isinstance(local, torch.Tensor):
                total = local.clone()
                dist.all_reduce(total, group=shard.group)
                loss = loss.detach() + local - total
        self.loss_buf.copy_(loss.detach())
        self._adam()

    def _build(self) -> None:
        from .ppl import primitives

        # creating the parameters runs the guide once; put the RNG streams back so that step 1 sees the draws
        # a fresh ``SVI.step`` would see
        rng_cuda, rng_cpu = torch.cuda.get_rng_state(self.device), torch.get_rng_state()
        self._flatten_params()
        torch.cuda.set_rng_state(rng_cuda, self.device)
        torch.set_rng_state(rng_cpu)
        if self._want_fast:
            from .faststep import FusedStep, model_code

            found = model_code(self.model, self.guide, self.mp)
            if found is not None:
                self._fast = FusedStep(self, found[0], found[1])
        self._validation_prev = primitives.validation_enabled()
        if not self._use_graph:
            self._built = True
            return
        primitives.enable_validation(False)
        prev = torch.distributions.Distribution._validate_args
        torch.distributions.Distribution.set_default_validate_args(False)
        try:
            # warm-up on a side stream (lazy kernel attribute setup, allocator pools), with state restored after
            snap = (self.flat_param.clone(), self.exp_avg.clone(), self.exp_avg_sq.clone(), self.step_dev.clone())
            rng = torch.cuda.get_rng_state(self.device)
            s = torch.cuda.Stream(device=self.device)
            s.wait_stream(torch.cuda.current_stream(self.device))
            with torch.cuda.stream(s):
                for _ in range(self._warmup_iters):
                    self._body()
            torch.cuda.current_stream(self.device).wait_stream(s)
            torch.cuda.synchronize(self.device)
            with torch.no_grad():
                self.flat_param.copy_(snap[0])
                self.exp_avg.copy_(snap[1])
                self.exp_avg_sq.copy_(snap[2])
                self.step_dev.copy_(snap[3])
            torch.cuda.set_rng_state(rng, self.device)
            self._graph = torch.cuda.CUDAGraph()
            with torch.cuda.graph(self._graph):
                self._body()
        finally:
            torch.distributions.Distribution.set_default_validate_args(prev)
            primitives.enable_validation(self._validation_prev)
        self._built = True

    # ------------------------------------------------------------------------------------------------------
    def step(self, *args, sync: bool = True):
        """One SVI step.  Returns the ELBO loss as a float (``sync=True``, like ``pyro.infer.SVI.step``) or the
        device scalar that will hold it (``sync=False``)."""
        if not self._built:
            self._build()
        if self._graph is not None:
            self._graph.replay()
        else:
            self._body()
        self.steps_done += 1
        return self.loss_buf.item() if sync else self.loss_buf

    def current_lr(self) -> float:
        return self.lr0 * self.lrd ** self.steps_done
