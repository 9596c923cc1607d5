"""GPU, 2 ranks over NCCL (skipped on a single-GPU box): a cell-sharded fit equals the single-GPU fit."""
import os
import socket

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from test_golden_cpu import load

pytestmark = pytest.mark.gpu


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    port = s.getsockname()[1]
    s.close()
    return port


def _build(inp, a, b, shard, dev, kind):
    from velocycle_b200.preprocessing import make_velocity_metaparams

    sl = slice(a, b)
    return make_velocity_metaparams(
        inp["S"][sl], inp["U"][sl], inp["mu_nu"], inp["sd_nu"], inp["phixy_prior"][sl], inp["mu_nw"], inp["sd_nw"],
        batch_id=inp["batch_id"][sl], cond_id=inp["cond_id"][sl], Nb=int(inp["Nb"]), Nx=int(inp["Nx"]),
        count_factor=inp["cf"][sl], model_type=kind, device=dev, shard=shard)


def _fit(mp_, steps, graphed):
    from velocycle_b200 import ppl as pyro
    from velocycle_b200.ppl.infer import SVI, Trace_ELBO
    from velocycle_b200.ppl.optim import ClippedAdam
    from velocycle_b200.svi import GraphedSVI

    args = {"lr": 0.03, "lrd": 0.999, "betas": (0.8, 0.99)}
    pyro.clear_param_store()
    pyro.set_rng_seed(2024)
    if graphed:
        svi = GraphedSVI(mp_.model_fn, mp_.guide_fn, args, mp_)
        losses = [svi.step() for _ in range(steps)]
    else:
        svi = SVI(mp_.model_fn, mp_.guide_fn, ClippedAdam(args), Trace_ELBO())
        losses = [svi.step(mp_) for _ in range(steps)]
    return losses, {k: v.detach().clone() for k, v in pyro.get_param_store().named_parameters()}


def _worker(rank, world, port, tmp):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    torch.cuda.set_device(rank)
    dev = torch.device("cuda", rank)
    dist.init_process_group("nccl", rank=rank, world_size=world, device_id=dev)
    from velocycle_b200.sharding import ShardInfo, shard_cells

    z, inp = load("case_multi")
    Nc = inp["S"].shape[0]
    for kind, graphed in (("normal", False), ("lrmn", True)):
        shard = ShardInfo.make(Nc)
        a, b = shard_cells(Nc, rank, world)
        _, sharded = _fit(_build(inp, a, b, shard, dev, kind), 6, graphed)
        _, single = _fit(_build(inp, 0, Nc, None, dev, kind), 6, graphed)
        for name, ref in single.items():
            got = sharded[name]
            if name == "ϕxy_locs":
                ref = ref[a:b]
            finite = torch.isfinite(ref)
            err = float((got[finite] - ref[finite]).abs().max())
            assert err <= 5e-3, (kind, name, err)
    open(os.path.join(tmp, f"ok{rank}"), "w").write("ok")
    dist.destroy_process_group()


@pytest.mark.skipif(torch.cuda.device_count() < 2, reason="needs 2 GPUs")
def test_two_gpu_sharded_fit_matches_single_gpu(tmp_path):
    world = 2
    mp.spawn(_worker, args=(world, _free_port(), str(tmp_path)), nprocs=world, join=True)
    assert all((tmp_path / f"ok{r}").exists() for r in range(world))
