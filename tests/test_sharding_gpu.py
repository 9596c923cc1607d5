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


def _comm_worker(rank, world, port, tmp):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    torch.cuda.set_device(rank)
    dev = torch.device("cuda", rank)
    dist.init_process_group("nccl", rank=rank, world_size=world, device_id=dev)
    from velocycle_b200.sharding import PeerComm, ShardInfo

    n = 22_532
    comm = PeerComm.create(ShardInfo.make(1000), n, dev)
    assert comm is not None
    g = torch.Generator(device=dev).manual_seed(100 + rank)
    for it in range(7):  # odd and even epochs, both buffer sets reused several times
        x = torch.randn(n, generator=g, device=dev) * (10.0 ** (it - 3))
        parts = [torch.empty_like(x) for _ in range(world)]
        dist.all_gather(parts, x)  # (through NCCL: the independent path)
        exact = torch.stack(parts).double().sum(0)
        scale = torch.stack(parts).double().abs().sum(0)
        comm.allreduce_(x)
        assert bool(((x.double() - exact).abs() <= 1e-6 * scale + 1e-30).all()), it
        seq = parts[0].clone()
        for t in parts[1:]:
            seq += t
        assert torch.equal(x, seq), "the sum is taken in rank order 0..world-1"
        same = [torch.empty_like(x) for _ in range(world)]
        dist.all_gather(same, x)
        assert all(torch.equal(same[0], t) for t in same), "the one-shot sum must be bitwise identical on every rank"
    # inside a CUDA graph: replays advance the device-side epoch
    buf = torch.zeros(n, device=dev)
    src = torch.full((n,), float(rank + 1), device=dev)
    s = torch.cuda.Stream()
    s.wait_stream(torch.cuda.current_stream())
    with torch.cuda.stream(s):
        buf.copy_(src)
        comm.allreduce_(buf)
    torch.cuda.current_stream().wait_stream(s)
    torch.cuda.synchronize()
    graph = torch.cuda.CUDAGraph()
    with torch.cuda.graph(graph):
        buf.copy_(src)
        comm.allreduce_(buf)
    for _ in range(5):
        graph.replay()
        torch.cuda.synchronize()
        assert float(buf[0]) == world * (world + 1) / 2 and float(buf[-1]) == world * (world + 1) / 2
    open(os.path.join(tmp, f"comm{rank}"), "w").write("ok")
    dist.barrier()
    os._exit(0)  # (peer mappings and NCCL communicators: skip the teardown)


@pytest.mark.skipif(torch.cuda.device_count() < 2, reason="needs 2 GPUs")
def test_one_shot_peer_allreduce_matches_nccl(tmp_path):
    world = min(torch.cuda.device_count(), 8)
    mp.spawn(_comm_worker, args=(world, _free_port(), str(tmp_path)), nprocs=world, join=True)
    assert all((tmp_path / f"comm{r}").exists() for r in range(world))


@pytest.mark.skipif(torch.cuda.device_count() < 2, reason="needs 2 GPUs")
def test_two_gpu_sharded_fit_matches_single_gpu(tmp_path):
    world = 2
    mp.spawn(_worker, args=(world, _free_port(), str(tmp_path)), nprocs=world, join=True)
    assert all((tmp_path / f"ok{r}").exists() for r in range(world))
