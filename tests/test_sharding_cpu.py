"""CPU, world_size 2, gloo: the host-side logic of the multi-GPU path (cell ranges, the single flat all-reduce,
RNG slicing).  The CUDA op cannot run here, so the per-rank evaluation is the oracle's closed form; everything
around it -- ``shard_cells``, ``allreduce_flat_``, ``ShardedNormal``, the flat-buffer layout -- is product code."""
import os
import socket

import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    port = s.getsockname()[1]
    s.close()
    return port


def _flat_local(p):
    """What fused_elbo_grad lays out in ``gene_flat`` (fused.py), computed by the oracle on a cell slice."""
    from oracle.likelihood import analytic_gradients

    a = analytic_gradients(p)
    parts = [a["lp_S"], a["lp_U"], a["d_shape_inv"], a["d_logbeta"], a["d_gamma"], a["d_nu"].reshape(-1),
             a["d_dnu"].reshape(-1), a["d_nu_omega"].reshape(-1)]
    return torch.cat([t.reshape(-1) for t in parts]).float(), a


def _worker(rank, world, port, tmp):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    from velocycle_b200.sharding import ShardInfo, ShardedNormal, allreduce_flat_, shard_cells
    from velocycle_b200.synthetic import make_synthetic

    Nc, Ng = 301, 23
    d = make_synthetic(Nc, Ng, H=2, Hw=1, Nb=3, Nx=2, seed=4, device="cpu", sorted_batches=False)
    full = dict(S=d.S[:, :Ng], U=d.U[:, :Ng], phi=d.phi, cf=d.cf, batch_id=d.batch_id, cond_id=d.cond_id, nu=d.nu,
                dnu=d.dnu, shape_inv=d.shape_inv, logbeta=d.logbeta, gamma=torch.exp(d.loggamma) + 2.0,
                nu_omega=d.nu_omega)
    shard = ShardInfo.make(Nc)
    a, b = shard_cells(Nc, rank, world)
    assert shard.cell_offset == a and shard.Nc_local == b - a
    local = dict(full)
    for k in ("S", "U", "phi", "cf", "batch_id", "cond_id"):
        local[k] = full[k][a:b]
    flat, loc = _flat_local(local)
    allreduce_flat_(flat, shard)  # the step's single exchange
    ref_flat, ref = _flat_local(full)
    err = float((flat - ref_flat).abs().max() / ref_flat.abs().max())
    assert err < 1e-5, err
    # per-cell gradients never leave the rank and equal the corresponding slice of the full problem
    for k in ("d_phi", "d_cf", "d_omega"):
        assert torch.allclose(loc[k], ref[k][a:b], rtol=1e-9, atol=1e-9), k
    # RNG: every rank draws the global noise and keeps its slice -> identical to the single-process draw
    torch.manual_seed(77)
    gene_noise = torch.randn(Ng, 3)  # replicated draws come first and must agree across ranks
    locs = torch.arange(2 * Nc, dtype=torch.float32).reshape(Nc, 2)
    mine = ShardedNormal(locs[a:b], 1.0, shard, event_dims=1).rsample()
    torch.manual_seed(77)
    gene_noise2 = torch.randn(Ng, 3)
    whole = torch.distributions.Normal(locs, torch.ones(())).rsample()
    assert torch.equal(gene_noise, gene_noise2)
    assert torch.equal(mine, whole[a:b])
    lp = ShardedNormal(locs[a:b], 1.0, shard, event_dims=1).log_prob(mine)
    assert lp.shape == (b - a,)
    gathered = [torch.zeros_like(gene_noise) for _ in range(world)]
    dist.all_gather(gathered, gene_noise)
    assert all(torch.equal(g, gathered[0]) for g in gathered)
    # the fit drivers' early-exit decision is rank 0's on every rank (a rank leaving the loop alone would hang the others in
    # the step's all-reduce); PeerComm declines politely where there is no CUDA device / a single rank
    import collections

    from velocycle_b200.sharding import PeerComm
    from velocycle_b200.svi import agree_across_ranks

    MP = collections.namedtuple("MP", "shard device")
    assert agree_across_ranks(rank == 0, MP(shard, "cpu")) is True
    assert agree_across_ranks(rank != 0, MP(shard, "cpu")) is False
    assert agree_across_ranks(True, MP(None, "cpu")) is True
    assert PeerComm.create(ShardInfo(0, 1, 0, Nc), 128, "cpu") is None
    open(os.path.join(tmp, f"ok{rank}"), "w").write("ok")
    dist.destroy_process_group()


def test_two_rank_sharding_matches_single_process(tmp_path):
    world = 2
    mp.spawn(_worker, args=(world, _free_port(), str(tmp_path)), nprocs=world, join=True)
    assert all((tmp_path / f"ok{r}").exists() for r in range(world))


def test_shard_cells_partitions_exactly():
    from velocycle_b200.sharding import shard_cells

    for Nc in (0, 1, 7, 1000, 1_000_003):
        for world in (1, 2, 3, 8):
            edges = [shard_cells(Nc, r, world) for r in range(world)]
            assert edges[0][0] == 0 and edges[-1][1] == Nc
            assert all(edges[i][1] == edges[i + 1][0] for i in range(world - 1))
            sizes = [b - a for a, b in edges]
            assert max(sizes) - min(sizes) <= 1
