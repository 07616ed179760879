"""GPU, needs >= 2 devices (skipped on a 1-GPU box): the NCCL row-partitioned pass against the
single-GPU pass on the same arxiv-like graph — both backward layouts, with and without the
two-lane overlap.  One process per GPU; never more ranks than GPUs."""
import datetime
import os
import socket
import sys

import pytest
import torch

from conftest import ROOT

pytestmark = pytest.mark.gpu


def _free_port():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


def _guard(seconds: int = 240):
    """A rank that fails alone leaves its peers waiting in NCCL: dump every thread's stack and exit instead of
    sitting out the watchdog's default 10 minutes."""
    import faulthandler
    faulthandler.enable()
    faulthandler.dump_traceback_later(seconds, exit=True)


def _worker(rank, world, port, out):
    import torch.distributed as dist
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    sys.path.insert(0, ROOT)
    torch.cuda.set_device(rank)
    dev = torch.device("cuda", rank)
    _guard()
    dist.init_process_group("nccl", rank=rank, world_size=world, device_id=dev, timeout=datetime.timedelta(seconds=90))
    try:
        import laplace_gnn_b200 as L
        n, u, f, c, h, layers = 40_000, 300_000, 64, 10, 128, 3
        gen = torch.Generator(device=dev).manual_seed(0)
        src = torch.randint(0, n, (u,), device=dev, generator=gen)
        dst = torch.randint(0, n, (u,), device=dev, generator=gen)
        ei = torch.stack([torch.cat([src, dst]), torch.cat([dst, src])])
        X = torch.randn(n, f, device=dev, generator=gen)
        idx = torch.randperm(n, device=dev, generator=gen)[: int(0.6 * n)].sort().values
        y = torch.randint(0, c, (idx.numel(),), device=dev, generator=gen)
        torch.manual_seed(0)
        model = L.SparseGCN(f, h, c, layers, X, L.Graph.from_edge_index(ei, n)).to(dev)

        def fit(**kw):
            la = L.Laplace(model, "classification", backend=L.B200GGN, backend_kwargs=kw)
            la.fit(L.TensorBatchLoader(idx, y))
            return la, float(la.log_marginal_likelihood())

        ref, ref_ml = fit()
        # (switches, unit-compacted slabs expected).  Rows layout: the ranks exchange ragged unit-compacted rows —
        # one group of 10 x 128 (even-slot layout), two lanes of 6 x 128 and of 8 x 128 (16-byte slots) — or dense
        # slabs where the group is too narrow for the HBM budget given / the switch is off
        for kw, units in (({"backward_parallel": "rows", "overlap": False}, True),
                          ({"backward_parallel": "rows", "overlap": True, "rhs_tile_bytes": 64 << 20}, False),
                          ({"backward_parallel": "rows", "overlap": True, "unit_min_width": 0}, True),
                          ({"backward_parallel": "rows", "overlap": True, "unit_min_width": 0,
                            "unit_even_groups": False}, True),
                          ({"backward_parallel": "rows", "overlap": True, "unit_rows": False}, False),
                          ({"backward_parallel": "columns", "unit_min_width": 0}, True)):   # 5 columns per rank as 6 x 128
            la, ml = fit(process_group=dist.group.WORLD, **kw)
            for blk, rblk in zip(la.H_facs.kfacs, ref.H_facs.kfacs):
                for a, b in zip(blk, rblk):
                    err = float((a - b).abs().max() / b.abs().max())
                    assert err <= 2e-5, (kw, err)
            assert abs(ml - ref_ml) <= 1e-5 * abs(ref_ml), (kw, ml, ref_ml)
            assert abs(float(la.loss) - float(ref.loss)) <= 1e-6 * abs(float(ref.loss))
            assert (la.backend.last_stats["unit_slabs"] > 0) == units, (kw, la.backend.last_stats)
        assert ref.backend.last_stats["unit_slabs"] > 0
        out[rank] = ref_ml
    finally:
        dist.destroy_process_group()


@pytest.mark.skipif(torch.cuda.device_count() < 2, reason="needs at least 2 GPUs")
def test_nccl_partitioned_fit_matches_single_gpu():
    import torch.multiprocessing as mp
    world = min(torch.cuda.device_count(), 4)
    with mp.Manager() as m:
        out = m.dict()
        mp.spawn(_worker, args=(world, _free_port(), out), nprocs=world, join=True)
        assert len(out) == world


def _lab_worker(rank, world, port, out):
    """The multi-GPU switches that have not run over NCCL yet: eigendecompositions spread over the
    ranks, even column groups in the column-parallel backward, halo-only exchange on a graph with locality."""
    import numpy as np
    import torch.distributed as dist
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    sys.path.insert(0, ROOT)
    torch.cuda.set_device(rank)
    dev = torch.device("cuda", rank)
    _guard()
    dist.init_process_group("nccl", rank=rank, world_size=world, device_id=dev, timeout=datetime.timedelta(seconds=90))
    try:
        import laplace_gnn_b200 as L

        def fit(model, idx, y, **kw):
            la = L.Laplace(model, "classification", backend=L.B200GGN, backend_kwargs=kw)
            la.fit(L.TensorBatchLoader(idx, y))
            return la, float(la.log_marginal_likelihood())

        def same(la, ref, tag):
            for blk, rblk in zip(la.H_facs.kfacs, ref.H_facs.kfacs):
                for a, b in zip(blk, rblk):
                    assert float((a - b).abs().max() / b.abs().max()) <= 2e-5, tag

        # (1) uniform graph, 10 classes on 2 ranks = 5 columns each: even groups (6) + sharded eigh
        n, u, f, c, h, layers = 40_000, 300_000, 64, 10, 128, 3
        gen = torch.Generator(device=dev).manual_seed(0)
        src = torch.randint(0, n, (u,), device=dev, generator=gen)
        dst = torch.randint(0, n, (u,), device=dev, generator=gen)
        ei = torch.stack([torch.cat([src, dst]), torch.cat([dst, src])])
        X = torch.randn(n, f, device=dev, generator=gen)
        idx = torch.randperm(n, device=dev, generator=gen)[: int(0.6 * n)].sort().values
        y = torch.randint(0, c, (idx.numel(),), device=dev, generator=gen)
        torch.manual_seed(0)
        model = L.SparseGCN(f, h, c, layers, X, L.Graph.from_edge_index(ei, n)).to(dev)
        ref, ref_ml = fit(model, idx, y)
        la, ml = fit(model, idx, y, process_group=dist.group.WORLD, backward_parallel="columns", shard_eigh=True,
                     unit_even_groups=True, unit_min_width=0)
        same(la, ref, "even groups + sharded eigh")
        assert abs(ml - ref_ml) <= 1e-5 * abs(ref_ml)
        if world == 2:
            assert la.backend.last_stats["group"] == 6
        # every rank holds bit-identical eigenpairs
        flat = torch.cat([t.reshape(-1) for blk in la.H.eigenvalues for t in blk])
        gathered = [torch.empty_like(flat) for _ in range(world)]
        dist.all_gather(gathered, flat)
        assert all(torch.equal(gathered[0], t) for t in gathered)
        # (2) banded graph: halo-only exchange in both layouts
        n2 = 60_000
        s2 = torch.arange(n2 - 3, device=dev).repeat_interleave(3)
        d2 = s2 + torch.arange(1, 4, device=dev).repeat(n2 - 3)
        X2 = torch.randn(n2, f, device=dev, generator=gen)
        torch.manual_seed(1)
        model2 = L.SparseGCN(f, h, c, layers, X2, L.Graph.from_edge_index(torch.stack([s2, d2]), n2)).to(dev)
        idx2 = torch.randperm(n2, device=dev, generator=gen)[: int(0.6 * n2)].sort().values
        y2 = torch.randint(0, c, (idx2.numel(),), device=dev, generator=gen)
        ref2, ref2_ml = fit(model2, idx2, y2)
        for mode in ("rows", "columns"):
            la2, ml2 = fit(model2, idx2, y2, process_group=dist.group.WORLD, backward_parallel=mode, sparse_halo=True)
            assert la2.backend.last_stats["partition"].sparse_halo
            same(la2, ref2, f"sparse halo {mode}")
            assert abs(ml2 - ref2_ml) <= 1e-5 * abs(ref2_ml)
        out[rank] = ref_ml
    finally:
        dist.destroy_process_group()


@pytest.mark.skipif(torch.cuda.device_count() < 2, reason="needs at least 2 GPUs")
def test_nccl_lab_switches_match_single_gpu():
    import torch.multiprocessing as mp
    world = min(torch.cuda.device_count(), 4)
    with mp.Manager() as m:
        out = m.dict()
        mp.spawn(_lab_worker, args=(world, _free_port(), out), nprocs=world, join=True)
        assert len(out) == world


@pytest.mark.skipif(torch.cuda.device_count() < 2, reason="needs at least 2 GPUs")
def test_model_on_another_device_than_the_current_one():
    """The backend, the GCN layers and the graph build run on the device of their tensors whichever device is
    current in the calling thread; a raw ops.* call with a foreign device's tensor raises instead of launching on the
    wrong stream (_lib.stream() is the current device's stream)."""
    sys.path.insert(0, ROOT)
    import laplace_gnn_b200 as L
    from laplace_gnn_b200 import ops
    from laplace_gnn_b200._lib import LgnnError
    torch.cuda.set_device(0)
    n, u, f, c, h, layers = 5_000, 30_000, 16, 5, 64, 3
    gen = torch.Generator().manual_seed(0)
    ei = torch.randint(0, n, (2, u), generator=gen)
    X = torch.randn(n, f, generator=gen)
    idx = torch.randperm(n, generator=gen)[: int(0.6 * n)].sort().values
    y = torch.randint(0, c, (idx.numel(),), generator=gen)
    results = []
    for d in ("cuda:0", "cuda:1"):
        graph = L.Graph.from_edge_index(ei.to(d), n)                     # current device stays 0
        torch.manual_seed(0)
        model = L.SparseGCN(f, h, c, layers, X.to(d), graph).to(d)
        out = model(idx.to(d))
        out.sum().backward()
        la = L.Laplace(model, "classification", backend=L.B200GGN)
        la.fit(L.TensorBatchLoader(idx.to(d), y.to(d)))
        results.append((float(la.log_marginal_likelihood()), [[t.cpu() for t in blk] for blk in la.H_facs.kfacs],
                        model.convs[0].lin.weight.grad.cpu()))
        assert torch.cuda.current_device() == 0
    assert results[0][0] == results[1][0]
    for ba, bb in zip(results[0][1], results[1][1]):
        for a, b in zip(ba, bb):
            assert torch.equal(a, b)
    assert torch.equal(results[0][2], results[1][2])
    with pytest.raises(LgnnError, match="current CUDA device"):
        ops.spmm(graph.ahat, torch.randn(n, 8, device="cuda:1"))        # graph lives on cuda:1, current device is 0
