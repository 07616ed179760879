"""CPU, world_size 2 and 3 over gloo: the row-partitioned multi-rank pass (halo all-gather of the
SpMM inputs, all-reduce of the factors and the loss) must reproduce the reference goldens in both
backward layouts, with the kernels replaced by the oracle-backed test double."""
import os
import socket
import sys

import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from conftest import ROOT, Golden


def _free_port():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


def _worker(rank, world, port, name, mode, out):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    sys.path.insert(0, ROOT)
    sys.path.insert(0, os.path.join(ROOT, "tests"))
    torch.set_num_threads(1)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        import fake_ops as F
        import laplace_gnn_b200 as L
        import laplace_gnn_b200.ops as ops
        for n in F.ALL:
            setattr(ops, n, getattr(F, n))
        from helpers import build_model, check_against_golden, loader_for
        g = Golden(name)
        model = build_model(g)
        la = L.Laplace(model, "classification", backend=L.B200GGN,
                       backend_kwargs={"process_group": dist.group.WORLD, "backward_parallel": mode,
                                       "rhs_tile_bytes": 40_000,        # several column groups
                                       "shard_eigh": mode == "columns"})  # eigh spread over the ranks / replicated
        la.fit(loader_for(g))
        ml = la.log_marginal_likelihood()
        check_against_golden(g, la.loss, la.H_facs.kfacs, ml)
        st = la.backend.last_stats
        assert st["world"] == world and 0.0 <= st["partition"].halo_fraction <= 1.0
        # every rank ends with bit-identical factors (same all-reduce result)
        flat = torch.cat([h.reshape(-1) for blk in la.H_facs.kfacs for h in blk])
        gathered = [torch.empty_like(flat) for _ in range(world)]
        dist.all_gather(gathered, flat)
        assert all(torch.equal(gathered[0], t) for t in gathered)
        out[rank] = float(ml)
    finally:
        dist.destroy_process_group()


@pytest.mark.parametrize("world", [2, 3])
@pytest.mark.parametrize("mode", ["rows", "columns"])
@pytest.mark.parametrize("name", ["tiny_directed_3l", "small_multibatch_2l"])
def test_partitioned_fit_matches_reference_golden(world, mode, name):
    port = _free_port()
    with mp.Manager() as m:
        out = m.dict()
        mp.spawn(_worker, args=(world, port, name, mode, out), nprocs=world, join=True)
        assert len(out) == world
        vals = list(out.values())
        assert all(v == vals[0] for v in vals)
        assert abs(vals[0] - Golden(name).marglik) <= 1e-3 * abs(Golden(name).marglik)


def _unit_worker(rank, world, port, out):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    sys.path.insert(0, ROOT)
    sys.path.insert(0, os.path.join(ROOT, "tests"))
    torch.set_num_threads(1)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        import fake_ops as F
        import laplace_gnn_b200 as L
        import laplace_gnn_b200.ops as ops
        for n in F.ALL:
            setattr(ops, n, getattr(F, n))
        from test_host_logic import _synthetic_model
        model, idx, y = _synthetic_model(500, 2000, 10, 32, 10, 3)
        ref = L.B200GGN(model, "classification", unit_slabs=False).kron(idx, y, N=len(y))
        for mode in ("columns", "rows"):
            be = L.B200GGN(model, "classification", process_group=dist.group.WORLD, backward_parallel=mode,
                           unit_min_width=0)
            loss, kron = be.kron(idx, y, N=len(y))
            # column-parallel backward: every rank holds the whole graph -> its 5 columns travel as one
            # zero-padded group through the unit-compacted slabs; the row layout packs each rank's rows back to
            # back (ragged rows, absolute slots in the all-gathered headers) and exchanges those
            assert be.last_stats["unit_slabs"] > 0
            assert abs(float(loss) - float(ref[0])) <= 1e-5 * abs(float(ref[0]))
            for fa, fb in zip(kron.kfacs, ref[1].kfacs):
                for a, b in zip(fa, fb):
                    assert float((a - b).abs().max()) <= 1e-5 * float(b.abs().max())
        # rows layout: groups of 12 (slots of 16 bytes) instead of 10 (even-slot layout), several narrow groups,
        # and the dense exchange with the switch off
        for kw, units in (({"unit_even_groups": False}, True), ({"rhs_tile_bytes": 600_000}, True),
                          ({"unit_rows": False}, False)):
            be = L.B200GGN(model, "classification", process_group=dist.group.WORLD, backward_parallel="rows",
                           unit_min_width=0, **kw)
            loss, kron = be.kron(idx, y, N=len(y))
            assert (be.last_stats["unit_slabs"] > 0) == units, (kw, be.last_stats)
            for fa, fb in zip(kron.kfacs, ref[1].kfacs):
                for a, b in zip(fa, fb):
                    assert float((a - b).abs().max()) <= 1e-5 * float(b.abs().max()), kw
        # unit_even_groups: the rank's 5 columns travel as a group of 6 instead of 8
        be = L.B200GGN(model, "classification", process_group=dist.group.WORLD, backward_parallel="columns",
                       unit_min_width=0, unit_even_groups=True)
        loss, kron = be.kron(idx, y, N=len(y))
        assert be.last_stats["group"] == 6 and be.last_stats["unit_slabs"] > 0
        for fa, fb in zip(kron.kfacs, ref[1].kfacs):
            for a, b in zip(fa, fb):
                assert float((a - b).abs().max()) <= 1e-5 * float(b.abs().max())
        out[rank] = True
    finally:
        dist.destroy_process_group()


def test_column_parallel_backward_uses_unit_compacted_slabs():
    port = _free_port()
    with mp.Manager() as m:
        out = m.dict()
        mp.spawn(_unit_worker, args=(2, port, out), nprocs=2, join=True)
        assert len(out) == 2


def _lanes_worker(rank, world, port, out):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    sys.path.insert(0, ROOT)
    sys.path.insert(0, os.path.join(ROOT, "tests"))
    torch.set_num_threads(1)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        import fake_ops as F
        import laplace_gnn_b200 as L
        import laplace_gnn_b200.ops as ops
        from laplace_gnn_b200 import curvature
        for n in F.ALL:
            setattr(ops, n, getattr(F, n))
        curvature._B200KFAC._lanes_on_cpu = True          # two column groups interleaved, as on the device
        from test_host_logic import _synthetic_model
        model, idx, y = _synthetic_model(1200, 5000, 12, 64, 11, 3)
        ref = L.B200GGN(model, "classification", unit_slabs=False).kron(idx, y, N=len(y))
        seen = []
        # (switches, group width expected, unit SpMMs expected)
        for kw, group, units in (({"unit_min_width": 0}, 6, True),                              # 6 + 6(5): even slots
                                 ({"unit_min_width": 0, "unit_even_groups": False}, 8, True),   # 8 + 4(3): 16-byte slots
                                 ({"unit_min_width": 0, "rhs_tile_bytes": 2_500_000}, 2, True),  # six narrow groups
                                 ({"unit_min_width": 0, "hess_sqrt": "ggn"}, 6, True),
                                 ({"unit_rows": False}, 6, False),
                                 # output layer from all-gathered softmax statistics (opt-in), both modes, narrow groups
                                 ({"unit_min_width": 0, "rows_hess_stats": True}, 6, True),
                                 ({"unit_min_width": 0, "rows_hess_stats": True, "hess_sqrt": "ggn",
                                   "unit_even_groups": False}, 8, True),
                                 ({"unit_min_width": 0, "rows_hess_stats": True, "rhs_tile_bytes": 2_500_000}, 2, True),
                                 ({"rows_hess_stats": True, "unit_rows": False}, 6, False)):    # needs unit_rows: dense
            ref_kw = ref if kw.get("hess_sqrt") != "ggn" else \
                L.B200GGN(model, "classification", unit_slabs=False, hess_sqrt="ggn").kron(idx, y, N=len(y))
            be = L.B200GGN(model, "classification", process_group=dist.group.WORLD, backward_parallel="rows",
                           overlap=True, **kw)
            loss, kron = be.kron(idx, y, N=len(y))
            st = be.last_stats
            assert st["group"] == group and (st["unit_slabs"] > 0) == units, (kw, st)
            assert abs(float(loss) - float(ref_kw[0])) <= 1e-5 * abs(float(ref_kw[0]))
            for fa, fb in zip(kron.kfacs, ref_kw[1].kfacs):
                for a, b in zip(fa, fb):
                    assert float((a - b).abs().max()) <= 1e-5 * float(b.abs().max()), kw
            seen.append(st["n_groups"])
        out[rank] = seen
    finally:
        dist.destroy_process_group()


@pytest.mark.parametrize("world", [2, 3])
def test_rows_layout_with_two_lanes_and_ragged_unit_rows(world):
    """The two-lane interleaving of the rows layout (one group's all-gather under the other's SpMM on the device)
    walked on the CPU double: plans for both slot layouts, narrow groups, both Hessian square roots, the dense
    exchange and the materialised output layer as A/B — same factors as the single-process dense pass."""
    port = _free_port()
    with mp.Manager() as m:
        out = m.dict()
        mp.spawn(_lanes_worker, args=(world, port, out), nprocs=world, join=True)
        assert len(out) == world and len({tuple(v) for v in out.values()}) == 1


def _tiny_worker(rank, world, port, out):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    sys.path.insert(0, ROOT)
    sys.path.insert(0, os.path.join(ROOT, "tests"))
    torch.set_num_threads(1)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        import fake_ops as F
        import laplace_gnn_b200 as L
        import laplace_gnn_b200.ops as ops
        for n in F.ALL:
            setattr(ops, n, getattr(F, n))
        from test_host_logic import _synthetic_model
        empties = 0
        for nodes, pairs in ((3, 2), (7, 5)):
            model, idx, y = _synthetic_model(nodes, pairs, 4, 32, 3, 3)
            ref = L.B200GGN(model, "classification", unit_slabs=False).kron(idx, y, N=len(y))
            for mode in ("rows", "columns"):
                be = L.B200GGN(model, "classification", process_group=dist.group.WORLD, backward_parallel=mode,
                               unit_min_width=0, rows_hess_stats=(nodes == 7))
                loss, kron = be.kron(idx, y, N=len(y))
                b = be.last_stats["partition"].bounds
                empties += sum(b[r + 1] == b[r] for r in range(world))
                if mode == "rows":              # (columns: 3 classes over 4 ranks leave one rank without a column)
                    assert be.last_stats["unit_slabs"] > 0
                assert abs(float(loss) - float(ref[0])) <= 1e-5 * abs(float(ref[0]))
                for fa, fb in zip(kron.kfacs, ref[1].kfacs):
                    for a, c in zip(fa, fb):
                        assert float((a - c).abs().max()) <= 1e-5 * float(c.abs().max()) + 1e-30, (nodes, mode)
        assert empties > 0                      # 3 nodes over 4 ranks: at least one rank owns no row
        out[rank] = True
    finally:
        dist.destroy_process_group()


def test_more_ranks_than_rows():
    """Ranks that own no row at all (3 nodes over 4 ranks) in both layouts, unit-compacted slabs on: empty plans,
    empty packs and empty slots of the all-gathers."""
    port = _free_port()
    with mp.Manager() as m:
        out = m.dict()
        mp.spawn(_tiny_worker, args=(4, port, out), nprocs=4, join=True)
        assert len(out) == 4


def test_column_share_covers_all_columns_once():
    from laplace_gnn_b200.dist import column_share
    for C in (1, 3, 7, 40, 47):
        for world in (1, 2, 3, 8, 64):
            seen = []
            for r in range(world):
                s, c = column_share(C, r, world)
                seen += list(range(s, s + c))
            assert seen == list(range(C))
            sizes = [column_share(C, r, world)[1] for r in range(world)]
            assert max(sizes) - min(sizes) <= 1


def _eigh_worker(rank, world, port, out):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    sys.path.insert(0, ROOT)
    torch.set_num_threads(1)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        from laplace_gnn_b200.kron import Kron
        gen = torch.Generator().manual_seed(7)
        blocks = []
        for n_out, n_in in ((16, 9), (16, 16), (5, 16)):
            a, b = torch.randn(40, n_out, generator=gen), torch.randn(40, n_in, generator=gen)
            G, A = a.T @ a, b.T @ b
            dup = G.clone()
            dup._dup_of = G
            blocks += [[G, A], [dup]]
        k = Kron(blocks)
        ref = k.decompose()
        got = k.decompose(process_group=dist.group.WORLD)
        for vr, vg, qr, qg in zip(ref.eigenvalues, got.eigenvalues, ref.eigenvectors, got.eigenvectors):
            for a, b in zip(vr + qr, vg + qg):
                assert torch.equal(a, b)             # same LAPACK call on the same input, whoever ran it
        assert float(got.logdet()) == float(ref.logdet())
        out[rank] = float((got + torch.tensor(0.5)).logdet())
    finally:
        dist.destroy_process_group()


@pytest.mark.parametrize("world", [2, 3])
def test_sharded_eigh_equals_replicated(world):
    port = _free_port()
    with mp.Manager() as m:
        out = m.dict()
        mp.spawn(_eigh_worker, args=(world, port, out), nprocs=world, join=True)
        vals = list(out.values())
        assert len(vals) == world and all(v == vals[0] for v in vals)


def test_eigh_assignment_balances_cubic_cost():
    from laplace_gnn_b200.kron import eigh_assignment
    sizes = [256, 100, 256, 256, 47, 256]                 # G_1, A_1, G_2, A_2, G_3, A_3 of the products shape
    own = eigh_assignment(sizes, 8)
    assert sorted(own) == [0, 1, 2, 3, 4, 5]              # six factors, six ranks, nobody holds two
    own = eigh_assignment(sizes, 2)
    load = [sum(s ** 3 for s, o in zip(sizes, own) if o == r) for r in range(2)]
    assert abs(load[0] - load[1]) <= 100 ** 3 + 47 ** 3
    assert eigh_assignment(sizes, 1) == [0] * 6 and eigh_assignment([], 4) == []


def _halo_worker(rank, world, port, out):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    sys.path.insert(0, ROOT)
    sys.path.insert(0, os.path.join(ROOT, "tests"))
    torch.set_num_threads(1)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        import numpy as np
        import fake_ops as F
        import laplace_gnn_b200 as L
        import laplace_gnn_b200.ops as ops
        for n in F.ALL:
            setattr(ops, n, getattr(F, n))
        # a graph with locality: node i is linked to i+1 .. i+3 (directed) -> a contiguous row block reads only a
        # few rows of its neighbours' blocks
        n, C = 240, 5
        src = np.repeat(np.arange(n - 3), 3)
        dst = src + np.tile(np.arange(1, 4), n - 3)
        graph = L.Graph.from_edge_index(torch.from_numpy(np.stack([src, dst]).astype(np.int64)), n)
        gen = torch.Generator().manual_seed(3)
        torch.manual_seed(3)
        model = L.SparseGCN(7, 16, C, 3, torch.randn(n, 7, generator=gen), graph)
        idx = torch.randperm(n, generator=gen)[:150].sort().values
        y = torch.randint(0, C, (150,), generator=gen)
        ref_loss, ref = L.B200GGN(model, "classification").kron(idx, y, N=150)
        for mode in ("rows", "columns"):
            be = L.B200GGN(model, "classification", process_group=dist.group.WORLD, backward_parallel=mode,
                           sparse_halo=True, rhs_tile_bytes=60_000)
            loss, kron = be.kron(idx, y, N=150)
            part = be.last_stats["partition"]
            assert part.sparse_halo and part.halo_fraction < 0.1
            plan = part.halo_plan()
            assert sum(plan["recv_counts"]) <= 6 and sum(plan["send_counts"]) <= 6      # 3 rows per neighbouring block
            assert abs(float(loss) - float(ref_loss)) <= 1e-5 * abs(float(ref_loss))
            for fa, fb in zip(kron.kfacs, ref.kfacs):
                for a, b in zip(fa, fb):
                    assert float((a - b).abs().max()) <= 1e-5 * float(b.abs().max())
        # the switch off: the same partition all-gathers whole slabs
        be = L.B200GGN(model, "classification", process_group=dist.group.WORLD, backward_parallel="rows")
        be.kron(idx, y, N=150)
        assert not be.last_stats["partition"].sparse_halo
        out[rank] = True
    finally:
        dist.destroy_process_group()


@pytest.mark.parametrize("world", [2, 3])
def test_sparse_halo_exchange_matches_the_single_rank_pass(world):
    port = _free_port()
    with mp.Manager() as m:
        out = m.dict()
        mp.spawn(_halo_worker, args=(world, port, out), nprocs=world, join=True)
        assert len(out) == world


def _ingest_worker(rank, world, port, out):
    """dist.ingest_edge_index / ingest_rows / row_block: the sharded copy + all-gather gives the graph of the
    unsharded list bit for bit, and a model holding only its node block of X (x_rows) fits to the same factors."""
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    sys.path.insert(0, ROOT)
    sys.path.insert(0, os.path.join(ROOT, "tests"))
    torch.set_num_threads(1)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        import fake_ops as F
        import laplace_gnn_b200 as L
        import laplace_gnn_b200.ops as ops
        from laplace_gnn_b200 import dist as D
        for n in F.ALL:
            setattr(ops, n, getattr(F, n))
        from helpers import build_model
        g = Golden("tiny_directed_3l")
        cpu = torch.device("cpu")
        host_ei = torch.from_numpy(g.edge_index.astype("int64"))              # E = 160: not a multiple of 3
        ei = D.ingest_edge_index(host_ei, cpu, dist.group.WORLD)
        assert ei.shape[0] == 2 and ei.shape[1] >= host_ei.shape[1]
        assert set(map(tuple, ei.t().tolist())) == set(map(tuple, host_ei.t().tolist()))
        whole, shard = L.Graph.from_edge_index(host_ei, g.n), L.Graph.from_edge_index(ei, g.n)
        for a, b in ((whole.ahat, shard.ahat), (whole.ahat_t, shard.ahat_t)):
            assert torch.equal(a.rowptr, b.rowptr) and torch.equal(a.col, b.col) and torch.equal(a.val, b.val)
        assert D.ingest_edge_index(host_ei[:, :1], cpu, dist.group.WORLD).shape == (2, world)   # fewer edges than ranks
        # a model that holds this rank's node block of X only
        ref = build_model(g)
        lo, hi = D.row_block(shard, dist.group.WORLD)
        x_loc = D.ingest_rows(torch.from_numpy(g.x), lo, hi, cpu)
        model = L.SparseGCN(g.F, g.h, g.C, g.L, x_loc, shard, x_rows=(lo, hi))
        model.load_state_dict(ref.state_dict())
        idx, y = torch.from_numpy(g.idx), torch.from_numpy(g.y)
        kw = {"process_group": dist.group.WORLD, "backward_parallel": "columns"}
        la = L.Laplace(model, "classification", backend=L.B200GGN, backend_kwargs=kw)
        la.fit(L.TensorBatchLoader(idx, y))
        la_ref = L.Laplace(ref, "classification", backend=L.B200GGN, backend_kwargs=kw)
        la_ref.fit(L.TensorBatchLoader(idx, y))
        for fa, fb in zip(la.H_facs.kfacs, la_ref.H_facs.kfacs):
            for a, b in zip(fa, fb):
                assert torch.equal(a, b)
        assert float(la.log_marginal_likelihood()) == float(la_ref.log_marginal_likelihood())
        with pytest.raises(RuntimeError):
            model(idx)                                                        # the training forward needs all rows
        with pytest.raises(ValueError):
            L.B200GGN(model, "classification").kron(idx, y, N=len(y))         # ... and so does a single-device fit
        out[rank] = 1
    finally:
        dist.destroy_process_group()


@pytest.mark.parametrize("world", [2, 3])
def test_sharded_ingest_gives_the_same_graph_and_fit(world):
    with mp.Manager() as m:
        out = m.dict()
        mp.spawn(_ingest_worker, args=(world, _free_port(), out), nprocs=world, join=True)
        assert len(out) == world
