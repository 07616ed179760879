"""CPU: the ragged unit-compacted row layout as the host side sees it — ``ops.unit_row_slots`` and the
CPU double of ``unit_pack_ragged`` / ragged ``spmm_units`` (tests/fake_ops.py, what the gloo tests of the rows layout
run on) against the numpy restatement of the layout that the GPU tests hold the kernels to (helpers.unit_layout)."""
import numpy as np
import pytest
import torch

import fake_ops as F
from helpers import unit_layout
from laplace_gnn_b200 import ops
from oracle import gcn_kfac_oracle as O


def _case(n, g, h, density, seed):
    gen = torch.Generator().manual_seed(seed)
    act = torch.randn(n, h, generator=gen) * (torch.rand(n, h, generator=gen) < density)
    act[1] = 0
    act[2] = 1
    dense = torch.randn(n, g, h, generator=gen) * (act > 0)[:, None, :]
    return act, dense


@pytest.mark.parametrize("g,h", [(2, 32), (4, 64), (6, 96), (10, 256), (16, 128)])
@pytest.mark.parametrize("density", [0.0, 0.5, 1.0])
def test_row_slots_and_cpu_double_follow_the_layout(g, h, density):
    n, base = 37, 12
    act, dense = _case(n, g, h, density, seed=g * 10 + h)
    live = (act > 0).numpy()
    slots = ops.unit_row_slots(act, g)
    want_slots = []
    for r in range(n):
        hdr, slot = unit_layout(live[r], g)
        k_last = int(live[r][-32:].sum())
        want_slots.append(hdr[-1][1] + ((k_last + 1) // 2 * 2 if g % 4 else k_last))
    assert slots.tolist() == want_slots
    if g % 4:
        assert all(s % 2 == 0 for s in want_slots)
    first = torch.cumsum(slots, 0) - slots + base
    flat = torch.full(((base + int(slots.sum())) * g + 4,), -7.0)
    words = torch.zeros(n, h // 32, 2, dtype=torch.int32)
    F.unit_pack_ragged(dense.reshape(n, g * h), act, g, first, flat, words)
    W = words.numpy().view(np.uint32)
    for r in range(n):
        hdr, slot = unit_layout(live[r], g)
        for w, (mask, rel) in enumerate(hdr):
            assert W[r, w, 0] == mask and W[r, w, 1] == int(first[r]) + rel
        for u in np.nonzero(live[r])[0]:
            at = (int(first[r]) + slot[u]) * g
            assert torch.equal(flat[at: at + g], dense[r][:, u])
    us = ops.UnitSlab(n, g, h, flat, words, None, ragged=True)
    assert torch.equal(F._ragged_to_dense(us), dense.reshape(n, g * h))


def test_cpu_double_of_the_ragged_spmm_matches_the_dense_product():
    n, g, h = 300, 6, 64
    act, dense = _case(n, g, h, 0.5, seed=3)
    ei = O.synthetic_edges(n, 1500, seed=1)
    a = F.csr_from_edge_index(torch.from_numpy(ei), n)
    a.val = torch.rand(a.col.numel())
    slots = ops.unit_row_slots(act, g)
    order = torch.randperm(n, generator=torch.Generator().manual_seed(0))          # any row order is a valid plan
    first = torch.empty(n, dtype=torch.int64)
    first[order] = torch.cumsum(slots[order], 0) - slots[order]
    flat = torch.zeros(int(slots.sum()) * g + 4)
    words = torch.zeros(n, h // 32, 2, dtype=torch.int32)
    F.unit_pack_ragged(dense.reshape(n, g * h), act, g, first, flat, words)
    got = F.spmm_units(a, ops.UnitSlab(n, g, h, flat, words, None, ragged=True))
    assert torch.equal(got, F.spmm(a, dense.reshape(n, g * h)))


def test_collectives_are_chained_across_streams_only(monkeypatch):
    """dist._after_previous / _issued: a blocking collective waits for the previous one's event only when it is issued
    on another stream (the two lanes of the rows layout); same-stream sequences and asynchronous collectives are left
    alone.  Streams and events are stand-ins: the bookkeeping is what is tested."""
    import types
    import laplace_gnn_b200.dist as D
    log = []

    class Ev:
        def record(self, s):
            log.append(("record", s.cuda_stream))

    class St:
        def __init__(self, i):
            self.cuda_stream = i

        def wait_event(self, e):
            log.append(("wait", self.cuda_stream))

    cur = [St(1)]
    shim = types.SimpleNamespace(cuda=types.SimpleNamespace(current_stream=lambda d=None: cur[0], Event=Ev),
                                 Tensor=torch.Tensor)
    monkeypatch.setattr(D, "torch", shim)
    monkeypatch.setattr(D, "_LAST_COLLECTIVE", {})
    t = types.SimpleNamespace(is_cuda=True, device=None)
    pg = object()

    def collective(async_op=False):
        D._after_previous(pg, t, async_op)
        D._issued(pg, t, object() if async_op else None)

    collective()                      # first, stream 1
    collective()                      # same stream: implicit order
    cur[0] = St(2)
    collective()                      # lane 2: waits for stream 1's event
    collective(async_op=True)         # asynchronous: untouched
    cur[0] = St(1)
    collective()                      # back on stream 1: waits for stream 2's event
    assert log == [("record", 1), ("record", 1), ("wait", 2), ("record", 2), ("wait", 1), ("record", 1)]
    cpu = types.SimpleNamespace(is_cuda=False, device=None)
    D._after_previous(pg, cpu)
    D._issued(pg, cpu)                # CPU tensors (gloo): nothing recorded
    assert len(log) == 6
