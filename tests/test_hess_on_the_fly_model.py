"""CPU: the arithmetic of csrc/spmm_hess.cu restated in numpy — the five per-node vectors lgnn_hess_stats_f32
writes and the per-edge reconstruction of the Hessian-sqrt columns lgnn_spmm_hess_f32 performs — against the
oracle's Hessian square root (oracle.hess_sqrt_rhs, pinned to the reference).  It checks the identity the kernel
rests on: v_c[k] = -(A_c P_k + S_c Q_k) for k != c, V_c on the diagonal, duplicates scaling A, S, V."""
import numpy as np
import pytest
import torch

from oracle import gcn_kfac_oracle as O


def stats_model(logits: np.ndarray, idx: np.ndarray, mode: str):
    """float32 restatement of hess_stats_kernel: returns P, Q, A, S, V as [n_nodes, C] arrays."""
    n, C = logits.shape
    P, Q, A, S, V = (np.zeros((n, C), np.float32) for _ in range(5))
    for node in idx:
        f = logits[node].astype(np.float32)
        e = np.exp(f - f.max(), dtype=np.float32)
        p = (e / e.sum(dtype=np.float32)).astype(np.float32)
        fbar = np.float32((p * f).sum(dtype=np.float32))
        fc = (f - fbar).astype(np.float32)
        s = np.sqrt(p, dtype=np.float32)
        a = (1 + 0.5 * fc).astype(np.float32) if mode == "reference" else np.ones(C, np.float32)
        v = (1 - p) * a - (p * fc if mode == "reference" else 0)
        P[node], Q[node] = p, (p * fc if mode == "reference" else 0)
        A[node] += s * a
        S[node] += s
        V[node] += s * v
    return P, Q, A, S, V


@pytest.mark.parametrize("C", [2, 3, 7, 40, 47, 64])
@pytest.mark.parametrize("mode", ["reference", "ggn"])
def test_rank_two_reconstruction_matches_the_hessian_sqrt(C, mode):
    rng = np.random.default_rng(C)
    n = 60
    logits = (3 * rng.standard_normal((n, C))).astype(np.float32)
    logits[7] = 0.0                       # uniform softmax
    logits[8, 0] = 25.0                   # a confidently classified node: 1 - p_0 ~ 1e-9
    idx = np.sort(rng.permutation(n)[:40])
    idx = np.concatenate([idx, idx[:3]])  # three nodes listed twice
    P, Q, A, S, V = stats_model(logits, idx, mode)
    want = np.zeros((n, C, C), np.float64)                         # [node, column c, class k]
    Vref = O.hess_sqrt_rhs(torch.from_numpy(logits[idx]).double(), mode).numpy()
    np.add.at(want, idx, Vref)
    got = -(A[:, :, None] * P[:, None, :] + S[:, :, None] * Q[:, None, :]).astype(np.float64)
    got[:, np.arange(C), np.arange(C)] = V
    scale = np.abs(want).max()
    assert np.abs(got - want).max() <= 1e-5 * scale          # fp32 statistics against the float64 oracle
    outside = np.setdiff1d(np.arange(n), idx)
    assert not got[outside].any()


def test_spmm_of_reconstructed_columns_matches_materialised_path():
    """The whole output-layer step on a small graph: sum_j a_ij v_{j,c} from the five vectors, column group by
    column group with zero padding columns, against SpMM of the materialised right-hand sides."""
    n, C, g = 80, 7, 4
    ei = O.synthetic_edges(n, 300, seed=2, directed=True)
    G = O.build_graph(ei, n)
    rng = np.random.default_rng(5)
    logits = rng.standard_normal((n, C)).astype(np.float32)
    idx = np.sort(rng.permutation(n)[:50])
    P, Q, A, S, V = stats_model(logits, idx, "reference")
    dense_t = np.zeros((n, n), np.float64)
    rows = np.repeat(np.arange(n), np.diff(G.t_rowptr))
    dense_t[rows, G.t_col] = G.t_val
    Vref = O.hess_sqrt_rhs(torch.from_numpy(logits[idx]).double(), "reference").numpy()
    full = np.zeros((n, C, C))
    full[idx] = Vref
    for c0 in range(0, C, g):
        ncols = min(g, C - c0)
        want = np.einsum("ij,jck->ick", dense_t, full[:, c0:c0 + ncols, :])
        acc = np.einsum("ij,jc,jk->ick", dense_t, A[:, c0:c0 + ncols].astype(np.float64), P.astype(np.float64))
        acc += np.einsum("ij,jc,jk->ick", dense_t, S[:, c0:c0 + ncols].astype(np.float64), Q.astype(np.float64))
        got = -acc
        d = dense_t @ V[:, c0:c0 + ncols].astype(np.float64)          # dacc of the kernel
        for c in range(ncols):
            got[:, c, c0 + c] = d[:, c]
        assert np.abs(got - want).max() <= 2e-6 * np.abs(want).max()
