"""TEST INFRASTRUCTURE — dev container only.  Golden of the reference's EPOCH LOOP.

Runs the reference's own `gnn.marglik_training.marglik_optimization` (gnn/marglik_training.py:41-329) — its Adam
step, its per-epoch `Laplace(...).fit` + `log_marginal_likelihood`, its validation forward, its model selection — on
the reference's dense `GCN` for a few epochs with dropout off (p = 0: the trajectory then depends on no RNG stream),
and stores per-epoch train loss, -marglik and validation loss plus the inputs and the initial weights:
    python oracle/make_golden_epochs.py      ->  tests/golden/epochs_*.npz
`tests/test_training_golden.py` holds `laplace_gnn_b200.training.marglik_training` to it (CPU double and GPU).
Nothing of the reference is copied; only its outputs are stored."""
from __future__ import annotations

import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from oracle import gcn_kfac_oracle as O  # noqa: E402
from oracle import ref_loader  # noqa: E402
from oracle.make_golden import dense_adj_from_edges  # noqa: E402

GOLDEN_DIR = os.path.join(ROOT, "tests", "golden")
CASES = {
    # name: n, undirected pairs, F, C, h, L, epochs, lr, weight decay
    "epochs_small_2l": dict(n=400, U=1200, F=24, C=5, h=16, L=2, epochs=6, lr=0.01, wd=5e-4),
    "epochs_cora_shape_2l": dict(n=2708, U=5278, F=1433, C=7, h=16, L=2, epochs=5, lr=0.01, wd=5e-4, bow=True),
}


def run_case(name, cfg, R, MT):
    rng = np.random.Generator(np.random.PCG64(23))
    n, F, C, h, L = cfg["n"], cfg["F"], cfg["C"], cfg["h"], cfg["L"]
    ei = O.synthetic_edges(n, cfg["U"], seed=5)
    x = (rng.random((n, F)) < 0.0127).astype(np.float32) if cfg.get("bow") else rng.standard_normal((n, F)).astype(np.float32)
    y_all = rng.integers(0, C, n).astype(np.int64)
    perm = rng.permutation(n)
    n_tr, n_va = int(0.6 * n), int(0.2 * n)
    tr, va = np.sort(perm[:n_tr]), np.sort(perm[n_tr:n_tr + n_va])
    torch.manual_seed(0)
    model = R.GCN(F, h, C, L, torch.from_numpy(x), dense_adj_from_edges(ei, n), dropout_p=0.0)
    W0 = [c.lin.weight.detach().numpy().copy() for c in model.convs]
    b0 = [c.lin.bias.detach().numpy().copy() for c in model.convs]
    args_dict = {"model_type": "gcn", "optimizer": "adam", "early_stop": False, "grad_norm": False,
                 "weight_decay_adj": 0.0, "momentum_adj": 0.0}
    best, losses, val_losses, neg_margliks = MT.marglik_optimization(
        model, torch.from_numpy(tr), torch.from_numpy(y_all[tr]), torch.from_numpy(va), torch.from_numpy(y_all[va]),
        y=torch.from_numpy(y_all), lr=cfg["lr"], weight_decay=cfg["wd"], n_epochs=cfg["epochs"], device="cpu",
        args_dict=args_dict)
    out = dict(edge_index=ei.astype(np.int32), n=np.int64(n), x=x if not cfg.get("bow") else np.zeros(0, np.float32),
               x_bits=np.packbits(x.astype(np.uint8), axis=1) if cfg.get("bow") else np.zeros(0, np.uint8),
               F=np.int64(F), train_idx=tr, train_y=y_all[tr], val_idx=va, val_y=y_all[va],
               lr=np.float64(cfg["lr"]), weight_decay=np.float64(cfg["wd"]), epochs=np.int64(cfg["epochs"]),
               losses=np.asarray(losses, np.float64), val_losses=np.asarray(val_losses, np.float64),
               neg_margliks=np.asarray(neg_margliks, np.float64),
               best_marglik_epoch=np.int64(best["marglik"]["epoch"]), best_valloss_epoch=np.int64(best["valloss"]["epoch"]))
    for l in range(L):
        out[f"W{l}"], out[f"b{l}"] = W0[l], b0[l]
        out[f"Wend{l}"] = model.convs[l].lin.weight.detach().numpy().copy()
    print(f"[golden] {name}: losses {np.round(losses, 5)} -marglik {np.round(neg_margliks, 3)} val {np.round(val_losses, 5)}")
    return out


def main():
    R = ref_loader.load()
    import gnn.marglik_training as MT
    os.makedirs(GOLDEN_DIR, exist_ok=True)
    for name, cfg in CASES.items():
        np.savez_compressed(os.path.join(GOLDEN_DIR, name + ".npz"), **run_case(name, cfg, R, MT))


if __name__ == "__main__":
    main()
