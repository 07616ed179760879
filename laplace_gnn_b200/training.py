"""Epoch loop of the reference's experiment driver on the sparse GCN
(gnn/marglik_training.py:159-329): per epoch one Adam step on the
cross-entropy of the train nodes (dropout active, forward / backward through ``GCNConvFunction``),
then ``Laplace(...).fit`` + ``log_marginal_likelihood`` on the B200 backend, then a validation
forward; model selection by marginal likelihood and by validation loss with patience.

SURVEY §8(f) row 1: this is the caller of the hot path.  What changes against the reference loop:
one full-graph forward per fit instead of three (the backend computes loss and factors from the
same forward, the n_outputs probe forward is gone), and the weight-independent input factor
A_0 = X^T X is computed once and reused across epochs (``cache_input_factor``).

With ``edge_scores`` (structure.EdgeScores) the loop also runs the reference's structure-learning
block (:194-224): every ``marglik_frequency`` epochs after the burn-in, ``n_hypersteps`` SGD steps
on the straight-through edge scores with the gradient of the negative log marginal likelihood,
the graph re-binarised and rebuilt after each step."""
from __future__ import annotations

from copy import deepcopy
from dataclasses import dataclass, field

import torch

from .data import TensorBatchLoader


@dataclass
class MarglikTrainingResult:
    losses: list = field(default_factory=list)          # training CE per epoch
    neg_margliks: list = field(default_factory=list)    # -log marginal likelihood per epoch
    val_losses: list = field(default_factory=list)
    val_accs: list = field(default_factory=list)
    best_marglik_epoch: int = 0
    best_valloss_epoch: int = 0
    best_marglik_state: dict | None = None
    best_valloss_state: dict | None = None
    stopped_epoch: int = 0
    n_edges: list = field(default_factory=list)         # (epoch, off-diagonal edges after the structure steps)


def marglik_training(model, train_idx, train_y, val_idx, val_y, n_epochs: int = 200, lr: float = 0.01,
                     weight_decay: float = 5e-4, patience: int = 50, early_stop: bool = False,
                     hessian_structure: str = "kron", prior_precision: float = 1.0,
                     backend_kwargs: dict | None = None, laplace=None, batch_size: int | None = None,
                     seed: int = 0, edge_scores=None, lr_adj: float = 0.1, n_hypersteps: int = 20,
                     n_epochs_burnin: int = 40, marglik_frequency: int = 20, n_hyper_stop: int | None = None,
                     momentum_adj: float = 0.0, weight_decay_adj: float = 0.0,
                     grad_norm: bool = False) -> MarglikTrainingResult:
    """``laplace`` is the factory to use (the reference's ``laplace.Laplace`` when that package is
    importable, default: the stand-in of this package); everything else mirrors the reference's
    argument meaning (lr / weight_decay: ``marglik_training.py:104-123``; PATIENCE: ``:40``)."""
    if laplace is None:
        from .kron import Laplace as laplace
    from .curvature import B200GGN
    kw = {"cache_input_factor": True}
    kw.update(backend_kwargs or {})
    opt = torch.optim.Adam([p for n, p in model.named_parameters() if "adj" not in n], lr=lr,
                           weight_decay=weight_decay)
    crit = torch.nn.CrossEntropyLoss()
    loader = TensorBatchLoader(train_idx, train_y, batch_size)
    res = MarglikTrainingResult()
    best_nml, best_val = float("inf"), float("inf")
    ml_pat = val_pat = 0
    torch.manual_seed(seed)
    backend_cache: dict = {}
    adj_opt = None
    if edge_scores is not None:                           # marglik_training.py:95-104
        adj_opt = torch.optim.SGD([edge_scores.score], lr=lr_adj, weight_decay=weight_decay_adj,
                                  momentum=momentum_adj)
    n_hyper_stop = n_epochs + 1 if n_hyper_stop is None else n_hyper_stop
    for epoch in range(1, n_epochs + 1):
        model.train()
        epoch_loss = 0.0
        for idx_b, y_b in loader:                         # marglik_training.py:163-181
            opt.zero_grad()
            loss = crit(model(idx_b), y_b)
            loss.backward()
            opt.step()
            epoch_loss += float(loss.detach())
        res.losses.append(epoch_loss)

        if adj_opt is not None and epoch < n_hyper_stop and epoch % marglik_frequency == 0 \
                and epoch >= n_epochs_burnin:             # :194-224
            model.eval()
            for _ in range(n_hypersteps):
                edge_scores.neg_marglik_step(model, train_idx, train_y, adj_opt, prior_precision,
                                             kw.get("hess_sqrt", "reference"), grad_norm, batch_size)
            res.n_edges.append((epoch, int(edge_scores.active.sum())))

        la = laplace(model, "classification", subset_of_weights="all", hessian_structure=hessian_structure,
                     prior_precision=prior_precision, backend=B200GGN,
                     backend_kwargs={**kw, "_shared_cache": backend_cache})   # :261-269
        la.fit(loader)
        nml = -float(la.log_marginal_likelihood())
        res.neg_margliks.append(nml)

        with torch.no_grad():                             # fit() left the model in eval mode (:273)
            val_f = model(val_idx)
            val_loss = float(crit(val_f, val_y))
            val_acc = float((val_f.argmax(1) == val_y).float().mean())
        res.val_losses.append(val_loss)
        res.val_accs.append(val_acc)

        if not early_stop or ml_pat < patience:           # :277-296
            if nml < best_nml:
                best_nml, res.best_marglik_epoch, ml_pat = nml, epoch, 0
                res.best_marglik_state = deepcopy(model.state_dict())
            else:
                ml_pat += 1
        if not early_stop or val_pat < patience:
            if val_loss < best_val:
                best_val, res.best_valloss_epoch, val_pat = val_loss, epoch, 0
                res.best_valloss_state = deepcopy(model.state_dict())
            else:
                val_pat += 1
        res.stopped_epoch = epoch
        if early_stop and ml_pat >= patience and val_pat >= patience:
            break
    return res
