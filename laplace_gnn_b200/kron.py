"""Minimal stand-ins for the pieces of the ``laplace`` package that sit on either side of the
curvature-backend boundary, for machines where that package is not installed (the GPU box).

When ``laplace`` *is* importable (the reference fork or upstream laplace-torch), ``B200GGN``
returns that package's own ``Kron`` and is driven by its own ``KronLaplace`` — this module is not
used.  Here only the algebra the hot path needs is restated (device-resident, no host syncs):

* ``Kron`` / ``KronDecomposed``: block lists ``[[G, A], [G], ...]``, ``+``, scalar ``*``,
  ``decompose`` (eigh with eigenvalues clamped at 0) and ``logdet`` with per-block prior
  precisions            (reference: laplace/utils/matrix.py:16-145, 277-394; utils.py:193-226)
* ``KronLaplace`` / ``DiagLaplace`` / ``Laplace``: ``fit``, ``log_marginal_likelihood``,
  ``optimize_prior_precision``   (reference: laplace/baselaplace.py:778-973, 342-492, 1507-1627,
  1838-1919; laplace/laplace.py:13-47)
"""
from __future__ import annotations

import math
from typing import Iterable, Sequence

import torch
from torch.nn.utils import parameters_to_vector


def _eigh(m: torch.Tensor):
    if m.is_cuda and m.dtype == torch.float32 and m.shape[-1] <= 1024:
        lam, q = torch.linalg.eigh(m.double(), UPLO="U")
        return lam.to(m.dtype), q.to(m.dtype)
    return torch.linalg.eigh(m, UPLO="U")


def _eigh_psd(m: torch.Tensor):
    """symeig of the reference (laplace/utils/utils.py:193-226): eigh(UPLO="U"); when the solver does not
    converge (rank-deficient factors with many repeated tiny eigenvalues do that to LAPACK's / cusolver's fp32
    divide-and-conquer) the reference's jitter fallback, W L W^T + I = W (L + I) W^T: decompose M + I and take 1
    off the eigenvalues; then eigenvalues clamped at 0, NaN -> 0.  On the device the decomposition itself runs in
    float64 for n <= 1024 and is cast back: cusolver's double-precision path is 2-3x FASTER than the Jacobi solver
    torch picks for small fp32 matrices (B200: n = 256: 2.4 vs 5.3 ms, n = 500: 5.5 vs 15.6 ms;
    profiles/r1f_eigh_lab.txt) and its eigenvalues are exact to fp32 rounding."""
    try:
        lam, q = _eigh(m)
    except RuntimeError:               # did not converge (torch's LinAlgError is a RuntimeError)
        lam, q = _eigh(m + torch.eye(m.shape[-1], device=m.device, dtype=m.dtype))
        lam = lam - 1.0
    return torch.nan_to_num(lam.clamp(min=0.0)), torch.nan_to_num(q)


def _eigh_cached(h: torch.Tensor):
    """``_eigh_psd(h)``, except for a factor the backend marked as ``scale * raw`` with ``raw`` held in a cache that
    outlives the fit (``_eig_of``: the weight-independent input factor A_0 = X^T X / N with
    ``cache_input_factor``): the raw matrix is decomposed once, later fits scale its eigenvalues."""
    tag = getattr(h, "_eig_of", None)
    if tag is None:
        return _eigh_psd(h)
    cache, key, scale = tag
    if key not in cache:                      # the backend replaced its cache entry: decompose what we were given
        return _eigh_psd(h)
    ek = ("eig",) + tuple(key)
    if ek not in cache:
        cache[ek] = _eigh_psd(cache[key])
    lam, q = cache[ek]
    return lam * scale, q


def eigh_assignment(sizes: Sequence[int], world: int) -> list[int]:
    """Owner rank of each factor: largest first onto the least-loaded rank (cost ~ n^3), ties to the lower
    rank — a pure function of the sizes, so every rank computes the same table."""
    load = [0] * world
    owner = [0] * len(sizes)
    for i in sorted(range(len(sizes)), key=lambda i: (-sizes[i], i)):
        r = min(range(world), key=lambda r: (load[r], r))
        owner[i] = r
        load[r] += sizes[i] ** 3
    return owner


def _sharded_eigh(factors: Sequence[torch.Tensor], process_group):
    """[(eigenvalues, eigenvectors)] of every factor, each computed by one rank and all-gathered."""
    import torch.distributed as dist
    rank, world = dist.get_rank(process_group), dist.get_world_size(process_group)
    sizes = [int(h.shape[-1]) for h in factors]
    owner = eigh_assignment(sizes, world)
    seg = max(sum(n + n * n for n, o in zip(sizes, owner) if o == r) for r in range(world))
    dev, dt = factors[0].device, factors[0].dtype
    buf = torch.zeros(world, seg, dtype=dt, device=dev)
    off = 0
    for h, n, o in zip(factors, sizes, owner):
        if o == rank:
            lam, q = _eigh_psd(h)
            buf[rank, off:off + n] = lam
            buf[rank, off + n:off + n + n * n] = q.reshape(-1)
            off += n + n * n
    mine = buf[rank] if buf.is_cuda else buf[rank].clone()      # gloo (CPU tests): no in-place guarantee
    dist.all_gather_into_tensor(buf.view(-1), mine, group=process_group)
    offs = [0] * world
    out = []
    for n, o in zip(sizes, owner):
        a = offs[o]
        out.append((buf[o, a:a + n].clone(), buf[o, a + n:a + n + n * n].reshape(n, n).clone()))
        offs[o] = a + n + n * n
    return out


class Kron:
    """Block-diagonal Kronecker-factored matrix: one block per parameter tensor, a block is
    ``[G, A]`` (weight: G ⊗ A, output-side factor first) or ``[G]`` (bias)."""

    def __init__(self, kfacs: Sequence[Sequence[torch.Tensor]]):
        self.kfacs = [list(block) for block in kfacs]

    @classmethod
    def init_from_model(cls, model_or_params, device=None, dtype=None) -> "Kron":
        params = model_or_params.parameters() if isinstance(model_or_params, torch.nn.Module) \
            else model_or_params
        blocks = []
        for p in params:
            dev = device if device is not None else p.device
            dt = dtype if dtype is not None else p.dtype
            if p.ndim == 1:
                blocks.append([torch.zeros(p.shape[0], p.shape[0], device=dev, dtype=dt)])
            elif p.ndim >= 2:
                out_f, in_f = p.shape[0], int(p[0].numel())
                blocks.append([torch.zeros(out_f, out_f, device=dev, dtype=dt),
                               torch.zeros(in_f, in_f, device=dev, dtype=dt)])
            else:
                raise ValueError("scalar parameters are not supported")
        return cls(blocks)

    def __len__(self):
        return len(self.kfacs)

    def __add__(self, other: "Kron") -> "Kron":
        if not isinstance(other, Kron) or len(other) != len(self):
            raise ValueError("can only add a Kron with the same block structure")
        return Kron([[a + b for a, b in zip(fa, fb)] for fa, fb in zip(self.kfacs, other.kfacs)])

    def __mul__(self, s) -> "Kron":
        if isinstance(s, torch.Tensor) and s.numel() != 1:
            raise ValueError("Kron can only be scaled by a scalar")
        # spread the scalar evenly over the factors of a block so that the block scales by s
        return Kron([[(s ** (1.0 / len(f))) * h for h in f] for f in self.kfacs])

    __rmul__ = __mul__

    def decompose(self, damping: bool = False, process_group=None) -> "KronDecomposed":
        """eigh per factor (matrix.py:118-145).  A bias block's G that the backend marked as a copy
        of the preceding weight block's G (``_dup_of``) reuses that decomposition instead of
        repeating the identical eigh — same values, one third fewer eigendecompositions.

        With a ``process_group`` (the multi-GPU pass: every rank holds the same all-reduced factors) the
        distinct factors are spread over the ranks, each rank decomposes its share and ONE all-gather hands
        everybody the same eigenpairs — bit-identical across ranks by construction, and the replicated
        2L sequential cusolver calls (12 ms on the products shape, 4 % of an 8-GPU step) become one."""
        distinct, owner_of = [], {}
        for f in self.kfacs:
            for h in f:
                src = getattr(h, "_dup_of", None)
                if src is not None and id(src) in owner_of:
                    owner_of[id(h)] = owner_of[id(src)]
                else:
                    owner_of[id(h)] = len(distinct)
                    distinct.append(h)
        world = 1
        if process_group is not None:
            import torch.distributed as dist
            world = dist.get_world_size(process_group)
        if world > 1 and len(distinct) > 1:
            pairs = _sharded_eigh(distinct, process_group)
        else:
            pairs = [_eigh_cached(h) for h in distinct]
        vecs, vals = [], []
        for f in self.kfacs:
            vals.append([pairs[owner_of[id(h)]][0] for h in f])
            vecs.append([pairs[owner_of[id(h)]][1] for h in f])
        return KronDecomposed(vecs, vals, damping=damping)

    def diag(self) -> torch.Tensor:
        parts = []
        for f in self.kfacs:
            if len(f) == 1:
                parts.append(f[0].diagonal())
            else:
                parts.append(torch.outer(f[0].diagonal(), f[1].diagonal()).reshape(-1))
        return torch.cat(parts)

    def logdet(self) -> torch.Tensor:
        total = 0.0
        for f in self.kfacs:
            if len(f) == 1:
                total = total + torch.logdet(f[0])
            else:
                (g, a) = f
                total = total + a.shape[0] * torch.logdet(g) + g.shape[0] * torch.logdet(a)
        return total


class KronDecomposed:
    """Eigendecomposed ``Kron`` plus an additive per-block ``delta`` (the prior precision)."""

    def __init__(self, eigenvectors, eigenvalues, deltas: torch.Tensor | None = None,
                 damping: bool = False):
        self.eigenvectors = eigenvectors
        self.eigenvalues = eigenvalues
        dev = eigenvalues[0][0].device
        self.deltas = torch.zeros(len(eigenvalues), device=dev) if deltas is None else deltas
        self.damping = damping

    def __len__(self):
        return len(self.eigenvalues)

    def __add__(self, deltas: torch.Tensor) -> "KronDecomposed":
        deltas = torch.as_tensor(deltas, device=self.deltas.device, dtype=self.deltas.dtype)
        if deltas.ndim == 0 or deltas.numel() == 1:
            deltas = deltas.reshape(-1).expand(len(self))
        if deltas.numel() != len(self):
            raise ValueError("prior precision must be a scalar or one value per parameter tensor")
        return KronDecomposed(self.eigenvectors, self.eigenvalues, self.deltas + deltas, self.damping)

    def __mul__(self, s) -> "KronDecomposed":
        # spread evenly over the eigenvalue lists of a block, like the reference (matrix.py:347-366): the same
        # products without damping, the same damped terms with it
        vals = [[(s ** (1.0 / len(f))) * lam for lam in f] for f in self.eigenvalues]
        return KronDecomposed(self.eigenvectors, vals, self.deltas, self.damping)

    __rmul__ = __mul__

    def bmm(self, W: torch.Tensor, exponent: float = 1.0) -> torch.Tensor:
        """``self ** exponent @ W`` row-wise for W [S, P] (matrix.py:396-475): weight blocks are
        flattened [d_out, d_in] row-major with factors [G, A]."""
        if W.ndim == 1:
            return self.bmm(W.unsqueeze(0), exponent).squeeze(0)
        if W.ndim != 2:
            raise ValueError("W must be [P] or [S, P]")
        out, cur = [], 0
        for lams, Qs, delta in zip(self.eigenvalues, self.eigenvectors, self.deltas):
            if len(lams) == 1:
                p = lams[0].numel()
                Wp = W[:, cur:cur + p].T
                out.append((Qs[0] @ (torch.pow(lams[0] + delta, exponent).reshape(-1, 1) * (Qs[0].T @ Wp))).T)
            else:
                l1, l2 = lams
                p = l1.numel() * l2.numel()
                if self.damping:
                    scale = torch.pow(torch.outer(l1 + delta.sqrt(), l2 + delta.sqrt()), exponent)
                else:
                    scale = torch.pow(torch.outer(l1, l2) + delta, exponent)
                Wp = W[:, cur:cur + p].reshape(-1, l1.numel(), l2.numel())
                Wp = (Qs[0].T @ Wp @ Qs[1]) * scale.unsqueeze(0)
                out.append((Qs[0] @ Wp @ Qs[1].T).reshape(-1, p))
            cur += p
        if cur != W.shape[1]:
            raise ValueError("W has the wrong number of parameters")
        return torch.cat(out, dim=1)

    def logdet(self) -> torch.Tensor:
        total = 0.0
        for lams, delta in zip(self.eigenvalues, self.deltas):
            if len(lams) == 1:
                total = total + torch.log(lams[0] + delta).sum()
            elif self.damping:
                total = total + torch.log(torch.outer(lams[0] + delta.sqrt(), lams[1] + delta.sqrt())).sum()
            else:
                total = total + torch.log(torch.outer(lams[0], lams[1]) + delta).sum()
        return total


class _ParametricLaplaceLite:
    """The subset of ``ParametricLaplace`` on the hot path (classification, all weights)."""

    def __init__(self, model: torch.nn.Module, likelihood: str = "classification",
                 prior_precision=1.0, prior_mean=0.0, temperature: float = 1.0,
                 backend=None, backend_kwargs: dict | None = None, **unused):
        if likelihood != "classification":
            raise NotImplementedError("the GCN Laplace hot path is classification only")
        # reference keywords that would change the result must not be swallowed: only their neutral values pass
        # (baselaplace.py:94-107: sigma_noise=1, enable_backprop=False; KronLaplace :1520-1540: damping=False)
        neutral = {"sigma_noise": (1, 1.0), "enable_backprop": (False,), "damping": (False,),
                   "dict_key_x": ("input_ids",), "dict_key_y": ("labels",), "asdl_fisher_kwargs": (None,)}
        for key, val in unused.items():
            if key not in neutral:
                raise TypeError(f"{type(self).__name__}: unexpected keyword argument {key!r}")
            if not any(val is ok or val == ok for ok in neutral[key]):
                raise NotImplementedError(f"{type(self).__name__}: {key}={val!r} is outside the hot path this stand-in "
                                          "restates; drive B200GGN from the reference's laplace package for it")
        if backend is None:
            from .curvature import B200GGN
            backend = B200GGN
        self.model = model
        self.likelihood = likelihood
        # same filter as the fork (baselaplace.py:118-122): skip adj / norms parameters
        self.params = [p for k, p in model.named_parameters()
                       if p.requires_grad and "adj" not in k and "norms" not in k]
        self.n_params = sum(p.numel() for p in self.params)
        self.n_layers = len(self.params)
        self._device = self.params[0].device
        self.prior_precision = prior_precision
        self.prior_mean = prior_mean
        self.temperature = temperature
        self._backend_cls = backend
        self._backend_kwargs = dict(backend_kwargs or {})
        self._backend = None
        self.loss = 0.0
        self.n_data = 0
        self.n_outputs = 0
        self.H = None

    @property
    def backend(self):
        if self._backend is None:
            self._backend = self._backend_cls(self.model, self.likelihood, **self._backend_kwargs)
        return self._backend

    # ---- prior
    @property
    def prior_precision_diag(self) -> torch.Tensor:
        pp = torch.as_tensor(self.prior_precision, device=self._device, dtype=torch.float32)
        if pp.ndim == 0 or pp.numel() == 1:
            return pp.reshape(-1) * torch.ones(self.n_params, device=self._device)
        if pp.numel() == self.n_params:
            return pp
        if pp.numel() == self.n_layers:
            return torch.cat([v * torch.ones(p.numel(), device=self._device)
                              for v, p in zip(pp, self.params)])
        raise ValueError("prior precision must be scalar, per parameter tensor, or diagonal")

    @property
    def _H_factor(self) -> float:
        return 1.0 / self.temperature

    @property
    def log_likelihood(self) -> torch.Tensor:
        return -self._H_factor * self.loss

    @property
    def scatter(self) -> torch.Tensor:
        d = self.mean - self.prior_mean
        return (d * self.prior_precision_diag) @ d

    @property
    def log_det_prior_precision(self) -> torch.Tensor:
        return self.prior_precision_diag.log().sum()

    @property
    def log_det_ratio(self) -> torch.Tensor:
        return self.log_det_posterior_precision - self.log_det_prior_precision

    def log_marginal_likelihood(self, prior_precision=None) -> torch.Tensor:
        if prior_precision is not None:
            self.prior_precision = prior_precision
        return self.log_likelihood - 0.5 * (self.log_det_ratio + self.scatter)

    # ---- fit
    def _init_H(self):
        raise NotImplementedError

    def _curv_closure(self, X, y, N):
        raise NotImplementedError

    def fit(self, train_loader: Iterable, override: bool = True) -> None:
        if override:
            self._init_H()
            self.loss = 0.0
            self.n_data = 0
        self.model.eval()
        self.mean = parameters_to_vector(self.params).detach()
        N = len(train_loader.dataset)
        for X, y in train_loader:
            X, y = X.to(self._device), y.to(self._device)
            loss_b, H_b = self._curv_closure(X, y, N)
            self.loss = self.loss + loss_b
            self.H = H_b if self.H is None else self.H + H_b
        self.n_outputs = getattr(self.backend, "n_outputs", self.n_outputs)
        self.n_data += N

    # ---- checkpoint / resume (baselaplace.py:1314-1374, :1664-1676)
    def state_dict(self) -> dict:
        if self.H is None:
            raise AttributeError("Laplace not fitted. Run fit() first.")
        return {"mean": self.mean, "H": self._H_state(), "loss": self.loss, "prior_mean": self.prior_mean,
                "prior_precision": self.prior_precision, "n_data": self.n_data, "n_outputs": self.n_outputs,
                "likelihood": self.likelihood, "temperature": self.temperature,
                "cls_name": self.__class__.__name__}

    def load_state_dict(self, state_dict: dict) -> None:
        if self.__class__.__name__ != state_dict["cls_name"]:
            raise ValueError("Loading a wrong Laplace type. Make sure `subset_of_weights` and"
                             " `hessian_structure` are correct!")
        if len(state_dict["mean"]) != self.n_params:
            raise ValueError("Attempting to load Laplace with different number of parameters than the model.")
        if self.likelihood != state_dict["likelihood"]:
            raise ValueError("Different likelihoods detected!")
        self.mean = state_dict["mean"]
        self.loss = state_dict["loss"]
        self.prior_mean = state_dict["prior_mean"]
        self.prior_precision = state_dict["prior_precision"]
        self.n_data = state_dict["n_data"]
        self.n_outputs = state_dict["n_outputs"]
        self.temperature = state_dict["temperature"]
        self._load_H_state(state_dict["H"])

    def _H_state(self):
        return self.H

    def _load_H_state(self, H) -> None:
        self.H = H

    def optimize_prior_precision(self, init_prior_prec=1.0, n_steps: int = 100, lr: float = 0.1,
                                 prior_structure: str = "scalar", verbose: bool = False):
        """Marginal-likelihood tuning of the prior precision (baselaplace.py:419-463): Adam on
        log prior precision; every step only re-evaluates the cheap marglik algebra."""
        if prior_structure == "scalar":
            shape = (1,)
        elif prior_structure == "layerwise":
            shape = (self.n_layers,)
        elif prior_structure == "diag":
            shape = (self.n_params,)
        else:
            raise ValueError(prior_structure)
        log_pp = (torch.ones(shape, device=self._device) * math.log(float(init_prior_prec))).requires_grad_(True)
        opt = torch.optim.Adam([log_pp], lr=lr)
        for _ in range(n_steps):
            opt.zero_grad()
            neg = -self.log_marginal_likelihood(prior_precision=log_pp.exp())
            neg.backward()
            opt.step()
        self.prior_precision = log_pp.detach().exp()
        return self.prior_precision


class KronLaplace(_ParametricLaplaceLite):
    """``Laplace(model, "classification", subset_of_weights="all", hessian_structure="kron")``."""

    def _init_H(self):
        self.H = None
        self.H_facs = None

    def _curv_closure(self, X, y, N):
        return self.backend.kron(X, y, N=N)

    @staticmethod
    def _rescale_factors(kron: Kron, factor: float) -> Kron:
        for f in kron.kfacs:                      # only the input-side factor carries the 1/N (baselaplace.py:1574-1578)
            if len(f) == 2:
                f[1] *= factor
        return kron

    def fit(self, train_loader, override: bool = True) -> None:
        """``override=False`` continues a fitted posterior with more data the way the reference does
        (baselaplace.py:1580-1610): the old factors are discounted by n_old / (n_old + n_new), the new ones by
        n_new / (n_new + n_old), loss and n_data accumulate."""
        if override:
            self.H_facs = None
        n_old, n_new = self.n_data, len(train_loader.dataset)
        if self.H_facs is not None:
            self.H_facs = self._rescale_factors(self.H_facs, n_old / (n_old + n_new))
        self.H = None                            # the batches of this call are summed on their own
        super().fit(train_loader, override=override)
        if self.H_facs is None:
            self.H_facs = self.H
        else:
            self.H_facs = self.H_facs + self._rescale_factors(self.H, n_new / (n_new + n_old))
        # multi-GPU pass: the factors are all-reduced, the eigendecompositions are spread over the ranks
        pg = getattr(self.backend, "process_group", None)
        shard = pg is not None and getattr(self.backend, "shard_eigh", False)
        self.H = self.H_facs.decompose(process_group=pg) if shard else self.H_facs.decompose()

    def _H_state(self):                          # the reference stores the undecomposed factors (:1664-1668)
        return self.H_facs.kfacs

    def _load_H_state(self, kfacs) -> None:      # ... and re-decomposes on load (:1670-1676)
        self.H_facs = Kron(kfacs)
        self.H = self.H_facs.decompose()

    @property
    def posterior_precision(self) -> KronDecomposed:
        pp = torch.as_tensor(self.prior_precision, device=self._device, dtype=torch.float32)
        if pp.ndim and pp.numel() == self.n_params and self.n_params != self.n_layers:
            raise ValueError("a diagonal prior is not representable in a Kron posterior")
        return self.H * self._H_factor + pp

    @property
    def log_det_posterior_precision(self) -> torch.Tensor:
        return self.posterior_precision.logdet()

    def sample(self, n_samples: int = 100, generator: torch.Generator | None = None) -> torch.Tensor:
        """Weight samples from N(mean, P^-1) (baselaplace.py:1646-1655)."""
        eps = torch.randn(n_samples, self.n_params, device=self._device, generator=generator)
        return self.mean.reshape(1, -1) + self.posterior_precision.bmm(eps, exponent=-0.5)

    def __call__(self, x: torch.Tensor, pred_type: str = "nn", link_approx: str = "mc", n_samples: int = 100,
                 generator: torch.Generator | None = None, samples: torch.Tensor | None = None) -> torch.Tensor:
        """MC predictive ``la(idx, pred_type="nn", link_approx="mc", n_samples=...)`` — the call of
        ``mc_eval`` (gnn/marglik_training.py:341-353; baselaplace.py:1183-1199): class probabilities
        of the nodes ``x`` averaged over weight samples.  All samples of a tile go through the graph
        together as extra right-hand-side columns of the SpMM."""
        if pred_type != "nn" or link_approx != "mc":
            raise NotImplementedError("only the sampling predictive (pred_type='nn', link_approx='mc') is built")
        from .predictive import mc_predictive
        if samples is None:
            samples = self.sample(n_samples, generator)
        return mc_predictive(self.model, samples, x)


class DiagLaplace(_ParametricLaplaceLite):
    """``hessian_structure="diag"``: H is the vector diag(GGN)."""

    def _init_H(self):
        self.H = None

    def _curv_closure(self, X, y, N):
        return self.backend.diag(X, y, N=N)

    @property
    def posterior_precision(self) -> torch.Tensor:
        return self._H_factor * self.H + self.prior_precision_diag

    @property
    def log_det_posterior_precision(self) -> torch.Tensor:
        return self.posterior_precision.log().sum()


def Laplace(model, likelihood="classification", subset_of_weights="all", hessian_structure="kron",
            **kwargs):
    """Factory with the reference's signature (laplace/laplace.py:13-47), hot-path subset.  One deliberate
    difference: the reference's default is ``subset_of_weights="last_layer"``; only "all" — what every call site of
    gnn/marglik_training.py passes (:197-201, :261-265, :653-657) — is built here, so it is the default."""
    if subset_of_weights != "all":
        raise NotImplementedError("only subset_of_weights='all' is on the hot path")
    table = {"kron": KronLaplace, "diag": DiagLaplace}
    if hessian_structure not in table:
        raise NotImplementedError(f"hessian_structure={hessian_structure!r} is outside the hot path")
    return table[hessian_structure](model, likelihood, **kwargs)
