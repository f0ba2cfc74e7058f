"""Variational posterior over the global parameters theta (A11 of SURVEY §8a) — host-side PyTorch.

Restates `tfd.TransformedDistribution(Normal(loc, scale), Chain(reversed([Invert(MAF), Permute, ...])))`
(AR.py:376-391; fitz_nag_NVP.py:480-494; SV_dense.py:428-442) with
`masked_autoregressive_default_template(hidden_layers=[5,5,5])`.  About 580 parameters and p x dtheta
numbers per step: far too small for a kernel, so it stays an autograd module whose parameters are
VIEWS into the tail of the flat device blob the Adamax kernel updates (one global-norm clip over
everything, like AR.py:228-234).  The arithmetic lives in tf.contrib.distributions (TensorFlow 1.8,
not vendored by the reference): parity unpinned, restated from the published algorithm
(Papamakarios et al. 2017; TF `masked_dense`/`_gen_mask` block masks).
"""
from __future__ import annotations

import math
from typing import List, Optional, Sequence, Tuple

import numpy as np
import torch


def _gen_mask(num_blocks: int, n_in: int, n_out: int, exclusive: bool) -> np.ndarray:
    """Block mask [n_in, n_out] of TF's masked_dense: output block b sees input blocks <= b (< b if exclusive)."""
    mask = np.zeros((n_out, n_in), dtype=np.float32)
    d_in, d_out = n_in // num_blocks, n_out // num_blocks
    row = d_out if exclusive else 0
    col = 0
    for _ in range(num_blocks):
        mask[row:, col:col + d_in] = 1.0
        col += d_in
        row += d_out
    return mask.T.copy()


class ThetaFlow:
    """num_bijectors inverse-MAF layers with fixed permutations in between.

    sample():  z ~ N(loc, scale)^d ; for each layer: z <- (z - shift(z)) * exp(-log_scale(z)) ; z <- z[perm]
    log_prob:  sum log N(z0) + sum_layers sum_k log_scale_k   (the forward log-det of Invert(MAF) is -sum log_scale)
    """

    HIDDEN = (5, 5, 5)

    def __init__(self, dtheta: int, num_bijectors: int, base_loc: float, base_scale: float, activation: str = "elu",
                 permutations: Optional[Sequence[Sequence[int]]] = None, tf_mask_grad: bool = True,
                 softplus_out: bool = False):
        # TensorFlow's masked_dense does not multiply the mask in the forward pass: masked kernel entries are zero through
        # the initialiser and a kernel_constraint re-applied after every update, so they DO receive a gradient (which
        # counts in tf.global_norm, AR.py:230) that the constraint then wipes.  tf_mask_grad=True reproduces that (call
        # `constrain()` after each update); False multiplies the mask in the forward pass (zero gradient there).
        self.tf_mask_grad = bool(tf_mask_grad)
        # the chain ends in tfb.Softplus(event_ndims=2) (lotka_volterra_partial_batch.py:741): theta = softplus(u) > 0 and
        # log q(theta) = log q_u(u) - sum log sigmoid(u)
        self.softplus_out = bool(softplus_out)
        self.d = dtheta
        self.nb = num_bijectors
        self.base_loc = float(base_loc)
        self.base_scale = float(base_scale)
        self.act = torch.nn.functional.elu if activation == "elu" else torch.relu
        if permutations is None:
            # the reference draws them from numpy's global stream while building the graph (AR.py:384-385)
            permutations = [np.random.permutation(np.arange(0, dtheta)) for _ in range(num_bijectors - 1)]
        self.perms = [np.asarray(pm, dtype=np.int64) for pm in permutations]
        dims = (dtheta,) + self.HIDDEN + (2 * dtheta,)
        self.shapes: List[Tuple[int, int]] = [(dims[i], dims[i + 1]) for i in range(len(dims) - 1)]
        self.masks_np = [_gen_mask(dtheta, a, b, exclusive=(i == 0)) for i, (a, b) in enumerate(self.shapes)]
        self.n_params = num_bijectors * sum(a * b + b for a, b in self.shapes)
        self.views: List[List[Tuple[torch.Tensor, torch.Tensor]]] = []
        self.masks: List[torch.Tensor] = []
        self.flat: Optional[torch.Tensor] = None

    # ------------------------------------------------------------------
    def init_values(self, gen: torch.Generator) -> torch.Tensor:
        """Glorot-normal masked kernels, zero biases (TF masked_dense defaults)."""
        out = []
        for _ in range(self.nb):
            for (a, b), m in zip(self.shapes, self.masks_np):
                std = math.sqrt(2.0 / (a + b))
                out.append((torch.randn(a, b, generator=gen) * std * torch.from_numpy(m)).reshape(-1))
                out.append(torch.zeros(b))
        return torch.cat(out)

    def bind(self, flat: torch.Tensor) -> None:
        """`flat`: a leaf tensor (requires_grad) of n_params values living inside the optimiser's blob."""
        assert flat.numel() == self.n_params
        self.flat = flat
        self.masks = [torch.from_numpy(m).to(flat.device) for m in self.masks_np]
        self._perm_t = [torch.from_numpy(pm).to(flat.device) for pm in self.perms]

    def mask_flat(self) -> torch.Tensor:
        """Per-variable multiplier of the kernel constraint: the block masks on kernels, 1 on biases."""
        one = []
        for (a, b), m in zip(self.shapes, self.masks_np):
            one.append(torch.from_numpy(m).reshape(-1))
            one.append(torch.ones(b))
        return torch.cat(one * self.nb)

    def constrain(self) -> None:
        """kernel_constraint=lambda x: mask * x of masked_dense, applied after an update of the bound variables."""
        if getattr(self, "_mask_flat", None) is None or self._mask_flat.device != self.flat.device:
            self._mask_flat = self.mask_flat().to(self.flat.device)
        with torch.no_grad():
            self.flat.mul_(self._mask_flat)

    def _unpack(self):
        layers = []
        off = 0
        for _ in range(self.nb):
            one = []
            for (a, b) in self.shapes:
                w = self.flat[off:off + a * b].reshape(a, b); off += a * b
                bias = self.flat[off:off + b]; off += b
                one.append((w, bias))
            layers.append(one)
        return layers

    def _shift_log_scale(self, layer, z):
        h = z
        for i, (w, b) in enumerate(layer):
            h = h @ (w if self.tf_mask_grad else w * self.masks[i]) + b
            if i < len(layer) - 1:
                h = self.act(h)
        h = h.reshape(z.shape[0], self.d, 2)
        shift, log_scale = h[..., 0], h[..., 1]
        # clip_by_value_preserve_gradient(log_scale, -5, 3)
        log_scale = log_scale + (log_scale.clamp(-5.0, 3.0) - log_scale).detach()
        return shift, log_scale

    def sample_and_log_prob(self, z0: torch.Tensor):
        """z0 [p, d] ~ N(base_loc, base_scale).  Returns (theta [p,d], log q(theta) [p])."""
        lp = (-0.5 * ((z0 - self.base_loc) / self.base_scale) ** 2 - 0.5 * math.log(2 * math.pi)
              - math.log(self.base_scale)).sum(dim=1)
        z = z0
        for k, layer in enumerate(self._unpack()):
            shift, log_scale = self._shift_log_scale(layer, z)
            z = (z - shift) * torch.exp(-log_scale)
            lp = lp + log_scale.sum(dim=1)
            if k < self.nb - 1:
                z = z[:, self._perm_t[k]]
        if self.softplus_out:
            lp = lp - torch.nn.functional.logsigmoid(z).sum(dim=1)
            z = torch.nn.functional.softplus(z)
        return z, lp

    def base_sample(self, p: int, gen: Optional[torch.Generator], device) -> torch.Tensor:
        return self.base_loc + self.base_scale * torch.randn(p, self.d, generator=gen, device=device)


def prior_tensors(priors: Sequence[Tuple[float, float]], device, dtype=torch.float32):
    """(mean, scale) of the diagonal Gaussian prior as device tensors (build once: no H2D copy per step)."""
    mean = torch.tensor([m for m, _ in priors], dtype=dtype, device=device)
    scale = torch.tensor([s for _, s in priors], dtype=dtype, device=device)
    return mean, scale


def prior_log_prob(theta: torch.Tensor, priors) -> torch.Tensor:
    """MultivariateNormalDiag(prior_mean, prior_scale).log_prob(theta) (AR.py:178-182).
    `priors`: list of (mean, scale) pairs, or the (mean, scale) tensors of `prior_tensors`."""
    if isinstance(priors, tuple) and len(priors) == 2 and isinstance(priors[0], torch.Tensor):
        mean, scale = priors
    else:
        mean, scale = prior_tensors(priors, theta.device, theta.dtype)
    return (-0.5 * ((theta - mean) / scale) ** 2 - 0.5 * math.log(2 * math.pi) - torch.log(scale)).sum(dim=1)
