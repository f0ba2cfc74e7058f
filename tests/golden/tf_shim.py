"""A torch-backed, EAGER stand-in for the slice of the TensorFlow-1.8 API that the reference's model code uses
(AR.py, fitz_nag_NVP.py, SV_dense.py, optimisers/adamax.py), so that the reference's OWN unmodified classes -
IAF._create_flow, Flow_Stack, VI_SSM._ELBO / build_flow, AdamaxOptimizer._apply_dense - can be executed here and
their outputs stored as golden vectors (tests/golden/make_golden_step.py).

TensorFlow 1.8 itself cannot be installed in this image (no wheel for CPython 3.12).  What this file restates is
the TF LIBRARY semantics of ~30 ops, each a one-liner below with the documented behaviour it follows; what it
does NOT restate is anything in /root/reference: which tensors are sliced how, which terms enter the ELBO with
which sign and scale, the order variables are created in, what the optimiser does to its slots - all of that runs
from the reference's source.  Graph construction is evaluation: placeholders, random samples and variable
initial values are INJECTED (queues filled by the caller before the reference code runs), every op computes
immediately in float64, `compute_gradients` is torch.autograd, and ops with side effects (`apply_gradients`,
`minimize`) return a callable the caller runs explicitly - the eager counterpart of `sess.run(train_step)`.

Test infrastructure only; needs /root/reference at generation time, never imported by the package.
"""
import math
import sys
import types
from contextlib import contextmanager
from unittest import mock

import numpy as np
import torch

DT = torch.float64


class State:
    """Injected inputs and recorded objects of one build."""

    def __init__(self):
        self.placeholders = []      # values handed out by tf.placeholder, in creation order
        self.samples = []           # values handed out by <distribution>.sample(), in call order
        self.blob = None            # flat initial values of tf.layers variables, in creation order
        self.cursor = 0
        self.variables = []         # Var objects in creation order
        self.var_shapes = []

    def next_placeholder(self, shape):
        v = self.placeholders.pop(0)
        v = torch.as_tensor(np.asarray(v), dtype=DT)
        assert list(v.shape) == [int(s) for s in shape], ("placeholder shape", list(v.shape), shape)
        return v

    def next_sample(self):
        return torch.as_tensor(np.asarray(self.samples.pop(0)), dtype=DT)

    def new_variable(self, shape, kind):
        n = int(np.prod(shape))
        seg = self.blob[self.cursor:self.cursor + n]
        assert seg.numel() == n, "parameter blob exhausted: the creation order differs from the product's layout"
        self.cursor += n
        v = Var(seg.reshape(shape).clone().to(DT), "%s_%d" % (kind, len(self.variables)))
        self.variables.append(v)
        self.var_shapes.append(tuple(shape))
        return v


STATE = State()


class _DType:
    def __init__(self, name):
        self.name = name
        self.base_dtype = self

    def __eq__(self, other):
        return isinstance(other, _DType) and other.name == self.name

    def __hash__(self):
        return hash(self.name)


float32 = _DType("float32")
float16 = _DType("float16")


def _t(x):
    if isinstance(x, Var):
        return x.value
    if isinstance(x, torch.Tensor):
        return x
    return torch.as_tensor(np.asarray(x, dtype=np.float64), dtype=DT)


class Var:
    """tf.Variable: a leaf tensor; arithmetic on it reads the current value (as in a TF graph)."""

    def __init__(self, value, name):
        self.value = value.detach().clone().requires_grad_(True)
        self.name = name
        self.dtype = float32            # the reference's variables are float32; only `== tf.float16` is asked of it

    def assign(self, new):              # state_ops.assign: returns the new value
        self.value = _t(new).detach().clone().requires_grad_(True)
        return self.value.detach()

    def _v(self):
        return self.value.detach()

    def __mul__(self, o): return self._v() * _t(o)
    def __rmul__(self, o): return _t(o) * self._v()
    def __add__(self, o): return self._v() + _t(o)
    def __radd__(self, o): return _t(o) + self._v()
    def __sub__(self, o): return self._v() - _t(o)
    def __rsub__(self, o): return _t(o) - self._v()
    def __truediv__(self, o): return self._v() / _t(o)


# ------------------------------------------------------------------------------------------------------------
# tf.layers  (TF 1.8 defaults: use_bias=True, kernel created before bias, channels_last, 'valid' padding)
# ------------------------------------------------------------------------------------------------------------
def _only_trainable(kw):
    assert set(kw) <= {"trainable"} and kw.get("trainable", True), kw      # the scripts pass trainable=True at most


def dense(inputs, units, activation=None, **kw):
    """tf.layers.dense: outputs = activation(inputs . kernel + bias) on the LAST axis; kernel [in, units]."""
    _only_trainable(kw)
    x = _t(inputs)
    k = STATE.new_variable((x.shape[-1], units), "dense_kernel")
    b = STATE.new_variable((units,), "dense_bias")
    y = x @ k.value + b.value
    return activation(y) if activation is not None else y


def conv1d(inputs, filters, kernel_size, strides=1, padding="valid", activation=None, **kw):
    """tf.layers.conv1d, channels_last: out[n, m, f] = sum_k sum_c in[n, m*s + k, c] kernel[k, c, f] + bias[f]
    (cross-correlation, no kernel flip), kernel [kernel_size, in_channels, filters], 'valid': no padding."""
    _only_trainable(kw)
    assert padding.lower() == "valid"
    x = _t(inputs)
    k = STATE.new_variable((kernel_size, x.shape[-1], filters), "conv_kernel")
    b = STATE.new_variable((filters,), "conv_bias")
    y = torch.nn.functional.conv1d(x.transpose(1, 2), k.value.permute(2, 1, 0), b.value, stride=strides).transpose(1, 2)
    return activation(y) if activation is not None else y


def batch_normalization(inputs, training=False, **kw):
    """tf.layers.batch_normalization with training=False and freshly initialised statistics (the reference never
    runs the update ops): gamma * (x - moving_mean 0) / sqrt(moving_variance 1 + epsilon 1e-3) + beta;
    variables in creation order gamma, beta (the moving statistics are not trainable)."""
    x = _t(inputs)
    g = STATE.new_variable((x.shape[-1],), "bn_gamma")
    be = STATE.new_variable((x.shape[-1],), "bn_beta")
    return x * (g.value / math.sqrt(1.0 + 1e-3)) + be.value


# ------------------------------------------------------------------------------------------------------------
# tf.contrib.distributions
# ------------------------------------------------------------------------------------------------------------
class Normal:
    """tfd.Normal(loc, scale): elementwise; log_prob = -0.5 ((x - loc) / scale)^2 - 0.5 log(2 pi) - log(scale)."""

    def __init__(self, loc, scale, **kw):
        self.loc, self.scale = _t(loc), _t(scale)

    def sample(self, n=None, **kw):
        return STATE.next_sample()          # injected base noise (the caller's eps)

    def log_prob(self, x):
        x = _t(x)
        z = (x - self.loc) / self.scale
        return -0.5 * z * z - 0.5 * math.log(2.0 * math.pi) - torch.log(self.scale)


class MultivariateNormalDiag:
    """tfd.MultivariateNormalDiag(loc, scale_diag): log_prob sums the Normal log-densities over the LAST axis."""

    def __init__(self, loc, scale_diag, **kw):
        self.n = Normal(loc, scale_diag)

    def log_prob(self, x):
        return self.n.log_prob(x).sum(dim=-1)


# How a bijector reduces its log-determinant.  "literal": over the last `event_ndims` axes of whatever tensor it is
# handed - the TF-1.8 behaviour as remembered (Bijector._event_dims_tensor).  "per_state": the same, except that the
# leading (batch) axis is never reduced - the change of variables the LV script's comments describe.  The two differ only where a
# script hands a bijector with event_ndims = 2 a rank-2 [states, 2] tensor (lotka_volterra_partial_batch_fix_theta.py:
# 303-314: one scalar for all states, versus one value per state).  Which one TF 1.8 really computes cannot be settled
# without running it; tests/golden/make_golden_step_models.py stores both.
EVENT_REDUCTION = "literal"


def _reduce_event(v, event_ndims):
    if not event_ndims:
        return v
    first = v.dim() - event_ndims
    if EVENT_REDUCTION == "per_state":
        first = max(first, 1)
    return v.sum(dim=tuple(range(first, v.dim())))


class AffineBijector:
    """tfb.Affine(shift, scale_diag) (event_ndims = 1) and tfb.AffineScalar(shift, scale) (event_ndims = 0):
    forward(x) = scale * x + shift; inverse_log_det_jacobian = -sum log|scale| (a constant)."""

    def __init__(self, shift=0.0, scale_diag=None, scale=None, **kw):
        self.shift = _t(shift)
        self.scale = _t(scale_diag if scale_diag is not None else (1.0 if scale is None else scale))
        self.event_ndims = 1 if scale_diag is not None else 0

    def forward(self, x):
        return _t(x) * self.scale + self.shift

    def inverse(self, y):
        return (_t(y) - self.shift) / self.scale

    def inverse_log_det_jacobian(self, y):
        return -torch.log(torch.abs(self.scale)).sum()


class ChainBijector:
    """tfb.Chain([b_0, ..., b_n]): forward = b_0 o ... o b_n; inverse applies b_0^-1 first; the inverse log-determinant is
    the sum of every member's own (each reduced by the member's own event_ndims), evaluated along the inverse pass."""

    def __init__(self, bijectors, **kw):
        self.bijectors = list(bijectors)

    def forward(self, x):
        for b in reversed(self.bijectors):
            x = b.forward(x)
        return x

    def inverse(self, y):
        for b in self.bijectors:
            y = b.inverse(y)
        return y

    def inverse_log_det_jacobian(self, y):
        total = 0.0
        for b in self.bijectors:
            total = total + b.inverse_log_det_jacobian(y)
            y = b.inverse(y)
        return total


class TransformedDistribution:
    """tfd.TransformedDistribution(distribution, bijector): log_prob(y) = distribution.log_prob(bijector.inverse(y)) +
    bijector.inverse_log_det_jacobian(y)."""

    def __init__(self, distribution, bijector, **kw):
        self.distribution, self.bijector = distribution, bijector

    def log_prob(self, y):
        return self.distribution.log_prob(self.bijector.inverse(y)) + self.bijector.inverse_log_det_jacobian(y)


class SoftplusBijector:
    """tfb.Softplus(event_ndims): forward(x) = log(1 + exp(x)); inverse_log_det_jacobian(y) = -log(1 - exp(-y)) summed
    over the last `event_ndims` axes of y (the 1.8 bijectors take event_ndims in the constructor and reduce their
    log-determinants over that many trailing axes)."""

    def __init__(self, event_ndims=0, **kw):
        self.event_ndims = int(event_ndims)

    def forward(self, x):
        return torch.nn.functional.softplus(_t(x))

    def inverse(self, y):
        y = _t(y)
        return y + torch.log(-torch.expm1(-y))          # log(exp(y) - 1) without the overflow

    def inverse_log_det_jacobian(self, y):
        return _reduce_event(-torch.log(-torch.expm1(-_t(y))), self.event_ndims)


class _Bijectors:
    """tf.contrib.distributions.bijectors: Softplus, Affine, AffineScalar and Chain are real, everything else (the
    theta posterior's masked autoregressive flows, built in the scripts' module-level part) is a mock."""
    Softplus = SoftplusBijector
    Affine = AffineBijector
    AffineScalar = AffineBijector
    Chain = ChainBijector

    def __getattr__(self, name):
        return mock.MagicMock(name="bijectors." + name)


class InjectedDistribution:
    """Stands in for the theta posterior (tfd.TransformedDistribution of masked autoregressive flows, built in the
    scripts' main(), A11 of SURVEY section 8): hands out an injected sample and an injected log-density."""

    def __init__(self, theta, log_prob):
        self._theta, self._lp = torch.as_tensor(theta, dtype=DT), torch.as_tensor(log_prob, dtype=DT)

    def sample(self, n=None):
        self._theta = self._theta.detach().clone().requires_grad_(True)
        return self._theta

    def log_prob(self, x):
        return self._lp


# ------------------------------------------------------------------------------------------------------------
# optimizer base class (tensorflow.python.training.optimizer.Optimizer)
# ------------------------------------------------------------------------------------------------------------
class Optimizer:
    def __init__(self, use_locking, name):
        self._name = name
        self._slots = {}

    def _zeros_slot(self, var, slot_name, op_name):
        d = self._slots.setdefault(slot_name, {})
        if id(var) not in d:
            d[id(var)] = Var(torch.zeros_like(var.value), "%s/%s" % (var.name, slot_name))
        return d[id(var)]

    def get_slot(self, var, name):
        return self._slots[name][id(var)]

    def compute_gradients(self, loss):
        """tf.gradients of a non-scalar loss differentiates its SUM; one (gradient, variable) pair per trainable
        variable in creation order (None for a variable the loss does not depend on)."""
        vs = list(STATE.variables)
        gs = torch.autograd.grad(_t(loss).sum(), [v.value for v in vs], retain_graph=True, allow_unused=True)
        return list(zip(gs, vs))

    def apply_gradients(self, grads_and_vars):
        gv = [(g, v) for g, v in grads_and_vars if g is not None]

        def run():                              # the eager counterpart of sess.run(train_step)
            self._prepare()
            self._create_slots([v for _, v in gv])
            for g, v in gv:
                self._apply_dense(g.detach(), v)
        return run

    def minimize(self, loss):
        gv = self.compute_gradients(loss)
        return self.apply_gradients(gv)


def global_norm(t_list):
    """tf.global_norm: sqrt(sum of squared L2 norms), None entries ignored."""
    return torch.sqrt(sum((g.detach() ** 2).sum() for g in t_list if g is not None))


def clip_by_global_norm(t_list, clip_norm):
    """tf.clip_by_global_norm: t_i * clip_norm / max(global_norm, clip_norm); returns (list, global_norm)."""
    gn = global_norm(t_list)
    scale = clip_norm / torch.clamp(gn, min=clip_norm)
    return [None if g is None else g * scale for g in t_list], gn


@contextmanager
def name_scope(*a, **k):
    yield


def placeholder(dtype=None, shape=None, name=None):
    """tf.placeholder(dtype, shape): the next injected feed value."""
    return STATE.next_placeholder([shape] if isinstance(shape, int) else shape)


def split(value, num_or_size_splits, axis=0):
    v = _t(value)
    return list(torch.split(v, v.shape[axis] // num_or_size_splits, dim=axis))


def tile(x, multiples):
    return _t(x).repeat(*[int(m) for m in multiples])


def scatter_nd(indices, updates, shape):
    """tf.scatter_nd with unique indices: zeros(shape)[indices[i]] = updates[i]."""
    idx = torch.as_tensor(np.asarray(indices), dtype=torch.long)
    out = torch.zeros([int(s) for s in shape], dtype=DT)
    upd = _t(updates)
    out[tuple(idx[..., d] for d in range(idx.shape[-1]))] = upd
    return out


def install():
    """Registers the stand-in as `tensorflow` (+ the submodules the reference imports) in sys.modules."""
    tf = types.ModuleType("tensorflow")
    tf.float32, tf.float16, tf.int32 = float32, float16, _DType("int32")
    tf.shape = lambda x, **k: [int(d) for d in _t(x).shape]
    tf.zeros = lambda shape, dtype=None, **k: torch.zeros([int(d) for d in shape], dtype=DT)
    tf.ones = lambda shape, dtype=None, **k: torch.ones([int(d) for d in shape], dtype=DT)
    tf.reduce_min = lambda x, axis=None, **k: _t(x).min() if axis is None else _t(x).min(dim=axis).values
    tf.reduce_max = lambda x, axis=None, **k: _t(x).max() if axis is None else _t(x).max(dim=axis).values
    tf.reduce_prod = lambda x, axis=None, **k: _t(x).prod() if axis is None else _t(x).prod(dim=axis)
    tf.matrix_inverse = lambda x, **k: torch.linalg.inv(_t(x))
    tf.matrix_diag_part = lambda x, **k: torch.diagonal(_t(x), dim1=-2, dim2=-1)
    tf.eye = lambda n, **k: torch.eye(int(n), dtype=DT)
    # a host array converted at float32 is ROUNDED to float32 (then carried in float64 like everything else)
    tf.convert_to_tensor = lambda v, dtype=None, **k: _t(np.asarray(v, dtype=np.float32) if dtype == float32 and not isinstance(v, (torch.Tensor, Var)) else v)
    tf.set_random_seed = lambda *a, **k: None
    tf.InteractiveSession = lambda *a, **k: mock.MagicMock(name="session")
    tf.Session = tf.InteractiveSession
    tf.placeholder = placeholder
    tf.reduce_sum = lambda x, axis=None, **k: _t(x).sum() if axis is None else _t(x).sum(dim=axis)
    tf.reduce_mean = lambda x, axis=None, **k: _t(x).mean() if axis is None else _t(x).mean(dim=axis)
    tf.concat = lambda values, axis=0, **k: torch.cat([_t(v) for v in values], dim=axis)
    tf.expand_dims = lambda x, axis=None, **k: _t(x).unsqueeze(axis)
    tf.squeeze = lambda x, axis=None, **k: _t(x).squeeze() if axis is None else _t(x).squeeze(axis)
    tf.reshape = lambda x, shape, **k: _t(x).reshape([int(s) for s in shape])
    tf.transpose = lambda x, perm=None, **k: _t(x).permute(*perm) if perm is not None else _t(x).t()
    tf.split, tf.tile, tf.scatter_nd = split, tile, scatter_nd
    tf.log, tf.exp, tf.abs = (lambda x, **k: torch.log(_t(x))), (lambda x, **k: torch.exp(_t(x))), (lambda x, **k: torch.abs(_t(x)))
    tf.square, tf.sqrt = (lambda x, **k: _t(x) ** 2), (lambda x, **k: torch.sqrt(_t(x)))
    tf.maximum = lambda a, b, **k: torch.maximum(_t(a), _t(b))
    tf.ones_like, tf.zeros_like = (lambda x, **k: torch.ones_like(_t(x))), (lambda x, **k: torch.zeros_like(_t(x)))
    tf.constant = lambda v, dtype=None, **k: _t(v)
    tf.cast = lambda x, dtype=None, **k: x
    tf.name_scope = name_scope
    tf.global_norm, tf.clip_by_global_norm = global_norm, clip_by_global_norm
    tf.global_variables_initializer = lambda: None

    tf.nn = types.SimpleNamespace(
        elu=lambda x, **k: torch.nn.functional.elu(_t(x)),                # exp(x) - 1 for x < 0, x otherwise
        softplus=lambda x, **k: torch.nn.functional.softplus(_t(x)),      # log(1 + exp(x))
        relu=lambda x, **k: torch.relu(_t(x)))
    tf.layers = types.SimpleNamespace(dense=dense, conv1d=conv1d, batch_normalization=batch_normalization)
    tf.summary = mock.MagicMock(name="summary")
    tf.train = mock.MagicMock(name="train")
    bij = _Bijectors()
    tf.contrib = types.SimpleNamespace(distributions=types.SimpleNamespace(
        Normal=Normal, MultivariateNormalDiag=MultivariateNormalDiag, bijectors=bij,
        TransformedDistribution=TransformedDistribution))
    client = types.ModuleType("tensorflow.python.client")
    client.timeline = mock.MagicMock(name="timeline")
    sys.modules.setdefault("matplotlib", mock.MagicMock(name="matplotlib"))          # imported by the scripts, unused here
    sys.modules.setdefault("matplotlib.pyplot", mock.MagicMock(name="matplotlib.pyplot"))

    py = types.ModuleType("tensorflow.python")
    opsm = types.ModuleType("tensorflow.python.ops")
    clip_ops = types.ModuleType("tensorflow.python.ops.clip_ops")
    clip_ops.clip_by_global_norm, clip_ops.global_norm = clip_by_global_norm, global_norm
    control_flow_ops = types.ModuleType("tensorflow.python.ops.control_flow_ops")
    control_flow_ops.group = lambda *a, **k: None
    math_ops = types.ModuleType("tensorflow.python.ops.math_ops")
    math_ops.cast = lambda x, dtype=None, **k: x
    state_ops = types.ModuleType("tensorflow.python.ops.state_ops")
    state_ops.assign_sub = lambda var, delta, **k: var.assign(var._v() - _t(delta))
    state_ops.assign = lambda var, value, **k: var.assign(value)
    fw = types.ModuleType("tensorflow.python.framework")
    fw_ops = types.ModuleType("tensorflow.python.framework.ops")
    fw_ops.convert_to_tensor = lambda v, name=None, **k: v
    tr = types.ModuleType("tensorflow.python.training")
    tr_opt = types.ModuleType("tensorflow.python.training.optimizer")
    tr_opt.Optimizer = Optimizer
    opsm.clip_ops, opsm.control_flow_ops, opsm.math_ops, opsm.state_ops = clip_ops, control_flow_ops, math_ops, state_ops
    fw.ops, tr.optimizer = fw_ops, tr_opt
    py.ops, py.framework, py.training, py.client = opsm, fw, tr, client
    tf.python = py
    mods = {"tensorflow": tf, "tensorflow.python": py, "tensorflow.python.ops": opsm,
            "tensorflow.python.ops.clip_ops": clip_ops, "tensorflow.python.ops.control_flow_ops": control_flow_ops,
            "tensorflow.python.ops.math_ops": math_ops, "tensorflow.python.ops.state_ops": state_ops,
            "tensorflow.python.framework": fw, "tensorflow.python.framework.ops": fw_ops,
            "tensorflow.python.training": tr, "tensorflow.python.training.optimizer": tr_opt,
            "tensorflow.python.client": client}
    sys.modules.update(mods)
    return tf
