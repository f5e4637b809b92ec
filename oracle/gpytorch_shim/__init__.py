"""Minimal stand-in for the gpytorch / linear_operator API.  TEST INFRASTRUCTURE ONLY.

Purpose: gpytorch==1.11 and linear_operator==0.5.0 (reference requirements.txt:1-2)
cannot be installed in the build container (no network, no wheel).  This shim lets
the reference's OWN source file ``projectedlmc/projected_lmc.py`` be imported and
executed unmodified -- ``ProjectedGPModel`` / ``ProjectedLMCmll`` / ``LMCMixingMatrix`` and
their helpers run line for line -- while the third-party layer underneath (kernel
evaluation, Gaussian likelihood, MultivariateNormal.log_prob with the Cholesky path,
exact prediction, Kronecker sums) is provided here from the restated gpytorch
semantics in ``oracle/plmc_oracle.py``.  It is used ONLY by
``tests/golden/make_golden_reference.py`` to generate fixtures in the build container
(the reference checkout does not exist on the GPU box).

What this pins: every line of the reference's model/MLL layer (projection, the three
correction terms, the task-noise assembly, the prediction assembly, initial values and
parametrisations).  What it cannot pin: gpytorch's own arithmetic, which is restated.
"""
from __future__ import annotations

import importlib.util
import math
import sys
import types

import torch

from .. import plmc_oracle as O

_INSTALLED = False


# ----------------------------------------------------------------------------- helpers
def _auto_module(name):
    """Module whose unknown attributes resolve to inert placeholder classes (enough for the
    class statements / default arguments / annotations of the out-of-scope models)."""
    m = types.ModuleType(name)

    def __getattr__(attr, _m=m):
        if attr.startswith("__"):
            raise AttributeError(attr)
        cls = type(attr, (torch.nn.Module,), {"__init__": lambda self, *a, **k: torch.nn.Module.__init__(self)})
        setattr(_m, attr, cls)
        return cls

    m.__getattr__ = __getattr__
    return m


class DenseOp:
    """Dense stand-in for a LinearOperator."""

    def __init__(self, t):
        self.t = t.t if isinstance(t, DenseOp) else t

    def evaluate(self):
        return self.t

    to_dense = evaluate

    @property
    def shape(self):
        return self.t.shape

    @property
    def dtype(self):
        return self.t.dtype

    @property
    def device(self):
        return self.t.device

    def add_jitter(self, eps=1e-3):
        n = self.t.shape[-1]
        return DenseOp(self.t + eps * torch.eye(n, dtype=self.t.dtype, device=self.t.device))

    def sum(self, dim):
        return DenseOp(self.t.sum(dim))

    def diagonal(self, **kw):
        return torch.diagonal(self.t, dim1=-2, dim2=-1)


def _dense(x):
    return x.t if isinstance(x, DenseOp) else x


class GPModule(torch.nn.Module):
    def register_constraint(self, param_name, constraint, replace=True):
        self.add_module(param_name + "_constraint", constraint)

    def initialize(self, **kwargs):
        for k, v in kwargs.items():
            setattr(self, k, v)
        return self


# ----------------------------------------------------------------------------- constraints
class GreaterThan(torch.nn.Module):
    def __init__(self, lower_bound, **kw):
        super().__init__()
        self.register_buffer("lower_bound", torch.as_tensor(float(lower_bound)))
        self.register_buffer("upper_bound", torch.as_tensor(math.inf))

    def transform(self, raw):
        return torch.nn.functional.softplus(raw) + self.lower_bound

    def inverse_transform(self, v):
        x = v - self.lower_bound
        return x + torch.log(-torch.expm1(-x))


class Positive(GreaterThan):
    def __init__(self, **kw):
        super().__init__(0.0)


# ----------------------------------------------------------------------------- means
class Mean(GPModule):
    def __init__(self, *a, **k):
        super().__init__()


class ZeroMean(Mean):
    def __init__(self, batch_shape=torch.Size(), **kwargs):
        super().__init__()
        self.batch_shape = batch_shape

    def forward(self, x):
        return torch.zeros(*self.batch_shape, x.shape[-2], dtype=x.dtype, device=x.device)


class ConstantMean(Mean):
    pass


# ----------------------------------------------------------------------------- kernels
class Kernel(GPModule):
    has_lengthscale = True
    _kind = None

    def __init__(self, ard_num_dims=None, batch_shape=torch.Size(), active_dims=None, lengthscale_prior=None,
                 lengthscale_constraint=None, **kwargs):
        super().__init__()
        self.ard_num_dims = ard_num_dims
        self.batch_shape = torch.Size(batch_shape)
        d = 1 if ard_num_dims is None else ard_num_dims
        self.register_parameter("raw_lengthscale", torch.nn.Parameter(torch.zeros(*self.batch_shape, 1, d)))
        self.register_constraint("raw_lengthscale", lengthscale_constraint or Positive())

    @property
    def lengthscale(self):
        return self.raw_lengthscale_constraint.transform(self.raw_lengthscale)

    @lengthscale.setter
    def lengthscale(self, value):
        value = torch.as_tensor(value, dtype=self.raw_lengthscale.dtype)
        with torch.no_grad():
            self.raw_lengthscale.copy_(self.raw_lengthscale_constraint.inverse_transform(
                value.expand_as(self.raw_lengthscale)))

    def forward(self, x1, x2):
        same = torch.equal(x1, x2)
        # gpytorch zero-fills the diagonal only when x1 == x2 and the scaled inputs carry no grad
        carries_grad = torch.is_grad_enabled() and (x1.requires_grad or x2.requires_grad
                                                    or self.raw_lengthscale.requires_grad)
        zero_diag = same and not carries_grad
        return O.base_kernel(self._kind, x1, x2, self.lengthscale, zero_diag=zero_diag)

    def __call__(self, x1, x2=None, **kw):
        if x1.dim() == 1:
            x1 = x1.unsqueeze(-1)
        x2 = x1 if x2 is None else (x2.unsqueeze(-1) if x2.dim() == 1 else x2)
        return DenseOp(self.forward(x1, x2))


class RBFKernel(Kernel):
    _kind = "rbf"


class MaternKernel(Kernel):
    def __init__(self, nu=2.5, **kwargs):
        super().__init__(**kwargs)
        self.nu = nu
        self._kind = {2.5: "matern52", 1.5: "matern32", 0.5: "matern12"}[nu]


class ScaleKernel(Kernel):
    has_lengthscale = False

    def __init__(self, base_kernel, batch_shape=torch.Size(), **kwargs):
        GPModule.__init__(self)
        self.base_kernel = base_kernel
        self.batch_shape = torch.Size(batch_shape)
        self.register_parameter("raw_outputscale", torch.nn.Parameter(torch.zeros(*self.batch_shape)))
        self.register_constraint("raw_outputscale", Positive())

    @property
    def outputscale(self):
        return self.raw_outputscale_constraint.transform(self.raw_outputscale)

    def forward(self, x1, x2):
        return self.base_kernel.forward(x1, x2) * self.outputscale[..., None, None]


# ----------------------------------------------------------------------------- distributions
class Distribution:
    pass


class MultivariateNormal(Distribution):
    def __init__(self, mean, covariance_matrix):
        self._mean = mean
        self._covar = DenseOp(covariance_matrix)

    @property
    def mean(self):
        return self._mean

    loc = mean

    @property
    def lazy_covariance_matrix(self):
        return self._covar

    @property
    def covariance_matrix(self):
        return self._covar.t

    @property
    def batch_shape(self):
        return self._mean.shape[:-1]

    @property
    def event_shape(self):
        return self._mean.shape[-1:]

    @property
    def variance(self):
        return torch.diagonal(self._covar.t, dim1=-2, dim2=-1)

    def log_prob(self, value):
        tries = settings.cholesky_max_tries.value()
        return O.mvn_log_prob(self._covar.t, value - self._mean, max_tries=tries)


class MultitaskMultivariateNormal(Distribution):
    """mean [n, t]; covariance over the point-major / task-minor vectorisation (interleaved)."""

    def __init__(self, mean, covariance_matrix, **kw):
        self._mean = mean
        self._covar = DenseOp(covariance_matrix)

    @property
    def mean(self):
        return self._mean

    @property
    def lazy_covariance_matrix(self):
        return self._covar

    @property
    def variance(self):
        return torch.diagonal(self._covar.t, dim1=-2, dim2=-1).reshape(self._mean.shape)

    @property
    def stddev(self):
        return self.variance.sqrt()

    def confidence_region(self):
        s2 = self.stddev * 2
        return self.mean - s2, self.mean + s2


# ----------------------------------------------------------------------------- likelihoods
class Likelihood(GPModule):
    pass


class _GaussianLikelihoodBase(Likelihood):
    pass


class _HomoskedasticNoise(GPModule):
    def __init__(self, noise_constraint=None, batch_shape=torch.Size()):
        super().__init__()
        self.register_parameter("raw_noise", torch.nn.Parameter(torch.zeros(*batch_shape, 1)))
        self.register_constraint("raw_noise", noise_constraint or GreaterThan(1e-4))

    @property
    def noise(self):
        return self.raw_noise_constraint.transform(self.raw_noise)


class GaussianLikelihood(_GaussianLikelihoodBase):
    def __init__(self, noise_constraint=None, batch_shape=torch.Size(), **kwargs):
        super().__init__()
        self.noise_covar = _HomoskedasticNoise(noise_constraint, torch.Size(batch_shape))

    @property
    def noise(self):
        return self.noise_covar.noise

    def forward(self, dist, *params, **kw):
        K = dist.covariance_matrix
        n = K.shape[-1]
        noise = self.noise  # [*batch, 1]
        return MultivariateNormal(dist.mean, K + torch.diag_embed(noise.expand(*K.shape[:-2], n)))


class MultitaskGaussianLikelihood(Likelihood):
    def __init__(self, num_tasks, rank=0, has_global_noise=True, **kwargs):
        super().__init__()
        assert rank > 0 and not has_global_noise, "shim supports the configuration full_likelihood() uses"
        self.num_tasks = num_tasks
        self.register_parameter("task_noise_covar_factor", torch.nn.Parameter(torch.randn(num_tasks, rank)))

    @property
    def task_noise_covar(self):
        Fm = self.task_noise_covar_factor
        return Fm @ Fm.transpose(-1, -2)

    def forward(self, dist, *params, **kw):
        n = dist.mean.shape[-2]
        noise = torch.kron(torch.eye(n, dtype=dist.mean.dtype), self.task_noise_covar)
        return MultitaskMultivariateNormal(dist.mean, dist.lazy_covariance_matrix.t + noise)


# ----------------------------------------------------------------------------- models / mlls
class ExactGP(GPModule):
    def __init__(self, train_inputs, train_targets, likelihood):
        super().__init__()
        if torch.is_tensor(train_inputs):
            train_inputs = (train_inputs,)
        self.train_inputs = tuple(t.unsqueeze(-1) if t.ndimension() == 1 else t for t in train_inputs)
        self.train_targets = train_targets
        self.likelihood = likelihood
        self.prediction_strategy = None

    def set_train_data(self, inputs=None, targets=None, strict=True):
        if inputs is not None:
            if torch.is_tensor(inputs):
                inputs = (inputs,)
            self.train_inputs = tuple(t.unsqueeze(-1) if t.ndimension() == 1 else t for t in inputs)
        if targets is not None:
            self.train_targets = targets
        self.prediction_strategy = None

    def __call__(self, *args, **kwargs):
        inputs = [a.unsqueeze(-1) if a.ndimension() == 1 else a for a in args]
        if self.training:
            if not all(torch.equal(a, b) for a, b in zip(self.train_inputs, inputs)):
                raise RuntimeError("You must train on the training inputs!")
            return torch.nn.Module.__call__(self, *inputs, **kwargs)
        # exact prediction (gpytorch DefaultPredictionStrategy): joint prior on [X; X*]
        X, Xs = self.train_inputs[0], inputs[0]
        n = X.shape[-2]
        joint = torch.nn.Module.__call__(self, torch.cat([X, Xs], dim=-2), **kwargs)
        mean, K = joint.mean, joint.covariance_matrix
        train = self.likelihood(MultivariateNormal(mean[..., :n], K[..., :n, :n]))
        L = O.psd_safe_cholesky(train.covariance_matrix, settings.cholesky_max_tries.value())
        resid = (self.train_targets - mean[..., :n]).unsqueeze(-1)
        mean_cache = torch.cholesky_solve(resid, L).squeeze(-1)
        Kst = K[..., n:, :n]
        pred_mean = mean[..., n:] + (Kst @ mean_cache.unsqueeze(-1)).squeeze(-1)
        V = torch.linalg.solve_triangular(L, Kst.transpose(-1, -2), upper=False)
        pred_cov = K[..., n:, n:] - V.transpose(-1, -2) @ V
        return MultivariateNormal(pred_mean, pred_cov)


class MarginalLogLikelihood(GPModule):
    def __init__(self, likelihood, model):
        super().__init__()
        self.likelihood = likelihood
        self.model = model


class ExactMarginalLogLikelihood(MarginalLogLikelihood):
    def _add_other_terms(self, res, params):
        return res  # no added-loss terms / priors in the configurations exercised


# ----------------------------------------------------------------------------- settings
class _Setting:
    _default = None
    _value = None

    def __init__(self, value=None, *a, **k):
        self._new = value

    @classmethod
    def value(cls):
        return cls._default if cls._value is None else cls._value

    def __enter__(self):
        self._old = type(self)._value
        type(self)._value = self._new
        return self

    def __exit__(self, *exc):
        type(self)._value = self._old
        return False


settings = types.ModuleType("gpytorch.settings")


class _cholesky_max_tries(_Setting):
    _default = 3


settings.cholesky_max_tries = _cholesky_max_tries
for _n in ("cholesky_jitter", "max_cholesky_size", "fast_computations", "skip_posterior_variances",
           "skip_logdet_forward", "cg_tolerance", "eval_cg_tolerance", "debug"):
    setattr(settings, _n, type(_n, (_Setting,), {}))


# ----------------------------------------------------------------------------- linear_operator
class RootLinearOperator(DenseOp):
    def __init__(self, root):
        super().__init__(root @ root.transpose(-1, -2))


class KroneckerProductLinearOperator(DenseOp):
    def __init__(self, A, B):
        a, b = _dense(A), _dense(B)
        out = torch.einsum("...ij,...kl->...ikjl", a, b)
        super().__init__(out.reshape(*out.shape[:-4], a.shape[-2] * b.shape[-2], a.shape[-1] * b.shape[-1]))


def to_linear_operator(x):
    return DenseOp(x)


# ----------------------------------------------------------------------------- installation
def install():
    """Register the shim under the names the reference imports."""
    global _INSTALLED
    if _INSTALLED:
        return
    gp = _auto_module("gpytorch")
    gp.Module = GPModule
    gp.settings = settings
    table = {
        "means": dict(Mean=Mean, ZeroMean=ZeroMean, ConstantMean=ConstantMean),
        "means.mean": dict(Mean=Mean),
        "kernels": dict(Kernel=Kernel, RBFKernel=RBFKernel, MaternKernel=MaternKernel, ScaleKernel=ScaleKernel),
        "kernels.kernel": dict(Kernel=Kernel),
        "likelihoods": dict(Likelihood=Likelihood, GaussianLikelihood=GaussianLikelihood,
                            MultitaskGaussianLikelihood=MultitaskGaussianLikelihood),
        "likelihoods.likelihood": dict(Likelihood=Likelihood),
        "likelihoods.gaussian_likelihood": dict(_GaussianLikelihoodBase=_GaussianLikelihoodBase,
                                                GaussianLikelihood=GaussianLikelihood),
        "mlls": dict(ExactMarginalLogLikelihood=ExactMarginalLogLikelihood,
                     MarginalLogLikelihood=MarginalLogLikelihood),
        "mlls.exact_marginal_log_likelihood": dict(ExactMarginalLogLikelihood=ExactMarginalLogLikelihood),
        "distributions": dict(Distribution=Distribution, MultivariateNormal=MultivariateNormal,
                              MultitaskMultivariateNormal=MultitaskMultivariateNormal),
        "distributions.multivariate_normal": dict(MultivariateNormal=MultivariateNormal),
        "models": dict(ExactGP=ExactGP),
        "constraints": dict(GreaterThan=GreaterThan, Positive=Positive),
        "variational": {},
        "priors": {},
    }
    mods = {"gpytorch": gp}
    for path, names in table.items():
        full = "gpytorch." + path
        m = _auto_module(full)
        for k, v in names.items():
            setattr(m, k, v)
        mods[full] = m
    for full, m in mods.items():
        if full.count(".") >= 1:
            parent, leaf = full.rsplit(".", 1)
            setattr(mods[parent], leaf, m)
    mods["gpytorch.settings"] = settings
    lo = _auto_module("linear_operator")
    loo = _auto_module("linear_operator.operators")
    for k, v in dict(KroneckerProductLinearOperator=KroneckerProductLinearOperator,
                     RootLinearOperator=RootLinearOperator).items():
        setattr(loo, k, v)
    lod = _auto_module("linear_operator.operators.dense_linear_operator")
    lod.to_linear_operator = to_linear_operator
    lo.operators = loo
    loo.dense_linear_operator = lod
    mods.update({"linear_operator": lo, "linear_operator.operators": loo,
                 "linear_operator.operators.dense_linear_operator": lod})
    for name, m in mods.items():
        if name in sys.modules:
            raise RuntimeError(f"{name} is already imported; the shim must not shadow a real install")
        sys.modules[name] = m
    _INSTALLED = True


def load_reference(path="/root/reference/projectedlmc/projected_lmc.py"):
    """Import the reference's module, unmodified, on top of the shim."""
    install()
    spec = importlib.util.spec_from_file_location("_reference_projected_lmc", path)
    mod = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mod)
    return mod
