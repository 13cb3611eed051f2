"""Simulation plugins accepted for UserInput.model['simulateByInputParametersOnlyFunction'].

The reference calls an arbitrary Python function once per point (PEUQSE/InverseProblem.py:2825).  Here the
simulator must be either
  * one of the built-in CUDA model kernels below (evaluated inside the fused kernels), or
  * a vectorised torch callable ``f(theta[B,k] float64 cuda) -> sim[B, n_total] float64 cuda`` wrapped in
    TorchSimulator (responses concatenated in order); the likelihood then runs on the returned block.
There is no CPU path.
"""
import numpy as np

from . import _lib

KB_KJ = 1.380649e-26 * 6.0221409e+23   # kJ mol^-1 K^-1, as in the reference's Examples/Example01/tprmodel.py:17


class BuiltinModel:
    model_id = None
    n_params = None

    def n_total(self):
        raise NotImplementedError

    def constants(self, order):
        """Flat float64 constants with per-observation rows permuted by `order` (the kernels' row order)."""
        raise NotImplementedError

    def with_conditions(self, values):
        """A copy of this model under other independent variables (design of experiments): the batched counterpart of the
        reference's populateIndependentVariablesFunction plugin, which pushes the values into the simulator's module
        namespace before every run (PEUQSE/InverseProblem.py:1403)."""
        raise NotImplementedError(f"{type(self).__name__} has no independent variables to modulate")

    def __call__(self, theta):
        """Single-point simulation through the CUDA kernel (keeps the plugin callable like the reference's)."""
        if getattr(self, "_owner", None) is None:
            raise _lib.PeuqseError("built-in model is not attached to a parameter_estimation object yet")
        return self._owner.getSimulatedResponses(theta)


class QuadraticModel(BuiltinModel):
    """y_j = (x_j - a)**2 + b  (reference: Examples/Example00/simulation_model_00.py:6-10, 19-24)."""
    model_id = _lib.MODEL_QUADRATIC
    n_params = 2

    def __init__(self, x):
        self.x = np.asarray(x, dtype=float).reshape(-1)

    def n_total(self):
        return len(self.x)

    def with_conditions(self, values):
        return QuadraticModel(values)

    def constants(self, order):
        return np.ascontiguousarray(self.x[order])


class LangmuirLogModel(BuiltinModel):
    """log10 of a Langmuir coverage at the temperatures T + offsets: the simulator of the reference's design-of-experiments
    example (Examples/Example13doeFunctionExample/simulation_functions_Langmuir_CO_H2O_four_parameters.py:23-52):
        delta_G = dH - T_j*dS;  K = exp(-delta_G/(kB*T_j));  theta_A = K*pA/(1 + K*pA);  sim_j = log10(theta_A).
    theta = (dH, dS); the independent variables are (T, pA)."""
    model_id = _lib.MODEL_LANGMUIR_LOG
    n_params = 2
    KB_EV = 8.61733035E-5       # eV/K, simulation_functions_Langmuir_CO_H2O_four_parameters.py:31

    def __init__(self, T=398.15, pA=0.1, offsets=(0.0, -25.0, 25.0)):
        self.T, self.pA = float(T), float(pA)
        self.offsets = tuple(float(o) for o in offsets)

    def n_total(self):
        return len(self.offsets)

    def with_conditions(self, values):
        return LangmuirLogModel(values[0], values[1], self.offsets)

    def populateIndependentVariables(self, values):
        """Usable as UserInput.model['populateIndependentVariablesFunction'] (takes effect at the next parameter_estimation)."""
        self.T, self.pA = float(values[0]), float(values[1])

    def constants(self, order):
        Tj = np.array([self.T + o for o in self.offsets])
        return np.ascontiguousarray(np.concatenate([Tj[order], (self.KB_EV * Tj)[order], np.full(len(Tj), self.pA)]))


class Tpd2SiteModel(BuiltinModel):
    """Two-site first-order desorption rate with prescribed coverages:
    sim_j = sum_i cov_ij * exp(-(Ea_i - kBT_j*logA_i - gamma_i*cov_ij)/kBT_j), the rate law of the
    reference's Examples/Example01/tprmodel.py:20.  theta = (Ea1, Ea2, logA1, logA2, gamma1, gamma2)."""
    model_id = _lib.MODEL_TPD2SITE
    n_params = 6

    def __init__(self, kBT, cov1, cov2):
        self.kBT = np.asarray(kBT, dtype=float).reshape(-1)
        self.cov1 = np.asarray(cov1, dtype=float).reshape(-1)
        self.cov2 = np.asarray(cov2, dtype=float).reshape(-1)
        assert len(self.kBT) == len(self.cov1) == len(self.cov2)

    @classmethod
    def linear_ramp(cls, n=1024, t_end=400.0, T0=150.0, beta=0.5, tau1=150.0, tau2=300.0, cov0=0.5):
        t = np.linspace(0, t_end, n)
        T = T0 + beta * t
        return cls(KB_KJ * T, cov0 * np.exp(-t / tau1), cov0 * np.exp(-t / tau2))

    def n_total(self):
        return len(self.kBT)

    def constants(self, order):
        return np.ascontiguousarray(np.concatenate([self.kBT[order], self.cov1[order], self.cov2[order]]))


class LinearModel(BuiltinModel):
    """sim = Phi @ theta, accumulated parameter by parameter (d ascending)."""
    model_id = _lib.MODEL_LINEAR

    def __init__(self, Phi):
        self.Phi = np.ascontiguousarray(np.asarray(Phi, dtype=float))
        assert self.Phi.ndim == 2
        self.n_params = self.Phi.shape[1]

    def n_total(self):
        return self.Phi.shape[0]

    def constants(self, order):
        return np.ascontiguousarray(self.Phi[order]).reshape(-1)


class TorchSimulator:
    """Marks `fn` as a vectorised torch simulator: fn(theta[B,k]) -> sim[B, n_total] (float64, CUDA)."""
    model_id = _lib.MODEL_EXTERNAL

    def __init__(self, fn):
        self.fn = fn

    def __call__(self, theta):
        return self.fn(theta)
