"""peuqse_b200: B200-native posterior-evaluation engine with the PEUQSE user interface.

    import peuqse_b200 as PEUQSE
    import peuqse_b200.UserInput as UserInput
    ...
    PE_object = PEUQSE.parameter_estimation(UserInput)
    PE_object.doMetropolisHastings()

The hot path (log-prior + simulation + log-likelihood and the samplers / sweeps around it) runs in
libpeuqse_b200.so (hand-written sm_100a CUDA behind the C ABI of include/peuqse_b200.h).
"""
from . import UserInput  # noqa: F401
from .inverse_problem import parameter_estimation  # noqa: F401
from .models import BuiltinModel, LangmuirLogModel, LinearModel, QuadraticModel, TorchSimulator, Tpd2SiteModel  # noqa: F401
from ._lib import PeuqseError  # noqa: F401

__version__ = "0.1.0"
