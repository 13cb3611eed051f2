"""ctypes binding of libpeuqse_b200.so (the C ABI declared in include/peuqse_b200.h).

There is no CPU fallback: if the shared library is missing or a call fails, an exception is raised.
PyTorch is used by the callers only for device memory and streams; no torch type crosses this boundary.
"""
import ctypes as C
import os

MAX_K = 16
MAX_RESP = 8

MODEL_EXTERNAL, MODEL_QUADRATIC, MODEL_TPD2SITE, MODEL_LINEAR, MODEL_LANGMUIR_LOG = 0, 1, 2, 3, 4
RESP_MVN_DIAG, RESP_PER_VALUE, RESP_MVN_FULL = 0, 1, 2
PRIOR_NONE, PRIOR_DIAG, PRIOR_DENSE = 0, 1, 2

c_double_p = C.POINTER(C.c_double)
c_int32_p = C.POINTER(C.c_int32)


class ProblemDesc(C.Structure):
    _fields_ = [
        ("k", C.c_int32), ("scaling_on", C.c_int32), ("has_lb", C.c_int32), ("has_ub", C.c_int32),
        ("lb", c_double_p), ("ub", c_double_p), ("scale", c_double_p), ("mu_scaled", c_double_p),
        ("prior_mode", C.c_int32), ("kp", C.c_int32),
        ("prior_gidx", c_int32_p), ("prior_w", c_double_p), ("prior_U", c_double_p), ("prior_c", C.c_double),
        ("n_resp", C.c_int32), ("n_total", C.c_int32),
        ("resp_n", c_int32_p), ("resp_mode", c_int32_p), ("resp_c", c_double_p),
        ("y", c_double_p), ("w", c_double_p), ("cval", c_double_p), ("src", c_int32_p),
        ("resp_U", C.POINTER(c_double_p)),
        ("model_id", C.c_int32), ("n_model_const", C.c_int32), ("model_const", c_double_p),
    ]


class SweepSpec(C.Structure):
    _fields_ = [
        ("kind", C.c_int32), ("k", C.c_int32),
        ("center", C.c_double * MAX_K), ("interval", C.c_double * MAX_K), ("radii", C.c_int64 * MAX_K),
        ("spread", C.c_double), ("std_prior", C.c_double * MAX_K), ("seed", C.c_uint64),
    ]


class Best(C.Structure):
    _fields_ = [("logp", C.c_double), ("index", C.c_int64)]


class MHSpec(C.Structure):
    _fields_ = [
        ("n_chains", C.c_int64), ("n_steps", C.c_int64), ("burn_in", C.c_int64), ("first_step", C.c_int64),
        ("chain_offset", C.c_int64), ("seed", C.c_uint64), ("modulate", C.c_double),
        ("k", C.c_int32), ("thin", C.c_int32), ("step_chol", C.c_double * (MAX_K * MAX_K)),
        ("map_finding", C.c_int32), ("reserved", C.c_int32),
    ]


class StatsSpec(C.Structure):
    _fields_ = [("k", C.c_int32), ("filter", C.c_int32), ("coef", C.c_double), ("kl_bins", C.c_int32), ("reserved", C.c_int32)]


FILTER_NONE, FILTER_AUTO, FILTER_MAP, FILTER_MU_AP = 0, 1, 2, 3
# offsets inside the packed statistics result (PEUQSE_STATS_* macros)
STATS = dict(N_TOTAL=0, N_KEPT=1, CUT=2, THRESHOLD=3, MEAN_LOGNEG=4, STD_LOGNEG=5, MAX_LOGP_ALL=6, MAP_LOGP=7, MAP_INDEX=8,
             MAP_INDEX_ALL=9, EVIDENCE=10, INFO_GAIN_LOG_RATIO=11, INFO_GAIN_KL=12, SUM_EXP=13, FROB2=14, MEAN_ALL=16, MU_AP=32,
             STDAP=48, MAP_THETA=64, MIN=80, MAX=96, KL_EACH=112, JOINT_LO=128, JOINT_HI=144, DOUBLES=160)


# offsets inside one scenario's statistics block (PEUQSE_SCEN_STATS_* macros)
def scen_stats_layout(k):
    return dict(N_TOTAL=0, N_KEPT=1, THRESHOLD=2, MAP_LOGP=3, MAP_INDEX=4, EVIDENCE=5, INFO_GAIN_LOG_RATIO=6, CUT=7, MU_AP=8, STDAP=8 + k,
                MAP_THETA=8 + 2 * k, DOUBLES=8 + 3 * k)


class ESSSpec(C.Structure):
    _fields_ = [
        ("n_walkers", C.c_int64), ("n_iters", C.c_int64), ("first_iter", C.c_int64), ("seed", C.c_uint64),
        ("mu0", C.c_double), ("tune", C.c_int32), ("max_rounds", C.c_int32), ("k", C.c_int32), ("driver", C.c_int32),
    ]


def acc_layout(k):
    """Offsets inside the pass-2 accumulator (PEUQSE_ACC_* macros)."""
    return dict(W=0, WEXP=1, NKEPT=2, S1=3, S2=3 + k, MIN=3 + k + k * k, MAX=3 + 2 * k + k * k,
                SIZE=3 + 3 * k + k * k)


def mh_stat_layout(k):
    """Offsets inside one chain's statistics block (PEUQSE_MH_STAT_* macros)."""
    return dict(NACC=0, MAPLOGP=1, MAPROW=2, LASTLOGP=3, SUMEXP=4, MAPTHETA=5, LAST=5 + k, S1=5 + 2 * k,
                S2=5 + 3 * k, DYN=5 + 3 * k + k * k, SIZE=6 + 3 * k + k * k)


EXPORTS = [
    "peuqse_last_error", "peuqse_version", "peuqse_device_info", "peuqse_bench_fp64", "peuqse_problem_create", "peuqse_problem_destroy",
    "peuqse_logp_batch", "peuqse_logprior_batch", "peuqse_logp_from_sim", "peuqse_loglik_simvar", "peuqse_transform_transient", "peuqse_simulate_batch", "peuqse_sweep_work_bytes",
    "peuqse_sweep_pass1", "peuqse_sweep_best_point", "peuqse_sweep_points", "peuqse_sweep_pass2", "peuqse_sweep_edges", "peuqse_sweep_hist",
    "peuqse_comm_unique_id", "peuqse_comm_create", "peuqse_comm_destroy", "peuqse_comm_allgather",
    "peuqse_comm_allreduce_sum_i64", "peuqse_peer_alloc", "peuqse_peer_create", "peuqse_peer_destroy",
    "peuqse_peer_max_bytes", "peuqse_peer_status", "peuqse_peer_merge_best", "peuqse_peer_merge_acc", "peuqse_peer_sum_i64", "peuqse_merge_best", "peuqse_merge_acc", "peuqse_mh_run",
    "peuqse_mhs_state_doubles", "peuqse_mhs_begin", "peuqse_mhs_propose", "peuqse_mhs_accept", "peuqse_mhs_finish",
    "peuqse_philox_blocks", "peuqse_ess_work_bytes", "peuqse_ess_run",
    "peuqse_essr_create", "peuqse_essr_destroy", "peuqse_essr_begin_half", "peuqse_essr_round", "peuqse_essr_end_half",
    "peuqse_essr_end_iteration", "peuqse_essr_info",
    "peuqse_stats_work_bytes", "peuqse_stats_begin", "peuqse_stats_finish", "peuqse_stats_compact",
    "peuqse_scenarios_create", "peuqse_scenarios_destroy", "peuqse_scenarios_synthesize", "peuqse_scenarios_set_observed",
    "peuqse_scenarios_logp", "peuqse_mh_run_scenarios", "peuqse_scenarios_stats",
    "peuqse_geweke_work_bytes", "peuqse_geweke", "peuqse_scenarios_kl", "peuqse_act_concat",
]

_lib = None


class PeuqseError(RuntimeError):
    pass


def lib_path():
    return os.path.join(os.path.dirname(os.path.abspath(__file__)), "libpeuqse_b200.so")


def load():
    """Load the CUDA library (once). Raises if it is not built: there is no other implementation."""
    global _lib
    if _lib is not None:
        return _lib
    path = lib_path()
    if not os.path.exists(path):
        raise PeuqseError(f"{path} not found: build it with ./build.sh (python -c 'import __graft_entry__ as g; g.build()'). "
                          "peuqse_b200 has no CPU fallback.")
    L = C.CDLL(path)
    vp, i64, i32, u64, u32, dbl = C.c_void_p, C.c_int64, C.c_int32, C.c_uint64, C.c_uint32, C.c_double
    L.peuqse_last_error.restype = C.c_char_p
    L.peuqse_last_error.argtypes = []
    L.peuqse_version.restype = C.c_int
    L.peuqse_device_info.argtypes = [c_int32_p]
    L.peuqse_bench_fp64.argtypes = [i32, c_double_p, vp, vp]
    L.peuqse_problem_create.argtypes = [C.POINTER(ProblemDesc), C.POINTER(vp)]
    L.peuqse_problem_destroy.argtypes = [vp]
    L.peuqse_logp_batch.argtypes = [vp, vp, i64, vp, vp, vp, vp]
    L.peuqse_logprior_batch.argtypes = [vp, vp, i64, vp, vp]
    L.peuqse_logp_from_sim.argtypes = [vp, vp, vp, i64, vp, vp, vp, vp]
    L.peuqse_loglik_simvar.argtypes = [vp, vp, vp, vp, C.POINTER(i32), C.POINTER(i32), i32, i64, vp, vp]
    L.peuqse_transform_transient.argtypes = [vp, vp, C.POINTER(i32), C.POINTER(i32), i32, i64, vp]
    L.peuqse_simulate_batch.argtypes = [vp, vp, i64, vp, vp]
    L.peuqse_sweep_work_bytes.restype = i64
    L.peuqse_sweep_work_bytes.argtypes = [vp, C.POINTER(SweepSpec), i64, i64]
    L.peuqse_sweep_pass1.argtypes = [vp, C.POINTER(SweepSpec), i64, i64, vp, vp, vp, vp]
    L.peuqse_sweep_points.argtypes = [C.POINTER(SweepSpec), i64, i64, vp, vp]
    L.peuqse_sweep_best_point.argtypes = [C.POINTER(SweepSpec), vp, vp, vp]
    L.peuqse_sweep_pass2.argtypes = [vp, C.POINTER(SweepSpec), i64, i64, vp, vp, dbl, vp, vp, vp, i64, vp, vp, vp]
    L.peuqse_sweep_edges.argtypes = [vp, i32, i32, vp, vp, vp]
    L.peuqse_sweep_hist.argtypes = [vp, C.POINTER(SweepSpec), i64, i64, vp, vp, dbl, vp, i32, vp, vp, i32, vp]
    L.peuqse_comm_unique_id.argtypes = [vp]
    L.peuqse_comm_create.argtypes = [vp, i32, i32, C.POINTER(vp)]
    L.peuqse_comm_destroy.argtypes = [vp]
    L.peuqse_comm_allgather.argtypes = [vp, vp, vp, i64, vp]
    L.peuqse_comm_allreduce_sum_i64.argtypes = [vp, vp, i64, vp]
    L.peuqse_peer_alloc.argtypes = [C.POINTER(vp), vp]
    L.peuqse_peer_create.argtypes = [vp, vp, i32, i32, C.POINTER(vp)]
    L.peuqse_peer_destroy.argtypes = [vp]
    L.peuqse_peer_max_bytes.argtypes = []
    L.peuqse_peer_max_bytes.restype = i64
    L.peuqse_peer_status.argtypes = [vp]
    L.peuqse_peer_merge_best.argtypes = [vp, vp, vp]
    L.peuqse_peer_merge_acc.argtypes = [vp, vp, i32, vp]
    L.peuqse_peer_sum_i64.argtypes = [vp, vp, i64, vp]
    L.peuqse_merge_best.argtypes = [vp, i32, vp, vp]
    L.peuqse_merge_acc.argtypes = [vp, i32, i32, vp, vp]
    L.peuqse_mh_run.argtypes = [vp, C.POINTER(MHSpec), vp, vp, vp, vp, vp, vp, vp, vp, vp]
    L.peuqse_mhs_state_doubles.argtypes = [i32]
    L.peuqse_mhs_state_doubles.restype = i64
    L.peuqse_mhs_begin.argtypes = [C.POINTER(MHSpec), vp, vp, vp, vp, vp, vp, vp, vp, vp, vp]
    L.peuqse_mhs_propose.argtypes = [C.POINTER(MHSpec), i64, vp, vp, vp, vp, vp, vp]
    L.peuqse_mhs_accept.argtypes = [C.POINTER(MHSpec), i64, vp, vp, vp, vp, vp, vp, vp, vp, vp, vp]
    L.peuqse_mhs_finish.argtypes = [C.POINTER(MHSpec), vp, vp, vp]
    L.peuqse_philox_blocks.argtypes = [u64, u32, i64, i64, u32, vp, vp]
    L.peuqse_ess_work_bytes.restype = i64
    L.peuqse_ess_work_bytes.argtypes = [i64, i32]
    L.peuqse_ess_run.argtypes = [vp, C.POINTER(ESSSpec), vp, vp, vp, vp, vp, vp, vp]
    L.peuqse_essr_create.argtypes = [C.POINTER(ESSSpec), C.POINTER(vp)]
    L.peuqse_essr_destroy.argtypes = [vp]
    L.peuqse_essr_begin_half.argtypes = [vp, i64, i32, vp, vp, vp, C.POINTER(i64), vp]
    L.peuqse_essr_round.argtypes = [vp, vp, vp, C.POINTER(i64), vp]
    L.peuqse_essr_end_half.argtypes = [vp, vp, vp, vp]
    L.peuqse_essr_end_iteration.argtypes = [vp, vp]
    L.peuqse_essr_info.argtypes = [vp, vp, vp]
    L.peuqse_stats_work_bytes.restype = i64
    L.peuqse_stats_work_bytes.argtypes = []
    L.peuqse_stats_begin.argtypes = [C.POINTER(StatsSpec), vp, vp, i64, vp, vp, vp]
    L.peuqse_stats_finish.argtypes = [C.POINTER(StatsSpec), vp, vp, vp, i64, vp, vp, i64, vp, vp, vp]
    L.peuqse_stats_compact.argtypes = [C.POINTER(StatsSpec), vp, vp, vp, i64, vp, vp, vp, vp, vp, vp]
    L.peuqse_scenarios_create.argtypes = [vp, i64, c_double_p, C.POINTER(vp)]
    L.peuqse_scenarios_destroy.argtypes = [vp]
    L.peuqse_scenarios_synthesize.argtypes = [vp, vp, vp, i64, vp, vp]
    L.peuqse_scenarios_set_observed.argtypes = [vp, vp, vp, vp]
    L.peuqse_scenarios_logp.argtypes = [vp, vp, vp, i64, vp, vp, vp, vp]
    L.peuqse_mh_run_scenarios.argtypes = [vp, vp, C.POINTER(MHSpec), i64, vp, vp, vp, vp, vp, vp, vp, vp, vp]
    L.peuqse_scenarios_stats.argtypes = [i32, i32, dbl, i64, i64, i64, i64, vp, vp, vp, vp, vp]
    L.peuqse_scenarios_kl.argtypes = [i32, i32, i64, i64, i64, i64, vp, vp, vp, i64, vp, vp, vp]
    L.peuqse_geweke_work_bytes.restype = i64
    L.peuqse_geweke_work_bytes.argtypes = [i64, i64, i32]
    L.peuqse_geweke.argtypes = [vp, i64, i64, vp, vp, i32, vp, i32, vp, vp, vp]
    L.peuqse_act_concat.argtypes = [vp, i64, i64, i32, vp, vp, i64, vp]
    for name in EXPORTS:
        if name not in ("peuqse_last_error", "peuqse_version", "peuqse_sweep_work_bytes", "peuqse_ess_work_bytes",
                        "peuqse_stats_work_bytes", "peuqse_peer_max_bytes", "peuqse_geweke_work_bytes"):
            getattr(L, name).restype = C.c_int
    _lib = L
    return L


def check(rc):
    if rc != 0:
        msg = load().peuqse_last_error().decode(errors="replace")
        raise PeuqseError(f"libpeuqse_b200 error {rc}: {msg}")


def ptr(t):
    """Device pointer of a torch tensor (or None)."""
    if t is None:
        return None
    assert t.is_cuda and t.is_contiguous(), "device tensors must be contiguous CUDA tensors"
    return C.c_void_p(t.data_ptr())


def current_stream():
    import torch
    return C.c_void_p(torch.cuda.current_stream().cuda_stream)
