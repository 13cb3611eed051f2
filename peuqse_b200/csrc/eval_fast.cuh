// "Fast" evaluation for small problems (single MVN_DIAG response with a compile-time bound on the number of
// observations, diagonal prior): every constant is a kernel-parameter (constant-bank) operand, every loop is
// unrolled, and the prior is split per axis so that callers can hoist the terms of axes that do not change.
// Arithmetic and summation order are the same as eval.cuh (and therefore as scipy's logpdf).
#pragma once
#include "common.cuh"
#include "eval.cuh"

template <int MODEL, int KT>
struct FastSim;

template <int KT>
struct FastSim<PEUQSE_MODEL_QUADRATIC, KT> {
    template <int J>
    __device__ __forceinline__ static double at(const DevProblem& P, const double (&th)[KT]) {
        const double d = P.fmc[J] - th[0];
        return d * d + th[1];
    }
};

template <int KT>
struct FastSim<PEUQSE_MODEL_LINEAR, KT> {
    template <int J>
    __device__ __forceinline__ static double at(const DevProblem& P, const double (&th)[KT]) {
        double acc = P.fmc[J * PQ_MAXK] * th[0];
#pragma unroll
        for (int d = 1; d < KT; ++d)
            if (d < P.k) acc = acc + P.fmc[J * PQ_MAXK + d] * th[d];
        return acc;
    }
};

// Row/column split of the built-in models for grid sweeps: everything that depends only on the lower axes of a
// grid row is computed once per row (row_pre), the last-axis part per point (at_last).  Same operation order as at().
template <int MODEL, int KT>
struct FastSplit;

template <int KT>
struct FastSplit<PEUQSE_MODEL_QUADRATIC, KT> {           // (x_j - a)^2 + b, a = th[0] (row), b = th[1] (last axis)
    template <int J>
    __device__ __forceinline__ static double row_pre(const DevProblem& P, const double (&th)[KT]) {
        const double d = P.fmc[J] - th[0];
        return d * d;
    }
    template <int J>
    __device__ __forceinline__ static double at_last(const DevProblem& P, double pre, double thL) { return pre + thL; }
};

template <int KT>
struct FastSplit<PEUQSE_MODEL_LINEAR, KT> {              // ((Phi_j0 th_0 + Phi_j1 th_1) + ...) + Phi_jL th_L
    template <int J>
    __device__ __forceinline__ static double row_pre(const DevProblem& P, const double (&th)[KT]) {
        if (KT == 1) return 0.0;
        double acc = P.fmc[J * PQ_MAXK] * th[0];
#pragma unroll
        for (int d = 1; d < KT - 1; ++d) acc = acc + P.fmc[J * PQ_MAXK + d] * th[d];
        return acc;
    }
    template <int J>
    __device__ __forceinline__ static double at_last(const DevProblem& P, double pre, double thL) {
        const double t = P.fmc[J * PQ_MAXK + (KT - 1)] * thL;
        return (KT == 1) ? t : (pre + t);
    }
};

template <int MODEL, int KT, int NT, int J>
struct FastRowPre {
    __device__ __forceinline__ static void fill(const DevProblem& P, const double (&th)[KT], double* pre) {
        pre[J] = FastSplit<MODEL, KT>::template row_pre<J>(P, th);
        if constexpr (J + 1 < NT) FastRowPre<MODEL, KT, NT, J + 1>::fill(P, th, pre);
    }
};

template <int MODEL, int KT, int NT, bool EXACT, int J>
struct FastQLast {
    __device__ __forceinline__ static void fill(const DevProblem& P, const double (&pre)[NT], double thL, double (&q)[NT]) {
        if (EXACT || J < P.fast_n) {
            const double s = FastSplit<MODEL, KT>::template at_last<J>(P, pre[J], thL);
            const double d = P.fy[J] - s;                       // x - mean (IP.py:2960)
            const double t = d * P.fw[J];
            q[J] = t * t;
        } else {
            q[J] = 0.0;
        }
        if constexpr (J + 1 < NT) FastQLast<MODEL, KT, NT, EXACT, J + 1>::fill(P, pre, thL, q);
    }
};

template <int MODEL, int KT, int NT, bool EXACT, int J>
struct FastQ {
    __device__ __forceinline__ static void fill(const DevProblem& P, const double (&th)[KT], double (&q)[NT]) {
        if (EXACT || J < P.fast_n) {
            const double s = FastSim<MODEL, KT>::template at<J>(P, th);
            const double d = P.fy[J] - s;                       // x - mean (IP.py:2960)
            const double t = d * P.fw[J];
            q[J] = t * t;
        } else {
            q[J] = 0.0;
        }
        if constexpr (J + 1 < NT) FastQ<MODEL, KT, NT, EXACT, J + 1>::fill(P, th, q);
    }
};

// logLikelihood of the single diagonal response; NT = compile-time capacity, EXACT: n == NT
template <int MODEL, int KT, int NT, bool EXACT>
__device__ __forceinline__ double fast_loglik(const DevProblem& P, const double (&th)[KT]) {
    double q[NT];
    FastQ<MODEL, KT, NT, EXACT, 0>::fill(P, th, q);
    const double maha = ordered_sum_static<NT>(EXACT ? NT : P.fast_n, q);
    const double ll = -0.5 * (P.resp_c[0] + maha);
    return (ll != ll) ? pq_neg_inf() : ll;                      // NaN simulated value -> -inf (IP.py:2880-2882)
}

// one axis of the prior: bounds test and the squared whitened deviation of parameter d
__device__ __forceinline__ void prior_axis(const DevProblem& P, int d, double th, bool& ok, double& qn) {
    ok = true;
    if (P.has_ub) ok = ok && (th <= P.ub[d]);
    if (P.has_lb) ok = ok && (th >= P.lb[d]);
    qn = 0.0;
    if (P.gauss[d]) {
        const double z = (P.scaling_on ? th / P.scale[d] : th) - P.mus[d];
        const double t = z * P.pwn[d];
        qn = t * t;
    }
}

template <int KT>
__device__ __forceinline__ double prior_combine(const DevProblem& P, bool ok, const double (&qn)[KT]) {
    if (!ok) return pq_neg_inf();
    if (P.prior_mode == PEUQSE_PRIOR_NONE) return 0.0;
    double q[KT];
#pragma unroll
    for (int j = 0; j < KT; ++j) {
        q[j] = 0.0;
        if (j < P.kp) {
            double v = qn[j];
            if (!P.prior_identity) {
                const int g = P.gidx[j];
#pragma unroll
                for (int d = 0; d < KT; ++d)
                    if (d == g) v = qn[d];
            }
            q[j] = v;
        }
    }
    return -0.5 * (P.prior_c + ordered_sum_static<KT>(P.kp, q));
}
