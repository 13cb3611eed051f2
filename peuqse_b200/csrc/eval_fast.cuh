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
