// What the persistent Metropolis-Hastings kernels (mh.cu) and the lock-step driver (mh_steps.cu) share: the run description, the
// 'MAP_finding' step-size adaptation, the proposal increment and the draws of one (chain, step).
#pragma once
#include "common.cuh"
#include "philox.cuh"

struct DevMH {
    long long n_chains, n_steps, burn_in, first_step, chain_offset;
    unsigned long long seed;
    double modulate;
    int k, thin, chol_diag, map_finding;
    double chol[PQ_MAXK * PQ_MAXK];
};

// mcmc_mode 'MAP_finding' (IP.py:2631-2635, 2688-2713): every step draws a modulation m = u + 0.5 and proposes with the
// covariance Q * dyn * m * step, i.e. the increment of the unbiased mode times sqrt(dyn * m).  After every hundredth step the
// previous hundred (m, proposal log-posterior) pairs with a finite log-posterior are fitted by a straight line
// (np.polynomial.polynomial.polyfit, degree 1): positive slope -> dyn *= 1.05 while dyn < 10, negative -> dyn *= 0.95 while
// dyn > 0.1.  Only the SIGN of the slope matters: n sum(xy) - sum(x) sum(y) over the window, with y taken relative to the
// window's first finite value (no cancellation against |logP| ~ 1e4).  The first window starts with the entry (0, 0): index 0
// of the reference's zero-initialised history arrays lies inside its slice [i-100:i] at i = 100.
struct MapFind {
    double dyn, n, sx, sy, sxy, y0;
    __device__ __forceinline__ void init() { dyn = 1.0; n = 1.0; sx = 0.0; sy = 0.0; sxy = 0.0; y0 = 0.0; }
    __device__ __forceinline__ double scale(double m) const { return sqrt(dyn * m); }
    __device__ __forceinline__ void after_step(long long i, double m, double logp_drawn) {
        if (i % 100 == 0) {
            const double num = n * sxy - sx * sy;
            if (num > 0.0) { if (dyn < 10.0) dyn = dyn * 1.05; }
            else if (num < 0.0) { if (dyn > 0.1) dyn = dyn * 0.95; }
            n = 0.0; sx = 0.0; sy = 0.0; sxy = 0.0;
        }
        if (fabs(logp_drawn) < pq_pos_inf()) {                     // finite (NaN compares false)
            if (n == 0.0) y0 = logp_drawn;
            const double y = logp_drawn - y0;
            n += 1.0; sx += m; sy += y; sxy = fma(m, y, sxy);
        }
    }
};


// proposal increment delta = L z (L lower-triangular Cholesky factor of Q_covmat*step, row-major)
template <int KT>
__device__ __forceinline__ void make_delta(const DevMH& M, const double (&z)[KT], double (&delta)[KT]) {
#pragma unroll
    for (int d = 0; d < KT; ++d) {
        double a = 0.0;
        if (d < M.k) {
            if (M.chol_diag) {
                a = M.chol[d * M.k + d] * z[d];
            } else {
#pragma unroll
                for (int e = 0; e < KT; ++e)
                    if (e <= d) a = fma(M.chol[d * M.k + e], z[e], a);
            }
        }
        delta[d] = a;
    }
}

// draws for (chain, step): normals from blocks 8*step + j, accept uniform from block 8*step + 7
template <int KT>
__device__ __forceinline__ void draws_thread(const DevMH& M, unsigned long long chain, unsigned long long step,
                                             double (&z)[KT], double& u, double& um) {
#pragma unroll
    for (int b = 0; b < (KT + 1) / 2; ++b) {
        double z0 = 0.0, z1 = 0.0;
        if (2 * b < M.k) philox_normal_pair(philox_block(M.seed, PQ_STREAM_MH, chain, (uint32_t)(8ULL * step + b)), z0, z1);
        z[2 * b] = z0;
        if (2 * b + 1 < KT) z[2 * b + 1] = z1;
    }
    const Philox4 r = philox_block(M.seed, PQ_STREAM_MH, chain, (uint32_t)(8ULL * step + 7ULL));
    u = philox_u53(r.x, r.y);
    um = philox_u53(r.z, r.w);                                     // the other half of the block: 'MAP_finding' modulation
}

