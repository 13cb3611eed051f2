// Full-covariance likelihood (kernel K5 of SURVEY section 2): for a block of B simulated response rows,
//   maha_b = || (y - sim_b) @ U ||^2,   ll_b += -0.5*(rank*log(2pi) + log_pdet + maha_b)
// with U = u*sqrt(1/s) from scipy's _PSD (scipy/stats/_multivariate.py:208-236, 590-591) hoisted to set-up.
// Replaces one multivariate_normal.logpdf (an n x n eigendecomposition) per point (IP.py:2960).
#include "common.cuh"
#include "eval.cuh"
#include "launch.cuh"

int pq_logp_external(const peuqse_problem* p, const double* th, const double* sim, long long B, double* logp, double* lpr,
                     double* llk, cudaStream_t st);

#define QF_ROWS 8      // rows of the residual block per CTA
#define QF_THREADS 256

// v0 contraction on the fp64 FMA pipe: a CTA owns QF_ROWS residual rows (staged in shared memory) and its
// threads stride over the columns of U; every U element loaded feeds QF_ROWS FMAs.
__global__ void __launch_bounds__(QF_THREADS) k_quadform(const double* __restrict__ sim, int ld_sim, int off,
                                                        const int* __restrict__ src, const double* __restrict__ y,
                                                        const double* __restrict__ U, int n, double c, long long B,
                                                        double* __restrict__ ll) {
    extern __shared__ double sd[];   // [QF_ROWS][n] residuals
    __shared__ double s_red[QF_THREADS / 32][QF_ROWS];
    const long long b0 = (long long)blockIdx.x * QF_ROWS;
    for (int e = threadIdx.x; e < QF_ROWS * n; e += blockDim.x) {
        const int r = e / n, i = e - r * n;
        const long long b = b0 + r;
        sd[e] = (b < B) ? (y[off + i] - sim[b * ld_sim + src[off + i]]) : 0.0;   // x - mean (IP.py:2960)
    }
    __syncthreads();
    double acc[QF_ROWS];
#pragma unroll
    for (int r = 0; r < QF_ROWS; ++r) acc[r] = 0.0;
    for (int j = threadIdx.x; j < n; j += blockDim.x) {
        double t[QF_ROWS];
#pragma unroll
        for (int r = 0; r < QF_ROWS; ++r) t[r] = 0.0;
        for (int i = 0; i < n; ++i) {
            const double u = __ldg(U + (size_t)i * n + j);
#pragma unroll
            for (int r = 0; r < QF_ROWS; ++r) t[r] = fma(sd[r * n + i], u, t[r]);
        }
#pragma unroll
        for (int r = 0; r < QF_ROWS; ++r) acc[r] = fma(t[r], t[r], acc[r]);
    }
#pragma unroll
    for (int r = 0; r < QF_ROWS; ++r) {
        const double v = warp_sum(acc[r]);
        if ((threadIdx.x & 31) == 0) s_red[threadIdx.x >> 5][r] = v;
    }
    __syncthreads();
    if (threadIdx.x < QF_ROWS) {
        double m = 0.0;
        for (int w = 0; w < QF_THREADS / 32; ++w) m += s_red[w][threadIdx.x];
        const long long b = b0 + threadIdx.x;
        if (b < B) ll[b] = ll[b] + (-0.5 * (c + m));
    }
}

__global__ void k_fullcov_finish(long long B, const double* __restrict__ lp, double* __restrict__ ll,
                                 double* __restrict__ logp) {
    const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= B) return;
    double l = ll[i];
    if (l != l) l = pq_neg_inf();            // NaN simulated values -> -inf (IP.py:2880-2882)
    ll[i] = l;
    if (logp) logp[i] = combine_logp(lp[i], l);
}

int pq_fullcov_logp(const peuqse_problem* p, const double* d_theta, const double* d_sim, long long B, double* d_logp,
                    double* d_logprior, double* d_loglik, cudaStream_t st) {
    double* ll = d_loglik;
    double* lp = d_logprior;
    double* tmp = nullptr;
    const bool need_lp = d_logp != nullptr || d_logprior != nullptr;
    if (!ll || (need_lp && !lp)) {
        PQ_CUDA_OK(cudaMallocAsync((void**)&tmp, (size_t)2 * B * sizeof(double), st));
        if (!ll) ll = tmp;
        if (need_lp && !lp) lp = tmp + B;
    }
    // prior and the diagonal / per-value responses; full responses contribute 0 here
    int rc = pq_logp_external(p, need_lp ? d_theta : nullptr, d_sim, B, nullptr, need_lp ? lp : nullptr, ll, st);
    if (rc == PEUQSE_OK) {
        for (int r = 0; r < p->d.n_resp && rc == PEUQSE_OK; ++r) {
            if (p->d.resp_mode[r] != PEUQSE_RESP_MVN_FULL) continue;
            const int n = p->d.resp_n[r];
            const size_t smem = (size_t)QF_ROWS * n * sizeof(double);
            if (smem > (size_t)p->max_smem_optin) { rc = pq_set_error(PEUQSE_ERR_INVALID, "full-covariance response too long for the v0 contraction"); break; }
            cudaError_t e = pq_allow_smem((const void*)k_quadform, smem);
            if (e != cudaSuccess) { rc = pq_set_cuda_error(e, "cudaFuncSetAttribute", __LINE__); break; }
            const unsigned blocks = (unsigned)((B + QF_ROWS - 1) / QF_ROWS);
            k_quadform<<<blocks, QF_THREADS, smem, st>>>(d_sim, p->d.n_total, p->d.resp_off[r], p->d.src, p->d.y, p->d.U[r], n,
                                                       p->d.resp_c[r], B, ll);
            e = cudaGetLastError();
            if (e != cudaSuccess) rc = pq_set_cuda_error(e, "k_quadform", __LINE__);
        }
        if (rc == PEUQSE_OK) {
            k_fullcov_finish<<<(unsigned)((B + 255) / 256), 256, 0, st>>>(B, lp, ll, d_logp);
            cudaError_t e = cudaGetLastError();
            if (e != cudaSuccess) rc = pq_set_cuda_error(e, "k_fullcov_finish", __LINE__);
        }
    }
    if (tmp) cudaFreeAsync(tmp, st);
    return rc;
}
