// Full-covariance likelihood (kernel K5 of SURVEY section 2): for a block of B simulated response rows,
//   maha_b = || (y - sim_b) @ U ||^2,   ll_b += -0.5*(rank*log(2pi) + log_pdet + maha_b)
// with U = u*sqrt(1/s) from scipy's _PSD (scipy/stats/_multivariate.py:208-236, 590-591) hoisted to set-up.
// Replaces one multivariate_normal.logpdf (an n x n eigendecomposition) per point (IP.py:2960).
#include <cstdint>
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

// ---------------------------------------------------------------------------------------------------
// v1: tensor-core contraction.  T[B,n] = Rsd[B,n] @ U[n,n] with DMMA (mma.sync.m8n8k4 f64 -- there is no tcgen05 kind
// for fp64), fused with the row-wise sum of squares so that T is never written: every CTA owns a 64 x 64 tile of T,
// loops over K in steps of 16 with register-prefetched, double-buffered shared-memory tiles (the residual y - sim
// is formed while staging A), squares its accumulators and writes one partial row sum per (column tile, row).
// A second tiny kernel adds the partials in fixed order: ll_b += -0.5*(c + sum).
// ---------------------------------------------------------------------------------------------------
#define QD_BM 64
#define QD_BN 64
#define QD_BK 16
#define QD_LDA (QD_BK + 4)     // padded leading dimensions: operand fragment loads hit every bank exactly twice
#define QD_LDB (QD_BN + 8)

__device__ __forceinline__ void dmma884(double& c0, double& c1, double a, double b) {
    asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};"
                 : "+d"(c0), "+d"(c1)
                 : "d"(a), "d"(b));
}

#define QD_CHUNKS 4    // the K range is always accumulated as 4 separately rounded chunk sums, combined ((c0+c1)+c2)+c3

// row-wise sum of squares of a CTA's 64 x 64 tile held in accumulator-fragment order -> partial[blockIdx.x][row]
__device__ __forceinline__ void qd_epilogue(const double (&acc)[4][2][2], long long bm, long long B, double* __restrict__ partial) {
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int wm = warp >> 2, wn = warp & 3;
    __shared__ double s_row[4][QD_BM];
#pragma unroll
    for (int i = 0; i < 4; ++i) {
        double s = 0.0;
#pragma unroll
        for (int j = 0; j < 2; ++j) s = fma(acc[i][j][0], acc[i][j][0], fma(acc[i][j][1], acc[i][j][1], s));
        s += __shfl_xor_sync(0xffffffffu, s, 1);                   // the 4 lanes of a row group hold different columns
        s += __shfl_xor_sync(0xffffffffu, s, 2);
        if ((lane & 3) == 0) s_row[wn][wm * 32 + i * 8 + (lane >> 2)] = s;
    }
    __syncthreads();
    if (tid < QD_BM) {
        const long long row = bm + tid;
        if (row < B) partial[(size_t)blockIdx.x * B + row] = ((s_row[0][tid] + s_row[1][tid]) + s_row[2][tid]) + s_row[3][tid];
    }
}

// SPLIT = false: one CTA per tile walks the four K chunks one after the other.  SPLIT = true (small batches: fewer tiles
// than SMs, and a 256-step K loop is pure latency): blockIdx.z is the chunk, the chunk's accumulators go to
// frag[chunk][tile][16][256] and k_quadform_split_finish combines them -- the same additions in the same order, so a
// row's value does not depend on the size of the batch it is evaluated in.
template <bool SPLIT, bool VEC>
__global__ void __launch_bounds__(256, 3) k_quadform_dmma(const double* __restrict__ sim, int ld_sim, int off,
                                                      const int* __restrict__ src, const double* __restrict__ y,
                                                      const double* __restrict__ U, int n, long long B,
                                                      double* __restrict__ partial /* [gridDim.x][B] */,
                                                      double* __restrict__ frag) {
    __shared__ __align__(16) double As[2][QD_BM][QD_LDA];
    __shared__ __align__(16) double Bs[2][QD_BK][QD_LDB];
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int wm = warp >> 2, wn = warp & 3;                   // 2 x 4 warps: each owns 32 x 16 of the 64 x 64 tile
    const long long bm = (long long)blockIdx.y * QD_BM;
    const int bn = blockIdx.x * QD_BN;
    // staging assignments: A tile 64 x 16 (thread -> row ar, 4 consecutive k), B tile 16 x 64 (thread -> row br, 4 consecutive cols)
    const int ar = tid >> 2, ak = (tid & 3) * 4;
    const int br = tid >> 4, bc = (tid & 15) * 4;
    const long long arow = bm + ar;
    double ra[4], rb[4];
    auto load_tiles = [&](int k0) {
        if (VEC) {
            // simulated values in place (src is the identity), n a multiple of 4 and every row start 16-byte aligned: the four
            // consecutive elements of a thread are two 16-byte loads, no index gather in front of them
            const int kk = k0 + ak, kb = k0 + br, cc = bn + bc;
            ra[0] = ra[1] = ra[2] = ra[3] = 0.0;
            rb[0] = rb[1] = rb[2] = rb[3] = 0.0;
            if (arow < B && kk < n) {
                const double2* sp = reinterpret_cast<const double2*>(sim + arow * ld_sim + off + kk);
                const double2* yp = reinterpret_cast<const double2*>(y + off + kk);
                const double2 s0 = sp[0], s1 = sp[1], y0 = __ldg(yp), y1 = __ldg(yp + 1);
                ra[0] = y0.x - s0.x; ra[1] = y0.y - s0.y; ra[2] = y1.x - s1.x; ra[3] = y1.y - s1.y;   // x - mean (IP.py:2960)
            }
            if (kb < n && cc < n) {
                const double2* up = reinterpret_cast<const double2*>(U + (size_t)kb * n + cc);
                const double2 u0 = __ldg(up), u1 = __ldg(up + 1);
                rb[0] = u0.x; rb[1] = u0.y; rb[2] = u1.x; rb[3] = u1.y;
            }
            return;
        }
#pragma unroll
        for (int q = 0; q < 4; ++q) {
            const int kk = k0 + ak + q;
            ra[q] = (arow < B && kk < n) ? (y[off + kk] - sim[arow * ld_sim + src[off + kk]]) : 0.0;   // x - mean (IP.py:2960)
            const int kb = k0 + br, cc = bn + bc + q;
            rb[q] = (kb < n && cc < n) ? __ldg(U + (size_t)kb * n + cc) : 0.0;
        }
    };
    auto store_tiles = [&](int st) {
        // (ar * QD_LDA + ak) and (br * QD_LDB + bc) are multiples of 2 doubles: two 16-byte stores per operand
        double2* ap = reinterpret_cast<double2*>(&As[st][ar][ak]);
        double2* bp = reinterpret_cast<double2*>(&Bs[st][br][bc]);
        ap[0] = make_double2(ra[0], ra[1]); ap[1] = make_double2(ra[2], ra[3]);
        bp[0] = make_double2(rb[0], rb[1]); bp[1] = make_double2(rb[2], rb[3]);
    };
    const int nk = (n + QD_BK - 1) / QD_BK;
    double tot[4][2][2];
    const int c_first = SPLIT ? (int)blockIdx.z : 0, c_last = SPLIT ? (int)blockIdx.z + 1 : QD_CHUNKS;
    for (int ch = c_first; ch < c_last; ++ch) {
        const int kt0 = (int)((long long)nk * ch / QD_CHUNKS), kt1 = (int)((long long)nk * (ch + 1) / QD_CHUNKS);
        double acc[4][2][2];
#pragma unroll
        for (int i = 0; i < 4; ++i)
#pragma unroll
            for (int j = 0; j < 2; ++j) acc[i][j][0] = acc[i][j][1] = 0.0;
        if (kt0 < kt1) {
            load_tiles(kt0 * QD_BK);
            store_tiles(0);
            __syncthreads();
            for (int kt = kt0; kt < kt1; ++kt) {
                const int st = (kt - kt0) & 1;
                if (kt + 1 < kt1) load_tiles((kt + 1) * QD_BK);      // global loads of the next tile fly during the MMAs
#pragma unroll
                for (int ks = 0; ks < QD_BK; ks += 4) {
                    double af[4], bf[2];
#pragma unroll
                    for (int i = 0; i < 4; ++i) af[i] = As[st][wm * 32 + i * 8 + (lane >> 2)][ks + (lane & 3)];
#pragma unroll
                    for (int j = 0; j < 2; ++j) bf[j] = Bs[st][ks + (lane & 3)][wn * 16 + j * 8 + (lane >> 2)];
#pragma unroll
                    for (int i = 0; i < 4; ++i)
#pragma unroll
                        for (int j = 0; j < 2; ++j) dmma884(acc[i][j][0], acc[i][j][1], af[i], bf[j]);
                }
                if (kt + 1 < kt1) store_tiles(st ^ 1);
                __syncthreads();
            }
        }
        if (SPLIT) {
            double* f = frag + (((size_t)ch * gridDim.y + blockIdx.y) * gridDim.x + blockIdx.x) * (16 * 256) + tid;
#pragma unroll
            for (int i = 0; i < 4; ++i)
#pragma unroll
                for (int j = 0; j < 2; ++j) {
                    f[((i * 2 + j) * 2 + 0) * 256] = acc[i][j][0];
                    f[((i * 2 + j) * 2 + 1) * 256] = acc[i][j][1];
                }
        } else {
#pragma unroll
            for (int i = 0; i < 4; ++i)
#pragma unroll
                for (int j = 0; j < 2; ++j) {
                    tot[i][j][0] = (ch == 0) ? acc[i][j][0] : tot[i][j][0] + acc[i][j][0];
                    tot[i][j][1] = (ch == 0) ? acc[i][j][1] : tot[i][j][1] + acc[i][j][1];
                }
        }
    }
    if (!SPLIT) qd_epilogue(tot, bm, B, partial);
}

__global__ void __launch_bounds__(256) k_quadform_split_finish(const double* __restrict__ frag, long long B,
                                                              double* __restrict__ partial) {
    const size_t tile = (size_t)blockIdx.y * gridDim.x + blockIdx.x, tiles = (size_t)gridDim.x * gridDim.y;
    double tot[4][2][2];
#pragma unroll
    for (int ch = 0; ch < QD_CHUNKS; ++ch) {
        const double* f = frag + ((size_t)ch * tiles + tile) * (16 * 256) + threadIdx.x;
#pragma unroll
        for (int i = 0; i < 4; ++i)
#pragma unroll
            for (int j = 0; j < 2; ++j) {
                const double a0 = f[((i * 2 + j) * 2 + 0) * 256], a1 = f[((i * 2 + j) * 2 + 1) * 256];
                tot[i][j][0] = (ch == 0) ? a0 : tot[i][j][0] + a0;
                tot[i][j][1] = (ch == 0) ? a1 : tot[i][j][1] + a1;
            }
    }
    qd_epilogue(tot, (long long)blockIdx.y * QD_BM, B, partial);
}

__global__ void k_quadform_reduce(const double* __restrict__ partial, int ntn, long long B, double c, double* __restrict__ ll) {
    const long long b = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (b >= B) return;
    double m = 0.0;
    for (int t = 0; t < ntn; ++t) m += partial[(size_t)t * B + b];     // fixed order
    ll[b] = ll[b] + (-0.5 * (c + m));
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
            const int ntn = (n + QD_BN - 1) / QD_BN;
            double* part = nullptr;
            cudaError_t e = cudaMallocAsync((void**)&part, (size_t)ntn * B * sizeof(double), st);
            if (e != cudaSuccess) { rc = pq_set_cuda_error(e, "cudaMallocAsync(quadform partials)", __LINE__); break; }
            const dim3 grid((unsigned)ntn, (unsigned)((B + QD_BM - 1) / QD_BM));
            // fewer tiles than two per SM and a K loop long enough to be worth cutting: one CTA per (tile, K chunk)
            const bool split = (long long)grid.x * grid.y < 2LL * p->sm_count && n >= 16 * QD_BK;
            const bool vec = p->src_identity && n % 4 == 0 && p->d.n_total % 2 == 0 && p->d.resp_off[r] % 2 == 0 &&
                             ((uintptr_t)d_sim | (uintptr_t)(p->d.y) | (uintptr_t)(p->d.U[r])) % 16 == 0;
            if (split) {
                double* frag = nullptr;
                e = cudaMallocAsync((void**)&frag, (size_t)QD_CHUNKS * grid.x * grid.y * 16 * 256 * sizeof(double), st);
                if (e != cudaSuccess) { cudaFreeAsync(part, st); rc = pq_set_cuda_error(e, "cudaMallocAsync(quadform fragments)", __LINE__); break; }
                auto kern = vec ? k_quadform_dmma<true, true> : k_quadform_dmma<true, false>;
                kern<<<dim3(grid.x, grid.y, QD_CHUNKS), 256, 0, st>>>(d_sim, p->d.n_total, p->d.resp_off[r], p->d.src, p->d.y, p->d.U[r], n, B,
                                                                     part, frag);
                k_quadform_split_finish<<<grid, 256, 0, st>>>(frag, B, part);
                cudaFreeAsync(frag, st);
            } else {
                auto kern = vec ? k_quadform_dmma<false, true> : k_quadform_dmma<false, false>;
                kern<<<grid, 256, 0, st>>>(d_sim, p->d.n_total, p->d.resp_off[r], p->d.src, p->d.y, p->d.U[r], n, B, part, nullptr);
            }
            k_quadform_reduce<<<(unsigned)((B + 255) / 256), 256, 0, st>>>(part, ntn, B, p->d.resp_c[r], ll);
            e = cudaGetLastError();
            cudaFreeAsync(part, st);
            if (e != cudaSuccess) rc = pq_set_cuda_error(e, "k_quadform_dmma", __LINE__);
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
