// Specialised grid-sweep kernels (the C2 hot path): the bulk of combinationGenerator's list
// (CombinationGeneratorModule.py:40-81) is, for the last axis L with radius r, the set
//     g = M_L + b*2r + j,   b = index of a lower-axes combination (a "row"),  j in [0, 2r)
// so consecutive g walk along the last axis.  A CTA owns a fixed chunk of j's and strides over rows:
//   - every thread keeps its last-axis values theta_L(j) and their prior terms (incl. the fp64 division of
//     IP.py:2767) in registers for the whole kernel,
//   - the lower-axis point, its prior terms and every model sub-expression that depends on it are
//     block-uniform per row and are computed once per row (the 8 unrolled evaluations share them),
//   - per point only the last-axis part of the model, the residuals and the two reductions remain;
//     stores / loads of logP are fully coalesced along the row.
// Same arithmetic and summation order as the generic kernels (sweep.cu / eval.cuh): results are bit-identical.
#include "common.cuh"
#include "eval.cuh"
#include "eval_fast.cuh"
#include "launch.cuh"
#include "sweep_common.cuh"

#define GR_THREADS 256
#define GR_JPT 8
#define GR_CH (GR_THREADS * GR_JPT)

struct GridTiling {
    long long ML;          // first global index of the bulk of the last axis = prod_{d<L}(2 r_d + 1)
    long long r2;          // 2 r_L
    long long b_lo, b_hi;  // rows intersecting [first, first+count)
    int nchunks, nrg;      // chunks along j, row groups
};

__device__ __forceinline__ double last_axis_value(const DevSweep& S, int L, long long j) {
    const double step = S.interval[L] * (double)((j >> 1) + 1LL);          // (BaseFactor*i), CombinationGeneratorModule.py:59,71
    return (j & 1LL) ? (S.center[L] + step) : (S.center[L] - step);
}

// block-level merge of per-thread bests into *best_out (read-modify-write: several kernels feed one pass)
__device__ __forceinline__ void block_merge_best(Best best, Best* s_best, bool* s_last, Best* partial, unsigned int* ticket,
                                                 peuqse_best* best_out) {
    best = warp_best(best);
    if ((threadIdx.x & 31) == 0) s_best[threadIdx.x >> 5] = best;
    __syncthreads();
    if (threadIdx.x < 32) {
        Best b = (threadIdx.x < (blockDim.x >> 5)) ? s_best[threadIdx.x] : best_none();
        b = warp_best(b);
        if (threadIdx.x == 0) {
            partial[blockIdx.x] = b;
            __threadfence();
            const unsigned int t = atomicAdd(ticket, 1u);
            *s_last = (t == gridDim.x - 1);
        }
    }
    __syncthreads();
    if (*s_last && threadIdx.x < 32) {
        __threadfence();
        Best b = best_none();
        for (int j = threadIdx.x; j < (int)gridDim.x; j += 32) {
            Best o;
            o.v = __ldcg(&partial[j].v);
            o.i = __ldcg(&partial[j].i);
            b = best_merge(b, o);
        }
        b = warp_best(b);
        if (threadIdx.x == 0) {
            Best cur;
            cur.v = best_out->logp;
            cur.i = best_out->index < 0 ? 0x7fffffffffffffffLL : best_out->index;
            b = best_merge(cur, b);
            best_out->logp = b.v;
            best_out->index = (b.i == 0x7fffffffffffffffLL) ? -1 : b.i;
            *ticket = 0;
        }
    }
}

// ---------------------------------------------------------------------------------------------
// pass 1
// ---------------------------------------------------------------------------------------------
template <int MODEL, int KT, int NT, bool EXACT>
__global__ void __launch_bounds__(GR_THREADS) k_grid_pass1(const __grid_constant__ DevProblem P, const __grid_constant__ DevSweep S,
                                                           const __grid_constant__ GridTiling T, long long first, long long count,
                                                           double* __restrict__ logp_out, peuqse_best* __restrict__ best_out,
                                                           Best* __restrict__ partial, unsigned int* __restrict__ ticket) {
    __shared__ Best s_best[GR_THREADS / 32];
    __shared__ bool s_last;
    constexpr int L = KT - 1;
    const int chunk = blockIdx.x % T.nchunks, rg = blockIdx.x / T.nchunks;
    const long long j0 = (long long)chunk * GR_CH + threadIdx.x;
    double thL[GR_JPT], qL[GR_JPT];
    bool okL[GR_JPT];
#pragma unroll
    for (int m = 0; m < GR_JPT; ++m) {
        const long long j = j0 + m * GR_THREADS;
        thL[m] = last_axis_value(S, L, j);
        prior_axis(P, L, thL[m], okL[m], qL[m]);
        okL[m] = okL[m] && (j < T.r2);
    }
    Best best = best_none();
    for (long long b = T.b_lo + rg; b < T.b_hi; b += T.nrg) {
        double th[KT], qn[KT];
        grid_point<KT>(S, (unsigned long long)b, th, L);          // lower axes of this row (block-uniform)
        bool ok_low = true;
#pragma unroll
        for (int d = 0; d < KT; ++d) {
            qn[d] = 0.0;
            if (d < L) {
                bool okd;
                prior_axis(P, d, th[d], okd, qn[d]);
                ok_low = ok_low && okd;
            }
        }
        const long long gbase = T.ML + b * T.r2 + j0;
#pragma unroll
        for (int m = 0; m < GR_JPT; ++m) {
            const long long g = gbase + m * GR_THREADS;
            const long long i = g - first;
            if (j0 + m * GR_THREADS < T.r2 && i >= 0 && i < count) {
                th[L] = thL[m];
                qn[L] = qL[m];
                const double lp = prior_combine<KT>(P, ok_low && okL[m], qn);
                const double ll = fast_loglik<MODEL, KT, NT, EXACT>(P, th);
                const double v = combine_logp(lp, ll);              // IP.py:1649-1653
                if (logp_out) __stcs(logp_out + i, v);
                if (v > best.v) { best.v = v; best.i = g; }          // strict '>' in increasing g: first maximum
            }
        }
    }
    block_merge_best(best, s_best, &s_last, partial, ticket, best_out);
}

// ---------------------------------------------------------------------------------------------
// pass 2 (logP read back from HBM)
// ---------------------------------------------------------------------------------------------
__device__ __forceinline__ double blk_sum(double v, double* s_red) {
    v = warp_sum(v);
    __syncthreads();
    if ((threadIdx.x & 31) == 0) s_red[threadIdx.x >> 5] = v;
    __syncthreads();
    double t = 0.0;
    if (threadIdx.x < 32) {
        t = (threadIdx.x < (blockDim.x >> 5)) ? s_red[threadIdx.x] : 0.0;
        t = warp_sum(t);
    }
    return t;
}
__device__ __forceinline__ double blk_min(double v, double* s_red) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v = fmin(v, __shfl_xor_sync(0xffffffffu, v, o));
    __syncthreads();
    if ((threadIdx.x & 31) == 0) s_red[threadIdx.x >> 5] = v;
    __syncthreads();
    double t = pq_pos_inf();
    if (threadIdx.x < 32) {
        t = (threadIdx.x < (blockDim.x >> 5)) ? s_red[threadIdx.x] : pq_pos_inf();
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) t = fmin(t, __shfl_xor_sync(0xffffffffu, t, o));
    }
    return t;
}

template <int KT>
__global__ void __launch_bounds__(GR_THREADS) k_grid_pass2(const __grid_constant__ DevSweep S, const __grid_constant__ GridTiling T,
                                                           long long first, long long count, const double* __restrict__ logp_in,
                                                           const peuqse_best* __restrict__ best_in, double log_rel,
                                                           double* __restrict__ partial, double* __restrict__ acc_out,
                                                           unsigned int* __restrict__ ticket, long long* __restrict__ keep_idx,
                                                           double* __restrict__ keep_logp, long long keep_cap,
                                                           unsigned long long* __restrict__ keep_count) {
    __shared__ double s_red[GR_THREADS / 32];
    __shared__ bool s_last;
    constexpr int L = KT - 1;
    const long long best_index = best_in->index;
    if (best_index < 0) return;
    const double thr = best_in->logp + log_rel;                    // IP.py:3477
    double shift[KT];
    grid_point<KT>(S, (unsigned long long)best_index, shift, KT);  // moments about the MAP point
    const int chunk = blockIdx.x % T.nchunks, rg = blockIdx.x / T.nchunks;
    const long long j0 = (long long)chunk * GR_CH + threadIdx.x;
    double thL[GR_JPT];
#pragma unroll
    for (int m = 0; m < GR_JPT; ++m) thL[m] = last_axis_value(S, L, j0 + m * GR_THREADS);

    double aw = 0.0, awexp = 0.0, ank = 0.0;
    double s1[KT], s2[KT * (KT + 1) / 2], mn[KT], mx[KT];
#pragma unroll
    for (int d = 0; d < KT; ++d) { s1[d] = 0.0; mn[d] = pq_pos_inf(); mx[d] = pq_neg_inf(); }
#pragma unroll
    for (int e = 0; e < KT * (KT + 1) / 2; ++e) s2[e] = 0.0;

    for (long long b = T.b_lo + rg; b < T.b_hi; b += T.nrg) {
        double th[KT];
        grid_point<KT>(S, (unsigned long long)b, th, L);
        const long long gbase = T.ML + b * T.r2 + j0;
        double v[GR_JPT];
#pragma unroll
        for (int m = 0; m < GR_JPT; ++m) {                          // all loads of the row first (memory-level parallelism)
            const long long i = gbase + m * GR_THREADS - first;
            v[m] = pq_neg_inf();
            if (j0 + m * GR_THREADS < T.r2 && i >= 0 && i < count) v[m] = __ldcs(logp_in + i);
        }
#pragma unroll
        for (int m = 0; m < GR_JPT; ++m) {
            if (v[m] > thr) {                                       // IP.py:3481
                const double w = trunc(exp(v[m] - thr));            // int(np.exp(...)), IP.py:3485-3486
                th[L] = thL[m];
                aw += w;
                awexp = fma(w, exp(v[m]), awexp);
                ank += 1.0;
                int e = 0;
#pragma unroll
                for (int d = 0; d < KT; ++d) {
                    const double wd = w * (th[d] - shift[d]);
                    s1[d] += wd;
#pragma unroll
                    for (int d2 = d; d2 < KT; ++d2) { s2[e] = fma(wd, th[d2] - shift[d2], s2[e]); ++e; }
                    mn[d] = fmin(mn[d], th[d]);
                    mx[d] = fmax(mx[d], th[d]);
                }
                if (keep_idx) {
                    const unsigned long long slot = atomicAdd(keep_count, 1ULL);
                    if ((long long)slot < keep_cap) { keep_idx[slot] = gbase + m * GR_THREADS; keep_logp[slot] = v[m]; }
                }
            }
        }
    }
    double* mine = partial + (size_t)blockIdx.x * PQ_ACC_MAX;
    double t;
    t = blk_sum(aw, s_red);    if (threadIdx.x == 0) mine[PEUQSE_ACC_W] = t;
    t = blk_sum(awexp, s_red); if (threadIdx.x == 0) mine[PEUQSE_ACC_WEXP] = t;
    t = blk_sum(ank, s_red);   if (threadIdx.x == 0) mine[PEUQSE_ACC_NKEPT] = t;
    {
        int e = 0;
#pragma unroll
        for (int d = 0; d < KT; ++d) {
            t = blk_sum(s1[d], s_red);   if (threadIdx.x == 0) mine[PEUQSE_ACC_S1 + d] = t;
            t = blk_min(mn[d], s_red);   if (threadIdx.x == 0) mine[PEUQSE_ACC_MIN(KT) + d] = t;
            t = -blk_min(-mx[d], s_red); if (threadIdx.x == 0) mine[PEUQSE_ACC_MAX(KT) + d] = t;
#pragma unroll
            for (int d2 = d; d2 < KT; ++d2) {
                t = blk_sum(s2[e], s_red);
                if (threadIdx.x == 0) { mine[PEUQSE_ACC_S2(KT) + d * KT + d2] = t; mine[PEUQSE_ACC_S2(KT) + d2 * KT + d] = t; }
                ++e;
            }
        }
    }
    if (threadIdx.x == 0) {
        __threadfence();
        const unsigned int tk = atomicAdd(ticket, 1u);
        s_last = (tk == gridDim.x - 1);
    }
    __syncthreads();
    if (s_last) {
        __threadfence();
        constexpr int nacc = PEUQSE_ACC_DOUBLES(KT), mn0 = PEUQSE_ACC_MIN(KT), mx0 = PEUQSE_ACC_MAX(KT);
        for (int a = threadIdx.x; a < nacc; a += blockDim.x) {
            double r = acc_out[a];
            for (int b = 0; b < (int)gridDim.x; ++b) {             // fixed block order: deterministic
                const double x = __ldcg(partial + (size_t)b * PQ_ACC_MAX + a);
                if (a >= mx0) r = fmax(r, x);
                else if (a >= mn0) r = fmin(r, x);
                else r += x;
            }
            acc_out[a] = r;
        }
        if (threadIdx.x == 0) *ticket = 0;
    }
}

// ---------------------------------------------------------------------------------------------
// pass 3: marginal histograms without per-point atomics
//   lower axes: the bin is uniform per row -> one shared-memory atomic per warp per row;
//   last axis:  every thread owns fixed j's -> private column sums in registers, binned once at the end.
// ---------------------------------------------------------------------------------------------
__device__ __forceinline__ int hist_bin(const double* __restrict__ ed, int nbins, double x) {
    const double lo = ed[0], hi = ed[nbins];
    if (!(x >= lo && x <= hi)) return -1;
    int b = (int)(((x - lo) / (hi - lo)) * (double)nbins);         // numpy/lib/_histograms_impl.py
    if (b >= nbins) b = nbins - 1;
    if (b < 0) b = 0;
    if (x < ed[b]) --b;
    else if (b != nbins - 1 && x >= ed[b + 1]) ++b;
    return b;
}

template <int KT>
__global__ void __launch_bounds__(GR_THREADS) k_grid_hist(const __grid_constant__ DevSweep S, const __grid_constant__ GridTiling T,
                                                          long long first, long long count, const double* __restrict__ logp_in,
                                                          const peuqse_best* __restrict__ best_in, double log_rel,
                                                          const double* __restrict__ edges, int nbins,
                                                          unsigned long long* __restrict__ hist) {
    extern __shared__ unsigned long long sh[];                      // [KT][nbins]
    constexpr int L = KT - 1;
    if (best_in->index < 0) return;
    const double thr = best_in->logp + log_rel;
    for (int a = threadIdx.x; a < KT * nbins; a += blockDim.x) sh[a] = 0ULL;
    __syncthreads();
    const int chunk = blockIdx.x % T.nchunks, rg = blockIdx.x / T.nchunks;
    const long long j0 = (long long)chunk * GR_CH + threadIdx.x;
    int binL[GR_JPT];
    double colw[GR_JPT];
#pragma unroll
    for (int m = 0; m < GR_JPT; ++m) {
        binL[m] = hist_bin(edges + L * (nbins + 1), nbins, last_axis_value(S, L, j0 + m * GR_THREADS));
        colw[m] = 0.0;
    }
    for (long long b = T.b_lo + rg; b < T.b_hi; b += T.nrg) {
        const long long gbase = T.ML + b * T.r2 + j0;
        double v[GR_JPT];
#pragma unroll
        for (int m = 0; m < GR_JPT; ++m) {
            const long long i = gbase + m * GR_THREADS - first;
            v[m] = pq_neg_inf();
            if (j0 + m * GR_THREADS < T.r2 && i >= 0 && i < count) v[m] = __ldcs(logp_in + i);
        }
        double roww = 0.0;
#pragma unroll
        for (int m = 0; m < GR_JPT; ++m) {
            if (v[m] > thr) {
                const double w = trunc(exp(v[m] - thr));
                colw[m] += w;
                roww += w;
            }
        }
        if (KT > 1) {
            roww = warp_sum(roww);                                   // integers < 2^53: exact in fp64
            if ((threadIdx.x & 31) == 0 && roww > 0.0) {
                double th[KT];
                grid_point<KT>(S, (unsigned long long)b, th, L);
#pragma unroll
                for (int d = 0; d < L; ++d) {
                    const int bin = hist_bin(edges + d * (nbins + 1), nbins, th[d]);
                    if (bin >= 0) atomicAdd(sh + d * nbins + bin, (unsigned long long)roww);
                }
            }
        }
    }
#pragma unroll
    for (int m = 0; m < GR_JPT; ++m)
        if (binL[m] >= 0 && colw[m] > 0.0) atomicAdd(sh + L * nbins + binL[m], (unsigned long long)colw[m]);
    __syncthreads();
    for (int a = threadIdx.x; a < KT * nbins; a += blockDim.x)
        if (sh[a]) atomicAdd(hist + a, sh[a]);
}

// ---------------------------------------------------------------------------------------------
// host side
// ---------------------------------------------------------------------------------------------
// rows of the last axis' bulk that intersect [first, first+count); false if the fast path does not apply
bool pq_grid_tiling(const peuqse_problem* p, const DevSweep& S, long long first, long long count, int blocks_per_sm,
                    GridTiling* T, long long* slab_count) {
    if (S.kind != 0) return false;
    const int L = S.k - 1;
    const long long r2 = (long long)S.r2[L];
    if (r2 < GR_THREADS) return false;                 // rows too short to amortise the per-row work
    const long long ML = (long long)S.mprod[L];
    const long long end = first + count;
    *slab_count = first < ML ? (end < ML ? end : ML) - first : 0;      // points with the last axis at its centre
    T->ML = ML;
    T->r2 = r2;
    const long long g0 = first > ML ? first : ML;
    if (end <= g0) { T->b_lo = T->b_hi = 0; T->nchunks = 1; T->nrg = 1; return true; }
    T->b_lo = (g0 - ML) / r2;
    T->b_hi = (end - ML + r2 - 1) / r2;
    T->nchunks = (int)((r2 + GR_CH - 1) / GR_CH);
    long long rows = T->b_hi - T->b_lo;
    long long nrg = ((long long)p->sm_count * blocks_per_sm + T->nchunks - 1) / T->nchunks;
    if (nrg > rows) nrg = rows;
    if (nrg < 1) nrg = 1;
    if (nrg * T->nchunks > PQ_SWEEP_MAX_BLOCKS) nrg = PQ_SWEEP_MAX_BLOCKS / T->nchunks;
    T->nrg = (int)nrg;
    return true;
}

static bool fast_model_ok(const peuqse_problem* p) {
    const DevProblem& d = p->d;
    if (d.fast_n <= 0) return false;
    if (d.model_id == PEUQSE_MODEL_QUADRATIC) return true;
    if (d.model_id == PEUQSE_MODEL_LINEAR) return d.k == 2 || d.k == 3 || d.k == 4 || d.k == 6;
    return false;
}

template <int MODEL, int KT, int NT, bool EXACT>
static int launch_grid_pass1(const peuqse_problem* p, const DevSweep& S, const GridTiling& T, long long first, long long count,
                             double* logp, peuqse_best* best, void* work, cudaStream_t st) {
    unsigned int* ticket = (unsigned int*)work;
    Best* partial = (Best*)((char*)work + PQ_SWEEP_TICKET_BYTES);
    k_grid_pass1<MODEL, KT, NT, EXACT><<<T.nchunks * T.nrg, GR_THREADS, 0, st>>>(p->d, S, T, first, count, logp, best, partial, ticket);
    PQ_CUDA_OK(cudaGetLastError());
    return PEUQSE_OK;
}

// returns PEUQSE_OK and sets *handled when the specialised kernels ran for the bulk; the caller runs the generic
// kernel on the slab [first, first + *slab_count)
int pq_grid_pass1(const peuqse_problem* p, const DevSweep& S, long long first, long long count, double* logp,
                  peuqse_best* best, void* work, cudaStream_t st, bool* handled, long long* slab_count) {
    *handled = false;
    GridTiling T;
    if (!fast_model_ok(p) || !pq_grid_tiling(p, S, first, count, 4, &T, slab_count)) return PEUQSE_OK;
    *handled = true;
    if (T.b_hi <= T.b_lo) return PEUQSE_OK;
    const int n = p->d.fast_n;
    if (p->d.model_id == PEUQSE_MODEL_QUADRATIC) {
        switch (n) {
#define GQ(N) case N: return launch_grid_pass1<PEUQSE_MODEL_QUADRATIC, 2, N, true>(p, S, T, first, count, logp, best, work, st);
            GQ(1) GQ(2) GQ(3) GQ(4) GQ(5) GQ(6) GQ(7) GQ(8)
#undef GQ
        }
    } else {
#define GL(K)                                                                                                             \
    case K:                                                                                                               \
        return n <= 4 ? launch_grid_pass1<PEUQSE_MODEL_LINEAR, K, 4, false>(p, S, T, first, count, logp, best, work, st)   \
                      : launch_grid_pass1<PEUQSE_MODEL_LINEAR, K, 8, false>(p, S, T, first, count, logp, best, work, st);
        switch (p->d.k) { GL(2) GL(3) GL(4) GL(6) }
#undef GL
    }
    return pq_set_error(PEUQSE_ERR_INVALID, "grid fast path dispatch");
}

int pq_grid_pass2(const peuqse_problem* p, const DevSweep& S, long long first, long long count, const double* logp,
                  const peuqse_best* best, double log_rel, double* acc, long long* kidx, double* klogp, long long kcap,
                  long long* kcount, void* work, cudaStream_t st, bool* handled, long long* slab_count) {
    *handled = false;
    GridTiling T;
    if (!logp || !pq_grid_tiling(p, S, first, count, 4, &T, slab_count)) return PEUQSE_OK;
    if (!(S.k == 1 || S.k == 2 || S.k == 3 || S.k == 4 || S.k == 6)) return PEUQSE_OK;
    *handled = true;
    if (T.b_hi <= T.b_lo) return PEUQSE_OK;
    unsigned int* ticket = (unsigned int*)work;
    double* partial = (double*)((char*)work + PQ_SWEEP_TICKET_BYTES);
    const int blocks = T.nchunks * T.nrg;
#define G2(K) case K: k_grid_pass2<K><<<blocks, GR_THREADS, 0, st>>>(S, T, first, count, logp, best, log_rel, partial, acc, ticket, kidx, klogp, kcap, (unsigned long long*)kcount); break;
    switch (S.k) { G2(1) G2(2) G2(3) G2(4) G2(6) }
#undef G2
    PQ_CUDA_OK(cudaGetLastError());
    return PEUQSE_OK;
}

int pq_grid_hist(const peuqse_problem* p, const DevSweep& S, long long first, long long count, const double* logp,
                 const peuqse_best* best, double log_rel, const double* edges, int nbins, unsigned long long* hist,
                 cudaStream_t st, bool* handled, long long* slab_count) {
    *handled = false;
    GridTiling T;
    if (!logp || !pq_grid_tiling(p, S, first, count, 4, &T, slab_count)) return PEUQSE_OK;
    if (!(S.k == 1 || S.k == 2 || S.k == 3 || S.k == 4 || S.k == 6)) return PEUQSE_OK;
    *handled = true;
    if (T.b_hi <= T.b_lo) return PEUQSE_OK;
    const int blocks = T.nchunks * T.nrg;
    const size_t smem = (size_t)S.k * nbins * sizeof(unsigned long long);
#define G3(K) case K: k_grid_hist<K><<<blocks, GR_THREADS, smem, st>>>(S, T, first, count, logp, best, log_rel, edges, nbins, hist); break;
    switch (S.k) { G3(1) G3(2) G3(3) G3(4) G3(6) }
#undef G3
    PQ_CUDA_OK(cudaGetLastError());
    return PEUQSE_OK;
}
