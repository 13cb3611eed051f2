// Specialised grid-sweep kernels (the C2 hot path): the bulk of combinationGenerator's list
// (CombinationGeneratorModule.py:40-81) is, for the last axis L with radius r, the set
//     g = M_L + b*2r + j,   b = index of a lower-axes combination (a "row"),  j in [0, 2r)
// so consecutive g walk along the last axis.  A CTA owns a fixed chunk of j's and strides over rows:
//   - every thread keeps its last-axis values theta_L(j) and their prior terms (incl. the fp64 division of
//     IP.py:2767) in registers for the whole kernel,
//   - the lower-axis point, its prior terms and every model sub-expression that depends on it are
//     block-uniform per row and are computed once per row (the 8 unrolled, branch-free evaluations share them),
//   - per point only the last-axis part of the model, the residuals and the two reductions remain;
//     stores / loads of logP are fully coalesced along the row,
//   - pass 1 also records the maximum of every warp tile (8 x 32 points); passes 2 and 3 skip tiles whose
//     maximum is below the retention threshold, so they read only the retained region of the logP array.
// The slab g < M_L (last axis at its centre, a lower-dimensional grid) is handled per point at the end.
// Same arithmetic and summation order as the generic kernels (sweep.cu / eval.cuh): logP is bit-identical.
#include "common.cuh"
#include "eval.cuh"
#include "eval_fast.cuh"
#include "launch.cuh"
#include "sweep_common.cuh"

#define GR_THREADS 256
#define GR_WARPS (GR_THREADS / 32)
#define GR_JPT 8
#define GR_CH (GR_THREADS * GR_JPT)

struct GridTiling {
    long long ML;          // first global index of the bulk of the last axis = prod_{d<L}(2 r_d + 1)
    long long r2;          // 2 r_L
    long long b_lo, b_hi;  // rows intersecting [first, first+count)
    long long slab_first, slab_count;   // part of [first, first+count) below ML
    int nchunks, nrg;      // chunks along j, row groups
};

__device__ __forceinline__ double last_axis_value(const DevSweep& S, int L, long long j) {
    const double step = S.interval[L] * (double)((j >> 1) + 1LL);          // (BaseFactor*i), CombinationGeneratorModule.py:59,71
    return (j & 1LL) ? (S.center[L] + step) : (S.center[L] - step);
}

__device__ __forceinline__ bool pq_isnan_bits(double x) {                   // integer pipe, keeps the fp64 pipe free
    const unsigned hi = (unsigned)__double2hiint(x) & 0x7fffffffu;
    return hi > 0x7ff00000u || (hi == 0x7ff00000u && __double2loint(x) != 0);
}

__device__ __forceinline__ size_t tile_index(const GridTiling& T, long long b, int chunk, int warp) {
    return ((size_t)(b - T.b_lo) * T.nchunks + chunk) * GR_WARPS + warp;
}

// block-level merge of per-thread bests into *best_out
__device__ __forceinline__ void block_store_best(Best best, Best* s_best, bool* s_last, Best* partial, unsigned int* ticket,
                                                 peuqse_best* best_out) {
    best = warp_best(best);
    if ((threadIdx.x & 31) == 0) s_best[threadIdx.x >> 5] = best;
    __syncthreads();
    if (threadIdx.x < 32) {
        Best b = (threadIdx.x < GR_WARPS) ? s_best[threadIdx.x] : best_none();
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
            best_out->logp = b.v;
            best_out->index = (b.i == 0x7fffffffffffffffLL) ? -1 : b.i;
            *ticket = 0;
        }
    }
}

// logLikelihood with the NaN test on the integer pipe
template <int MODEL, int KT, int NT, bool EXACT>
__device__ __forceinline__ double grid_loglik(const DevProblem& P, const double (&th)[KT]) {
    double q[NT];
    FastQ<MODEL, KT, NT, EXACT, 0>::fill(P, th, q);
    const double maha = ordered_sum_static<NT>(EXACT ? NT : P.fast_n, q);
    const double ll = -0.5 * (P.resp_c[0] + maha);
    return pq_isnan_bits(ll) ? pq_neg_inf() : ll;                 // IP.py:2880-2882
}

// ---------------------------------------------------------------------------------------------
// pass 1
//   SIMPLE: every parameter has a Gaussian prior term, whitening in natural order, k < 8 -> the prior sum is
//   ((q_0 + q_1) + ...) + q_L with the lower-axis partial sum hoisted per row.
// ---------------------------------------------------------------------------------------------
template <int MODEL, int KT, int NT, bool EXACT, bool SIMPLE, bool STORE>
__global__ void __launch_bounds__(GR_THREADS) k_grid_pass1(const __grid_constant__ DevProblem P, const __grid_constant__ DevSweep S,
                                                           const __grid_constant__ GridTiling T, long long first, long long count,
                                                           double* __restrict__ logp_out, double* __restrict__ tilemax,
                                                           peuqse_best* __restrict__ best_out, Best* __restrict__ partial,
                                                           unsigned int* __restrict__ ticket) {
    __shared__ Best s_best[GR_WARPS];
    __shared__ bool s_last;
    constexpr int L = KT - 1;
    const int chunk = blockIdx.x % T.nchunks, rg = blockIdx.x / T.nchunks;
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const long long j0 = (long long)chunk * GR_CH + threadIdx.x;
    const long long last = first + count;
    double thL[GR_JPT], qL[GR_JPT];
    bool okL[GR_JPT], jv[GR_JPT];
#pragma unroll
    for (int m = 0; m < GR_JPT; ++m) {
        const long long j = j0 + m * GR_THREADS;
        thL[m] = last_axis_value(S, L, j);
        prior_axis(P, L, thL[m], okL[m], qL[m]);
        jv[m] = j < T.r2;
    }
    Best best = best_none();
    for (long long b = T.b_lo + rg; b < T.b_hi; b += T.nrg) {
        double th[KT], qn[KT];
        grid_point<KT>(S, (unsigned long long)b, th, L);          // lower axes of this row (block-uniform)
        bool ok_low = true;
        double plow = 0.0;
#pragma unroll
        for (int d = 0; d < KT; ++d) {
            qn[d] = 0.0;
            if (d < L) {
                bool okd;
                prior_axis(P, d, th[d], okd, qn[d]);
                ok_low = ok_low && okd;
                plow = (d == 0) ? qn[0] : (plow + qn[d]);         // numpy order for < 8 terms: sequential from the first
            }
        }
        const long long gbase = T.ML + b * T.r2 + j0;
        const bool inside = (T.ML + b * T.r2 >= first) && (T.ML + (b + 1) * T.r2 <= last);   // whole row in range
        double rowmax = pq_neg_inf();
        int rowm = 0;
#pragma unroll
        for (int m = 0; m < GR_JPT; ++m) {
            th[L] = thL[m];
            double lp;
            if (SIMPLE) {
                const double maha = (KT == 1) ? qL[m] : (plow + qL[m]);
                lp = -0.5 * (P.prior_c + maha);                      // _multivariate.py:591
            } else {
                qn[L] = qL[m];
                lp = prior_combine<KT>(P, true, qn);
            }
            const double ll = grid_loglik<MODEL, KT, NT, EXACT>(P, th);
            const double v = (ok_low && okL[m]) ? (ll + lp) : pq_neg_inf();   // IP.py:1649-1653
            const long long g = gbase + m * GR_THREADS;
            const bool valid = jv[m] && (inside || (g >= first && g < last));
            if (valid) {
                if (STORE) __stcs(logp_out + (g - first), v);
                if (v > rowmax) { rowmax = v; rowm = m; }          // strict '>', increasing g within the row
            }
        }
        if (rowmax > best.v) { best.v = rowmax; best.i = gbase + rowm * GR_THREADS; }   // rows in increasing g
        if (STORE) {
            double wm = rowmax;
#pragma unroll
            for (int o = 16; o > 0; o >>= 1) wm = fmax(wm, __shfl_xor_sync(0xffffffffu, wm, o));
            if (lane == 0) tilemax[tile_index(T, b, chunk, warp)] = wm;
        }
    }
    // slab: last axis at its centre (indices below M_L), per-point decode
    for (long long g = T.slab_first + (long long)blockIdx.x * GR_THREADS + threadIdx.x; g < T.slab_first + T.slab_count;
         g += (long long)gridDim.x * GR_THREADS) {
        double th[KT], qn[KT];
        grid_point<KT>(S, (unsigned long long)g, th, KT);
        bool ok = true;
#pragma unroll
        for (int d = 0; d < KT; ++d) {
            bool okd;
            prior_axis(P, d, th[d], okd, qn[d]);
            ok = ok && okd;
        }
        const double lp = prior_combine<KT>(P, true, qn);
        const double ll = grid_loglik<MODEL, KT, NT, EXACT>(P, th);
        const double v = ok ? (ll + lp) : pq_neg_inf();
        if (STORE) __stcs(logp_out + (g - first), v);
        if (v > best.v || (v == best.v && g < best.i)) { best.v = v; best.i = g; }
    }
    block_store_best(best, s_best, &s_last, partial, ticket, best_out);
}

// ---------------------------------------------------------------------------------------------
// pass 2 (logP read back from HBM, only tiles that hold retained points)
// ---------------------------------------------------------------------------------------------
__device__ __forceinline__ double blk_sum(double v, double* s_red) {
    v = warp_sum(v);
    __syncthreads();
    if ((threadIdx.x & 31) == 0) s_red[threadIdx.x >> 5] = v;
    __syncthreads();
    double t = 0.0;
    if (threadIdx.x < 32) {
        t = (threadIdx.x < GR_WARPS) ? s_red[threadIdx.x] : 0.0;
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
        t = (threadIdx.x < GR_WARPS) ? s_red[threadIdx.x] : pq_pos_inf();
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) t = fmin(t, __shfl_xor_sync(0xffffffffu, t, o));
    }
    return t;
}

template <int KT>
struct GridAcc {
    double w, wexp, nk;
    double s1[KT], s2[KT * (KT + 1) / 2], mn[KT], mx[KT];
    __device__ __forceinline__ void init() {
        w = wexp = nk = 0.0;
#pragma unroll
        for (int d = 0; d < KT; ++d) { s1[d] = 0.0; mn[d] = pq_pos_inf(); mx[d] = pq_neg_inf(); }
#pragma unroll
        for (int e = 0; e < KT * (KT + 1) / 2; ++e) s2[e] = 0.0;
    }
    __device__ __forceinline__ void add(double v, double thr, const double (&th)[KT], const double (&shift)[KT]) {
        const double wt = trunc(exp(v - thr));                      // int(np.exp(...)), IP.py:3485-3486
        w += wt;
        wexp = fma(wt, exp(v), wexp);
        nk += 1.0;
        int e = 0;
#pragma unroll
        for (int d = 0; d < KT; ++d) {
            const double wd = wt * (th[d] - shift[d]);
            s1[d] += wd;
#pragma unroll
            for (int d2 = d; d2 < KT; ++d2) { s2[e] = fma(wd, th[d2] - shift[d2], s2[e]); ++e; }
            mn[d] = fmin(mn[d], th[d]);
            mx[d] = fmax(mx[d], th[d]);
        }
    }
};

template <int KT>
__global__ void __launch_bounds__(GR_THREADS) k_grid_pass2(const __grid_constant__ DevSweep S, const __grid_constant__ GridTiling T,
                                                           long long first, long long count, const double* __restrict__ logp_in,
                                                           const double* __restrict__ tilemax,
                                                           const peuqse_best* __restrict__ best_in, double log_rel,
                                                           double* __restrict__ partial, double* __restrict__ acc_out,
                                                           unsigned int* __restrict__ ticket, long long* __restrict__ keep_idx,
                                                           double* __restrict__ keep_logp, long long keep_cap,
                                                           unsigned long long* __restrict__ keep_count) {
    __shared__ double s_red[GR_WARPS];
    __shared__ bool s_last;
    constexpr int L = KT - 1;
    const long long best_index = best_in->index;
    const bool none = best_index < 0;
    const double thr = none ? pq_pos_inf() : best_in->logp + log_rel;   // IP.py:3477
    double shift[KT];
    grid_point<KT>(S, (unsigned long long)(none ? 0 : best_index), shift, KT);   // moments about the MAP point
    const int chunk = blockIdx.x % T.nchunks, rg = blockIdx.x / T.nchunks;
    const int warp = threadIdx.x >> 5;
    const long long j0 = (long long)chunk * GR_CH + threadIdx.x;
    const long long last = first + count;
    double thL[GR_JPT];
#pragma unroll
    for (int m = 0; m < GR_JPT; ++m) thL[m] = last_axis_value(S, L, j0 + m * GR_THREADS);
    GridAcc<KT> A;
    A.init();
    for (long long b = T.b_lo + rg; b < T.b_hi && !none; b += T.nrg) {
        if (!(__ldg(tilemax + tile_index(T, b, chunk, warp)) > thr)) continue;     // nothing retained in this warp tile
        const long long gbase = T.ML + b * T.r2 + j0;
        double v[GR_JPT];
#pragma unroll
        for (int m = 0; m < GR_JPT; ++m) {                          // all loads of the row first (memory-level parallelism)
            const long long g = gbase + m * GR_THREADS;
            v[m] = pq_neg_inf();
            if (j0 + m * GR_THREADS < T.r2 && g >= first && g < last) v[m] = __ldcs(logp_in + (g - first));
        }
        double th[KT];
        grid_point<KT>(S, (unsigned long long)b, th, L);
#pragma unroll
        for (int m = 0; m < GR_JPT; ++m) {
            if (v[m] > thr) {                                       // IP.py:3481
                th[L] = thL[m];
                A.add(v[m], thr, th, shift);
                if (keep_idx) {
                    const unsigned long long slot = atomicAdd(keep_count, 1ULL);
                    if ((long long)slot < keep_cap) { keep_idx[slot] = gbase + m * GR_THREADS; keep_logp[slot] = v[m]; }
                }
            }
        }
    }
    for (long long g = T.slab_first + (long long)blockIdx.x * GR_THREADS + threadIdx.x; g < T.slab_first + T.slab_count && !none;
         g += (long long)gridDim.x * GR_THREADS) {
        const double v = __ldcs(logp_in + (g - first));
        if (v > thr) {
            double th[KT];
            grid_point<KT>(S, (unsigned long long)g, th, KT);
            A.add(v, thr, th, shift);
            if (keep_idx) {
                const unsigned long long slot = atomicAdd(keep_count, 1ULL);
                if ((long long)slot < keep_cap) { keep_idx[slot] = g; keep_logp[slot] = v; }
            }
        }
    }
    double* mine = partial + (size_t)blockIdx.x * PQ_ACC_MAX;
    double t;
    t = blk_sum(A.w, s_red);    if (threadIdx.x == 0) mine[PEUQSE_ACC_W] = t;
    t = blk_sum(A.wexp, s_red); if (threadIdx.x == 0) mine[PEUQSE_ACC_WEXP] = t;
    t = blk_sum(A.nk, s_red);   if (threadIdx.x == 0) mine[PEUQSE_ACC_NKEPT] = t;
    {
        int e = 0;
#pragma unroll
        for (int d = 0; d < KT; ++d) {
            t = blk_sum(A.s1[d], s_red);   if (threadIdx.x == 0) mine[PEUQSE_ACC_S1 + d] = t;
            t = blk_min(A.mn[d], s_red);   if (threadIdx.x == 0) mine[PEUQSE_ACC_MIN(KT) + d] = t;
            t = -blk_min(-A.mx[d], s_red); if (threadIdx.x == 0) mine[PEUQSE_ACC_MAX(KT) + d] = t;
#pragma unroll
            for (int d2 = d; d2 < KT; ++d2) {
                t = blk_sum(A.s2[e], s_red);
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
                                                          const double* __restrict__ tilemax,
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
    const int warp = threadIdx.x >> 5;
    const long long j0 = (long long)chunk * GR_CH + threadIdx.x;
    const long long last = first + count;
    double colw[GR_JPT];
#pragma unroll
    for (int m = 0; m < GR_JPT; ++m) colw[m] = 0.0;
    for (long long b = T.b_lo + rg; b < T.b_hi; b += T.nrg) {
        if (!(__ldg(tilemax + tile_index(T, b, chunk, warp)) > thr)) continue;
        const long long gbase = T.ML + b * T.r2 + j0;
        double v[GR_JPT];
#pragma unroll
        for (int m = 0; m < GR_JPT; ++m) {
            const long long g = gbase + m * GR_THREADS;
            v[m] = pq_neg_inf();
            if (j0 + m * GR_THREADS < T.r2 && g >= first && g < last) v[m] = __ldcs(logp_in + (g - first));
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
    for (int m = 0; m < GR_JPT; ++m) {
        if (colw[m] > 0.0) {
            const int bin = hist_bin(edges + L * (nbins + 1), nbins, last_axis_value(S, L, j0 + m * GR_THREADS));
            if (bin >= 0) atomicAdd(sh + L * nbins + bin, (unsigned long long)colw[m]);
        }
    }
    for (long long g = T.slab_first + (long long)blockIdx.x * GR_THREADS + threadIdx.x; g < T.slab_first + T.slab_count;
         g += (long long)gridDim.x * GR_THREADS) {
        const double v = __ldcs(logp_in + (g - first));
        if (v > thr) {
            const unsigned long long w = (unsigned long long)trunc(exp(v - thr));
            double th[KT];
            grid_point<KT>(S, (unsigned long long)g, th, KT);
#pragma unroll
            for (int d = 0; d < KT; ++d) {
                const int bin = hist_bin(edges + d * (nbins + 1), nbins, th[d]);
                if (bin >= 0) atomicAdd(sh + d * nbins + bin, w);
            }
        }
    }
    __syncthreads();
    for (int a = threadIdx.x; a < KT * nbins; a += blockDim.x)
        if (sh[a]) atomicAdd(hist + a, sh[a]);
}

// ---------------------------------------------------------------------------------------------
// host side
// ---------------------------------------------------------------------------------------------
static bool grid_tiling(const peuqse_problem* p, const DevSweep& S, long long first, long long count, int resident_blocks,
                        GridTiling* T) {
    if (S.kind != 0) return false;
    const int L = S.k - 1;
    const long long r2 = (long long)S.r2[L];
    if (r2 < GR_THREADS) return false;                 // rows too short to amortise the per-row work
    const long long ML = (long long)S.mprod[L];
    const long long end = first + count;
    T->ML = ML;
    T->r2 = r2;
    T->slab_first = first;
    T->slab_count = first < ML ? (end < ML ? end : ML) - first : 0;      // points with the last axis at its centre
    T->nchunks = (int)((r2 + GR_CH - 1) / GR_CH);
    const long long g0 = first > ML ? first : ML;
    if (end <= g0) {
        T->b_lo = T->b_hi = 0;
        T->nrg = 1;
        return true;
    }
    T->b_lo = (g0 - ML) / r2;
    T->b_hi = (end - ML + r2 - 1) / r2;
    const long long rows = T->b_hi - T->b_lo;
    long long nrg = resident_blocks / T->nchunks;       // one resident wave: no tail
    if (nrg > rows) nrg = rows;
    if (nrg < 1) nrg = 1;
    if (nrg * T->nchunks > PQ_SWEEP_MAX_BLOCKS) nrg = PQ_SWEEP_MAX_BLOCKS / T->nchunks;
    T->nrg = (int)nrg;
    return true;
}

template <class K>
static int resident_blocks(const peuqse_problem* p, K kern, size_t smem) {
    int per_sm = 1;
    if (cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, kern, GR_THREADS, smem) != cudaSuccess || per_sm < 1) per_sm = 1;
    return per_sm * p->sm_count;
}

static bool fast_model_ok(const peuqse_problem* p) {
    const DevProblem& d = p->d;
    if (d.fast_n <= 0) return false;
    if (d.model_id == PEUQSE_MODEL_QUADRATIC) return true;
    if (d.model_id == PEUQSE_MODEL_LINEAR) return d.k == 2 || d.k == 3 || d.k == 4 || d.k == 6;
    return false;
}

static bool grid_k_ok(int k) { return k == 1 || k == 2 || k == 3 || k == 4 || k == 6; }

// bytes of the tile-maximum array for this range (0 when the specialised kernels do not apply)
long long pq_grid_aux_bytes(const peuqse_problem* p, const DevSweep& S, long long first, long long count) {
    GridTiling T;
    if (!p || !fast_model_ok(p) || !grid_k_ok(S.k) || !grid_tiling(p, S, first, count, 1, &T)) return 0;
    return (long long)((T.b_hi - T.b_lo) * T.nchunks * GR_WARPS + 1) * (long long)sizeof(double);
}

template <int MODEL, int KT, int NT, bool EXACT>
static int launch_grid_pass1(const peuqse_problem* p, const DevSweep& S, long long first, long long count, double* logp,
                             double* tilemax, peuqse_best* best, void* work, cudaStream_t st) {
    const DevProblem& d = p->d;
    const bool simple = d.prior_mode == PEUQSE_PRIOR_DIAG && d.prior_identity && d.kp == d.k && KT < 8;
    unsigned int* ticket = (unsigned int*)work;
    Best* partial = (Best*)((char*)work + PQ_SWEEP_TICKET_BYTES);
    GridTiling T;
#define GP1_LAUNCH(SIMPLE, STORE)                                                                                            \
    {                                                                                                                         \
        auto kern = k_grid_pass1<MODEL, KT, NT, EXACT, SIMPLE, STORE>;                                                        \
        grid_tiling(p, S, first, count, resident_blocks(p, kern, 0), &T);                                                     \
        kern<<<T.nchunks * T.nrg, GR_THREADS, 0, st>>>(p->d, S, T, first, count, logp, tilemax, best, partial, ticket);        \
    }
    const bool store = logp != nullptr && tilemax != nullptr;
    if (simple && store) GP1_LAUNCH(true, true)
    else if (simple) GP1_LAUNCH(true, false)
    else if (store) GP1_LAUNCH(false, true)
    else GP1_LAUNCH(false, false)
#undef GP1_LAUNCH
    PQ_CUDA_OK(cudaGetLastError());
    return PEUQSE_OK;
}

// *handled = true when the specialised kernel covered the whole range [first, first+count)
int pq_grid_pass1(const peuqse_problem* p, const DevSweep& S, long long first, long long count, double* logp, double* tilemax,
                  peuqse_best* best, void* work, cudaStream_t st, bool* handled) {
    *handled = false;
    GridTiling T;
    if (!fast_model_ok(p) || !grid_tiling(p, S, first, count, 1, &T)) return PEUQSE_OK;
    if (logp && !tilemax) return PEUQSE_OK;            // storing logP without the tile maxima: generic path
    *handled = true;
    const int n = p->d.fast_n;
    if (p->d.model_id == PEUQSE_MODEL_QUADRATIC) {
        switch (n) {
#define GQ(N) case N: return launch_grid_pass1<PEUQSE_MODEL_QUADRATIC, 2, N, true>(p, S, first, count, logp, tilemax, best, work, st);
            GQ(1) GQ(2) GQ(3) GQ(4) GQ(5) GQ(6) GQ(7) GQ(8)
#undef GQ
        }
    } else {
#define GL(K)                                                                                                                    \
    case K:                                                                                                                      \
        return n <= 4 ? launch_grid_pass1<PEUQSE_MODEL_LINEAR, K, 4, false>(p, S, first, count, logp, tilemax, best, work, st)    \
                      : launch_grid_pass1<PEUQSE_MODEL_LINEAR, K, 8, false>(p, S, first, count, logp, tilemax, best, work, st);
        switch (p->d.k) { GL(2) GL(3) GL(4) GL(6) }
#undef GL
    }
    return pq_set_error(PEUQSE_ERR_INVALID, "grid fast path dispatch");
}

int pq_grid_pass2(const peuqse_problem* p, const DevSweep& S, long long first, long long count, const double* logp,
                  const double* tilemax, const peuqse_best* best, double log_rel, double* acc, long long* kidx, double* klogp,
                  long long kcap, long long* kcount, void* work, cudaStream_t st, bool* handled) {
    *handled = false;
    GridTiling T;
    if (!logp || !tilemax || !fast_model_ok(p) || !grid_k_ok(S.k) || !grid_tiling(p, S, first, count, 1, &T)) return PEUQSE_OK;
    *handled = true;
    unsigned int* ticket = (unsigned int*)work;
    double* partial = (double*)((char*)work + PQ_SWEEP_TICKET_BYTES);
#define G2(K)                                                                                                                    \
    case K: {                                                                                                                    \
        grid_tiling(p, S, first, count, resident_blocks(p, k_grid_pass2<K>, 0), &T);                                             \
        k_grid_pass2<K><<<T.nchunks * T.nrg, GR_THREADS, 0, st>>>(S, T, first, count, logp, tilemax, best, log_rel, partial, acc, \
                                                                  ticket, kidx, klogp, kcap, (unsigned long long*)kcount);        \
    } break;
    switch (S.k) { G2(1) G2(2) G2(3) G2(4) G2(6) }
#undef G2
    PQ_CUDA_OK(cudaGetLastError());
    return PEUQSE_OK;
}

int pq_grid_hist(const peuqse_problem* p, const DevSweep& S, long long first, long long count, const double* logp,
                 const double* tilemax, const peuqse_best* best, double log_rel, const double* edges, int nbins,
                 unsigned long long* hist, cudaStream_t st, bool* handled) {
    *handled = false;
    GridTiling T;
    if (!logp || !tilemax || !fast_model_ok(p) || !grid_k_ok(S.k) || !grid_tiling(p, S, first, count, 1, &T)) return PEUQSE_OK;
    *handled = true;
    const size_t smem = (size_t)S.k * nbins * sizeof(unsigned long long);
#define G3(K)                                                                                                                    \
    case K: {                                                                                                                    \
        grid_tiling(p, S, first, count, resident_blocks(p, k_grid_hist<K>, smem), &T);                                           \
        k_grid_hist<K><<<T.nchunks * T.nrg, GR_THREADS, smem, st>>>(S, T, first, count, logp, tilemax, best, log_rel, edges,      \
                                                                    nbins, hist);                                                 \
    } break;
    switch (S.k) { G3(1) G3(2) G3(3) G3(4) G3(6) }
#undef G3
    PQ_CUDA_OK(cudaGetLastError());
    return PEUQSE_OK;
}
