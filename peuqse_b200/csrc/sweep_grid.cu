// Specialised grid-sweep kernels (the C2 hot path): the bulk of combinationGenerator's list
// (CombinationGeneratorModule.py:40-81) is, for the last axis L with radius r, the set
//     g = M_L + b*2r + j,   b = index of a lower-axes combination (a "row"),  j in [0, 2r)
// so consecutive g walk along the last axis.
// pass 1: a CTA owns a fixed chunk of j's and strides over rows.
//   - every thread keeps its last-axis values theta_L(j) and their prior terms (incl. the fp64 division of
//     IP.py:2767) in registers for the whole kernel;
//   - everything that depends only on a row (lower-axis decode, its prior terms, the model's row part, e.g.
//     (x_j - a)^2) is computed ONCE per row per CTA into a shared-memory record and broadcast-read;
//   - per point remain the last-axis part of the model, residuals and the two reductions (~20 fp64 ops);
//     logP stores are coalesced; the maximum of each warp tile (8 x 32 points) is kept as a rounded-up float.
// pass 2: tiles are visited in flat order with a stride coprime to the number of chunks (even load for a compact
//   retained region); tiles whose maximum is below the threshold are skipped without touching logP; besides the
//   moments it accumulates per-row and per-column weight sums, from which a tiny kernel bins the marginal
//   histograms (pass 3) without reading logP again.
// The slab g < M_L (last axis at its centre, a lower-dimensional grid) is handled per point.
// Same arithmetic and summation order as the generic kernels (sweep.cu / eval.cuh): logP is bit-identical.
#include "common.cuh"
#include "eval.cuh"
#include "eval_fast.cuh"
#include "launch.cuh"
#include "sweep_common.cuh"
#include "tma.cuh"

#define GR_THREADS 256
#define GR_WARPS (GR_THREADS / 32)
#define GR_JPT 8
#define GR_CH (GR_THREADS * GR_JPT)
#define GR_RB 8         // rows per work unit of pass 1
#define PQ_GRID_CTR_WORD 24   // unit counters of pass 1: ticket words 24..63 (bytes 96..255 of the ticket area, after ShortHdr)
#define PQ_GRID_MAX_CTR 32
#define GR_MSTRIDE 32   // a thread's 8 points are 32 apart: a warp covers 256 CONTIGUOUS positions of a row (compact tiles)
// position of the thread's first point inside its chunk
#define GR_JL0 ((int)(threadIdx.x >> 5) * (32 * GR_JPT) + (int)(threadIdx.x & 31))

struct GridTiling {
    long long ML;          // first global index of the bulk of the last axis = prod_{d<L}(2 r_d + 1)
    long long r2;          // 2 r_L
    long long b_lo, b_hi;  // rows intersecting [first, first+count)
    long long slab_first, slab_count;   // part of [first, first+count) below ML
    int nchunks, nrg;      // chunks along j, row groups
    int dyn;               // pass 1 hands out its work units through atomic counters (nchunks <= PQ_GRID_MAX_CTR)
};

// auxiliary arrays inside d_work (after the base area)
struct GridAux {
    unsigned int* tilemax;   // [ntiles * GR_WARPS] order-preserving keys of float upper bounds
    double* colsum;          // [r2]   sum of weights per last-axis position
    double* rowsum;          // [rows] sum of weights per row
    double* rowsumL;         // [rows] sum of weights * (theta_L - shift_L) per row
    int* valid;              // rowsum/colsum hold the sums of the last pass 2
};

__device__ __forceinline__ double last_axis_value(const DevSweep& S, int L, long long j) {
    const double step = S.interval[L] * (double)((j >> 1) + 1LL);          // (BaseFactor*i), CombinationGeneratorModule.py:59,71
    return (j & 1LL) ? (S.center[L] + step) : (S.center[L] - step);
}

__device__ __forceinline__ bool pq_isnan_bits(double x) {                   // integer pipe, keeps the fp64 pipe free
    const unsigned hi = (unsigned)__double2hiint(x) & 0x7fffffffu;
    return hi > 0x7ff00000u || (hi == 0x7ff00000u && __double2loint(x) != 0);
}

// Tile maxima are kept as order-preserving keys of the HIGH WORD of a double (larger value <=> larger key; NaNs are
// excluded by the callers); key_value() returns an upper bound of every double with that key.  Key 0 decodes to NaN
// (compares false with everything: "no tile").
__device__ __forceinline__ unsigned int hi_okey(unsigned int h) { return (h & 0x80000000u) ? ~h : (h | 0x80000000u); }
__device__ __forceinline__ double key_value(unsigned int k) {
    if (k & 0x80000000u) {                                       // non-negative: largest low word
        const unsigned int h = k & 0x7fffffffu;
        return (h >= 0x7ff00000u) ? pq_pos_inf() : __hiloint2double((int)h, (int)0xffffffffu);
    }
    return __hiloint2double((int)~k, 0);                          // negative: smallest magnitude
}

// block-level merge of per-thread bests into *best_out
__device__ __forceinline__ void block_store_best(Best best, Best* s_best, bool* s_last, Best* partial, unsigned int* ticket,
                                                 peuqse_best* best_out, int nctr) {
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
            ticket[PQ_SHORT_VALID_WORD] = 0u;                  // this pass 1 leaves no shortlist (sweep_common.cuh)
        }
        if ((int)threadIdx.x < nctr) ticket[PQ_GRID_CTR_WORD + threadIdx.x] = 0u;    // unit counters ready for the next launch
    }
}

// logLikelihood of a full point with the NaN test on the integer pipe (slab points)
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
//   ((q_0 + q_1) + ...) + q_L with the lower-axis partial sum taken from the row record.
//   row record (doubles): [0] ok flag, [1] plow, [2 .. 2+KT-2] qn of the lower axes, then pre[NT]
// ---------------------------------------------------------------------------------------------
template <int KT, int NT>
struct RowRec {
    static constexpr int QN = 2;
    static constexpr int PRE = 2 + (KT - 1);
    static constexpr int N = ((2 + (KT - 1) + NT) + 1) & ~1;       // even: 16-byte aligned records
};

// Work distribution: a "unit" is GR_RB consecutive rows of one chunk of columns.  T.dyn: units are handed out by one
// atomic counter per chunk (ticket words PQ_GRID_CTR_WORD ...): a CTA starts on chunk blockIdx % nchunks, keeps its
// per-column registers while that chunk has units and then moves on to the next chunk, so SMs that were given fewer
// CTAs (or faster ones) absorb the difference.  Otherwise (more chunks than counters) unit u of a chunk goes to row
// group u % nrg.  Row records of the NEXT unit are staged by warp 0 after its share of the current unit (the atomic is
// issued before, its result consumed after), double-buffered: one __syncthreads per unit.
// Maximum tracking: the eight values of a thread are compared on the integer pipe through their high words (for negative
// finite doubles a larger high word is a smaller value); the exact fp64 scan runs only when a row could hold a new
// thread-best (high word <= that of the thread's best), a NaN / -inf, or a non-negative value.
template <int MODEL, int KT, int NT, bool EXACT, bool SIMPLE, bool STORE, bool BOUNDS>
__global__ void __launch_bounds__(GR_THREADS, 3) k_grid_pass1(const __grid_constant__ DevProblem P, const __grid_constant__ DevSweep S,
                                                           const __grid_constant__ GridTiling T, long long first, long long count,
                                                           double* __restrict__ logp_out, unsigned int* __restrict__ tilemax,
                                                           peuqse_best* __restrict__ best_out, Best* __restrict__ partial,
                                                           unsigned int* __restrict__ ticket) {
    using R = RowRec<KT, NT>;
    __shared__ __align__(16) double s_rec[2][GR_RB * R::N];
    __shared__ long long s_unit[2];
    __shared__ Best s_best[GR_WARPS];
    __shared__ bool s_last;
    constexpr int L = KT - 1;
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const long long last = first + count;
    const long long nrows = T.b_hi - T.b_lo;
    const long long nunits = (nrows + GR_RB - 1) / GR_RB;
    unsigned int* ctr = ticket + PQ_GRID_CTR_WORD;
    Best best = best_none();
    unsigned int lim = 0xffefffffu;            // exact scan when a row's smallest high word is <= lim (see above)
    const int tstride = T.nchunks * GR_WARPS;
    int chunk = blockIdx.x % T.nchunks;
    long long static_next = blockIdx.x / T.nchunks;               // !T.dyn: units rg, rg + nrg, ...
    const int attempts = (nrows > 0) ? (T.dyn ? T.nchunks : 1) : 0;
    for (int attempt = 0; attempt < attempts; ++attempt, chunk = (chunk + 1 == T.nchunks) ? 0 : chunk + 1) {
        __syncthreads();                                          // s_unit / s_rec of the previous chunk are no longer read
        if (threadIdx.x == 0) {
            long long u0;
            if (T.dyn) u0 = (long long)atomicAdd(ctr + chunk, 1u);
            else { u0 = static_next; static_next += T.nrg; }
            s_unit[0] = u0;
        }
        __syncthreads();
        if (s_unit[0] >= nunits) continue;                        // this chunk is finished (block-uniform)
        const long long j0 = (long long)chunk * GR_CH + GR_JL0;
        const bool warp_full = (long long)chunk * GR_CH + (long long)(warp + 1) * (32 * GR_JPT) <= T.r2;   // all 256 columns exist
        double thL[GR_JPT], qL[GR_JPT];
        bool okL[GR_JPT], jv[GR_JPT];
#pragma unroll
        for (int m = 0; m < GR_JPT; ++m) {
            const long long j = j0 + m * GR_MSTRIDE;
            thL[m] = last_axis_value(S, L, j);
            prior_axis(P, L, thL[m], okL[m], qL[m]);
            jv[m] = j < T.r2;
        }
        // row records of unit `u` into buffer `buf` (lanes < GR_RB of warp 0)
        auto stage = [&](long long u, int buf) {
            if (lane == 0) s_unit[buf] = u;
            const long long b = T.b_lo + u * GR_RB + lane;
            if (lane < GR_RB && u < nunits && b < T.b_hi) {
                double th[KT];
                grid_point<KT>(S, (unsigned long long)b, th, L);      // lower axes of this row
                double* rec = s_rec[buf] + lane * R::N;
                bool ok_low = true;
                double plow = 0.0;
#pragma unroll
                for (int d = 0; d < L; ++d) {
                    bool okd;
                    double q;
                    prior_axis(P, d, th[d], okd, q);
                    ok_low = ok_low && okd;
                    plow = (d == 0) ? q : (plow + q);                 // numpy order for < 8 terms: sequential from the first
                    rec[R::QN + d] = q;
                }
                rec[0] = ok_low ? 1.0 : 0.0;
                rec[1] = plow;
                FastRowPre<MODEL, KT, NT, 0>::fill(P, th, rec + R::PRE);
            }
        };
        if (warp == 0) stage(s_unit[0], 0);
        __syncthreads();
        int cur = 0;
        while (true) {
            const long long u = s_unit[cur];
            if (u >= nunits) break;
            long long un = 0;
            if (threadIdx.x == 0) {                                // next unit: issued now, consumed after this unit's rows
                if (T.dyn) un = (long long)atomicAdd(ctr + chunk, 1u);
                else { un = static_next; static_next += T.nrg; }
            }
            const long long bu = T.b_lo + u * GR_RB;
            const int nr = (int)((T.b_hi - bu < GR_RB) ? (T.b_hi - bu) : GR_RB);
            long long row0 = T.ML + bu * T.r2;
            const bool unit_inside = (row0 >= first) && (row0 + (long long)nr * T.r2 <= last);   // block-uniform
            unsigned int* tptr = tilemax + ((size_t)(bu - T.b_lo) * T.nchunks + chunk) * GR_WARPS + warp;
            for (int tt = 0; tt < nr; ++tt, row0 += T.r2, tptr += tstride) {
                const double* rec = s_rec[cur] + tt * R::N;           // broadcast reads
                const bool ok_low = rec[0] != 0.0;
                const double plow = rec[1];
                double pre[NT], qn[KT];
#pragma unroll
                for (int j = 0; j < NT; ++j) pre[j] = rec[R::PRE + j];
#pragma unroll
                for (int d = 0; d < KT; ++d) qn[d] = (d < L) ? rec[R::QN + (d < L ? d : 0)] : 0.0;
                const long long gbase = row0 + j0;
                double v[GR_JPT];
#pragma unroll
                for (int m = 0; m < GR_JPT; ++m) {                    // branch-free evaluation of the thread's 8 points
                    double lp;
                    if (SIMPLE) {
                        const double maha = (KT == 1) ? qL[m] : (plow + qL[m]);
                        lp = -0.5 * (P.prior_c + maha);                // _multivariate.py:591
                    } else {
                        qn[L] = qL[m];
                        lp = prior_combine<KT>(P, true, qn);
                    }
                    double q[NT];
                    FastQLast<MODEL, KT, NT, EXACT, 0>::fill(P, pre, thL[m], q);
                    const double maha_l = ordered_sum_static<NT>(EXACT ? NT : P.fast_n, q);
                    const double ll = -0.5 * (P.resp_c[0] + maha_l);
                    v[m] = ll + lp;                                    // IP.py:1653
                    if (BOUNDS) v[m] = (ok_low && okL[m]) ? v[m] : pq_neg_inf();   // IP.py:1649-1650
                }
                unsigned int okey;
                if (unit_inside && warp_full) {                       // every point of the warp's 256-column segment exists
                    unsigned int mn = (unsigned int)__double2hiint(v[0]), mx = mn;
#pragma unroll
                    for (int m = 1; m < GR_JPT; ++m) {
                        const unsigned int h = (unsigned int)__double2hiint(v[m]);
                        mn = min(mn, h);
                        mx = max(mx, h);
                    }
                    bool exact = false;
                    if (mx >= 0xfff00000u || mn < 0x80000000u) {      // a NaN / -inf or a non-negative value: rare
#pragma unroll
                        for (int m = 0; m < GR_JPT; ++m)
                            if (pq_isnan_bits(v[m])) v[m] = pq_neg_inf();     // NaN simulated value -> -inf (IP.py:2880-2882)
                        exact = true;
                    }
                    if (STORE) {
#pragma unroll
                        for (int m = 0; m < GR_JPT; ++m) __stcs(logp_out + (gbase + m * GR_MSTRIDE - first), v[m]);
                    }
                    if (exact || mn <= lim) {                         // the row may hold this thread's new best
                        double rowmax = v[0];
                        int rowm = 0;
#pragma unroll
                        for (int m = 1; m < GR_JPT; ++m)
                            if (v[m] > rowmax) { rowmax = v[m]; rowm = m; }    // strict '>', increasing g within the row
                        const long long g = gbase + rowm * GR_MSTRIDE;
                        if (rowmax > best.v || (rowmax == best.v && g < best.i && rowmax > pq_neg_inf())) {
                            best.v = rowmax;
                            best.i = g;
                            const unsigned int hb = (unsigned int)__double2hiint(rowmax);
                            lim = (hb & 0x80000000u) ? min(hb, 0xffefffffu) : 0x7fffffffu;
                        }
                        okey = hi_okey((unsigned int)__double2hiint(rowmax));
                    } else {
                        okey = ~mn;                                   // all negative finite: key of the smallest high word
                    }
                } else {                                              // ragged chunk or first / last rows of a rank's range
                    unsigned int emax = 0u;
#pragma unroll
                    for (int m = 0; m < GR_JPT; ++m) emax = max(emax, (unsigned int)__double2hiint(v[m]) & 0x7fffffffu);
                    if (emax >= 0x7ff00000u) {
#pragma unroll
                        for (int m = 0; m < GR_JPT; ++m)
                            if (pq_isnan_bits(v[m])) v[m] = pq_neg_inf();
                    }
                    double rowmax = pq_neg_inf();
                    int rowm = 0;
#pragma unroll
                    for (int m = 0; m < GR_JPT; ++m) {
                        const long long g = gbase + m * GR_MSTRIDE;
                        if (jv[m] && g >= first && g < last) {
                            if (STORE) __stcs(logp_out + (g - first), v[m]);
                            if (v[m] > rowmax) { rowmax = v[m]; rowm = m; }
                        }
                    }
                    const long long g = gbase + rowm * GR_MSTRIDE;
                    if (rowmax > best.v || (rowmax == best.v && g < best.i && rowmax > pq_neg_inf())) {
                        best.v = rowmax;
                        best.i = g;
                        const unsigned int hb = (unsigned int)__double2hiint(rowmax);
                        lim = (hb & 0x80000000u) ? min(hb, 0xffefffffu) : 0x7fffffffu;
                    }
                    okey = hi_okey((unsigned int)__double2hiint(rowmax));
                }
                if (STORE) {
                    const unsigned int wk = __reduce_max_sync(0xffffffffu, okey);
                    if (lane == 0) *tptr = wk;
                }
            }
            if (warp == 0) stage(__shfl_sync(0xffffffffu, un, 0), cur ^ 1);
            __syncthreads();
            cur ^= 1;
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
        if (v > best.v || (v == best.v && g < best.i && v > pq_neg_inf())) { best.v = v; best.i = g; }
    }
    block_store_best(best, s_best, &s_last, partial, ticket, best_out, T.dyn ? T.nchunks : 0);
}

// ---------------------------------------------------------------------------------------------
// pass 2 (logP read back from HBM, only tiles that hold retained points) + histogram weight sums
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
    double w, wr, nk;           // sum w, sum w*exp(logP - thr), retained points
    double s1[KT], s2[KT * (KT + 1) / 2], mn[KT], mx[KT];
    __device__ __forceinline__ void init() {
        w = wr = nk = 0.0;
#pragma unroll
        for (int d = 0; d < KT; ++d) { s1[d] = 0.0; mn[d] = pq_pos_inf(); mx[d] = pq_neg_inf(); }
#pragma unroll
        for (int e = 0; e < KT * (KT + 1) / 2; ++e) s2[e] = 0.0;
    }
    // generic (slab) path
    __device__ __forceinline__ double add(double v, double thr, const double (&th)[KT], const double (&shift)[KT]) {
        const double reps = exp(v - thr);                           // np.exp(...), IP.py:3485
        const double wt = trunc(reps);                              // int(...), IP.py:3486
        w += wt;
        wr = fma(wt, reps, wr);                                     // sum w*exp(logP) = exp(thr) * sum w*reps
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
        return wt;
    }
};

// index of (d, d2>=d) in the packed upper triangle
__host__ __device__ constexpr int tri_index(int KT, int d, int d2) { return d * KT - d * (d - 1) / 2 + (d2 - d); }

template <int KT, bool FUSE>
__global__ void __launch_bounds__(GR_THREADS, FUSE ? 2 : 3) k_grid_pass2(const __grid_constant__ DevSweep S,
                                                                         const __grid_constant__ GridTiling T, long long first,
                                                                         long long count, const double* __restrict__ logp_in,
                                                                         const GridAux aux,
                                                                         const peuqse_best* __restrict__ best_in, double log_rel,
                                                                         double* __restrict__ partial, double* __restrict__ acc_out,
                                                                         unsigned int* __restrict__ ticket,
                                                                         long long* __restrict__ keep_idx,
                                                                         double* __restrict__ keep_logp, long long keep_cap,
                                                                         unsigned long long* __restrict__ keep_count) {
    extern __shared__ double s_col[];                               // FUSE: column sums of this CTA, nchunks*GR_CH doubles
    __shared__ double s_red[GR_WARPS];
    __shared__ bool s_last;
    constexpr int L = KT - 1;
    const long long best_index = best_in->index;
    const bool none = best_index < 0;
    const double thr = none ? pq_pos_inf() : best_in->logp + log_rel;   // IP.py:3477
    double shift[KT];
    grid_point<KT>(S, (unsigned long long)(none ? 0 : best_index), shift, KT);   // moments about the MAP point
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const long long last = first + count;
    const long long ntiles = (T.b_hi - T.b_lo) * T.nchunks;
    if (FUSE) {
        for (int a = threadIdx.x; a < T.nchunks * GR_CH; a += blockDim.x) s_col[a] = 0.0;
        __syncthreads();
    }
    GridAcc<KT> A;
    A.init();
    long long t = blockIdx.x;
    unsigned int tk = (t < ntiles && !none) ? __ldg(aux.tilemax + (size_t)t * GR_WARPS + warp) : 0u;
    for (; t < ntiles && !none; t += gridDim.x) {
        const long long tn = t + gridDim.x;
        const unsigned int tk_next = (tn < ntiles) ? __ldg(aux.tilemax + (size_t)tn * GR_WARPS + warp) : 0u;   // prefetch
        if (key_value(tk) > thr) {                                   // something may be retained in this warp tile
            const long long row = t / T.nchunks;
            const long long b = T.b_lo + row;
            const int chunk = (int)(t - row * T.nchunks);
            const int jt = chunk * GR_CH + GR_JL0;
            const long long gbase = T.ML + b * T.r2 + jt;
            double v[GR_JPT];
#pragma unroll
            for (int m = 0; m < GR_JPT; ++m) {                      // all loads first (memory-level parallelism)
                const long long g = gbase + m * GR_MSTRIDE;
                v[m] = pq_neg_inf();
                if (jt + m * GR_MSTRIDE < T.r2 && g >= first && g < last) v[m] = __ldcs(logp_in + (g - first));
            }
            double roww = 0.0, rowwL = 0.0;                          // sum_m w, sum_m w*(theta_L - shift_L)
#pragma unroll
            for (int m = 0; m < GR_JPT; ++m) {
                if (v[m] > thr) {                                   // IP.py:3481
                    const double reps = exp(v[m] - thr);            // IP.py:3485
                    const double wt = trunc(reps);                  // IP.py:3486
                    const double thl = last_axis_value(S, L, jt + m * GR_MSTRIDE);
                    const double dl = thl - shift[L];
                    const double wdl = wt * dl;
                    A.wr = fma(wt, reps, A.wr);
                    A.nk += 1.0;
                    roww += wt;
                    rowwL += wdl;
                    A.s2[tri_index(KT, L, L)] = fma(wdl, dl, A.s2[tri_index(KT, L, L)]);
                    A.mn[L] = fmin(A.mn[L], thl);
                    A.mx[L] = fmax(A.mx[L], thl);
                    if (FUSE) s_col[jt + m * GR_MSTRIDE] += wt;      // position j of a row always belongs to this thread
                    if (keep_idx) {
                        const unsigned long long slot = atomicAdd(keep_count, 1ULL);
                        if ((long long)slot < keep_cap) { keep_idx[slot] = gbase + m * GR_MSTRIDE; keep_logp[slot] = v[m]; }
                    }
                }
            }
            if (roww > 0.0) {                                        // row part of the moments, once per row
                A.w += roww;
                A.s1[L] += rowwL;
                if (KT > 1) {
                    double th[KT];
                    grid_point<KT>(S, (unsigned long long)b, th, L);
#pragma unroll
                    for (int d = 0; d < L; ++d) {
                        const double dd = th[d] - shift[d];
                        const double wd = roww * dd;
                        A.s1[d] += wd;
#pragma unroll
                        for (int d2 = d; d2 < L; ++d2) A.s2[tri_index(KT, d, d2)] = fma(wd, th[d2] - shift[d2], A.s2[tri_index(KT, d, d2)]);
                        A.s2[tri_index(KT, d, L)] = fma(dd, rowwL, A.s2[tri_index(KT, d, L)]);
                        A.mn[d] = fmin(A.mn[d], th[d]);
                        A.mx[d] = fmax(A.mx[d], th[d]);
                    }
                }
            }
            if (FUSE && KT > 1) {
                const double rw = warp_sum(roww);                    // integers < 2^53: exact in fp64
                if (lane == 0 && rw > 0.0) atomicAdd(aux.rowsum + row, rw);
            }
        }
        tk = tk_next;
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
    if (FUSE) {
        __syncthreads();
        for (long long j = threadIdx.x; j < T.r2; j += blockDim.x) {
            const double cw = s_col[j];
            if (cw > 0.0) atomicAdd(aux.colsum + j, cw);
        }
    }
    double* mine = partial + (size_t)blockIdx.x * PQ_ACC_MAX;
    double r;
    r = blk_sum(A.w, s_red);  if (threadIdx.x == 0) mine[PEUQSE_ACC_W] = r;
    r = blk_sum(A.wr, s_red); if (threadIdx.x == 0) mine[PEUQSE_ACC_WEXP] = none ? 0.0 : r * exp(thr);
    r = blk_sum(A.nk, s_red); if (threadIdx.x == 0) mine[PEUQSE_ACC_NKEPT] = r;
    {
        int e = 0;
#pragma unroll
        for (int d = 0; d < KT; ++d) {
            r = blk_sum(A.s1[d], s_red);   if (threadIdx.x == 0) mine[PEUQSE_ACC_S1 + d] = r;
            r = blk_min(A.mn[d], s_red);   if (threadIdx.x == 0) mine[PEUQSE_ACC_MIN(KT) + d] = r;
            r = -blk_min(-A.mx[d], s_red); if (threadIdx.x == 0) mine[PEUQSE_ACC_MAX(KT) + d] = r;
#pragma unroll
            for (int d2 = d; d2 < KT; ++d2) {
                r = blk_sum(A.s2[e], s_red);
                if (threadIdx.x == 0) { mine[PEUQSE_ACC_S2(KT) + d * KT + d2] = r; mine[PEUQSE_ACC_S2(KT) + d2 * KT + d] = r; }
                ++e;
            }
        }
    }
    if (threadIdx.x == 0) {
        __threadfence();
        const unsigned int tkt = atomicAdd(ticket, 1u);
        s_last = (tkt == gridDim.x - 1);
    }
    __syncthreads();
    if (s_last) {
        __threadfence();
        constexpr int nacc = PEUQSE_ACC_DOUBLES(KT), mn0 = PEUQSE_ACC_MIN(KT), mx0 = PEUQSE_ACC_MAX(KT);
        for (int a = threadIdx.x; a < nacc; a += blockDim.x) {
            double rr = acc_out[a];
            for (int bb = 0; bb < (int)gridDim.x; ++bb) {          // fixed block order: deterministic
                const double x = __ldcg(partial + (size_t)bb * PQ_ACC_MAX + a);
                if (a >= mx0) rr = fmax(rr, x);
                else if (a >= mn0) rr = fmin(rr, x);
                else rr += x;
            }
            acc_out[a] = rr;
        }
        if (threadIdx.x == 0) {
            *ticket = 0;
            if (FUSE) *aux.valid = 1;
        }
    }
}

// ---------------------------------------------------------------------------------------------
// pass 2, TMA version: the retained tiles of logP are streamed through a shared-memory ring by bulk async copies
// (cp.async.bulk global->shared, completion on an mbarrier), so that HBM latency is hidden by the copy engine
// instead of by occupancy.  Per CTA: (1) the keys of its tiles are scanned once and the tiles with retained points
// are listed in shared memory, (2) one elected thread keeps P2_STAGES-1 copies in flight, (3) all warps consume a
// stage (a warp whose own tile maximum is below the threshold skips the arithmetic).
// ---------------------------------------------------------------------------------------------
#define P2_STAGES 4
#define P2_KLIST 1024
#define P2_STAGE_DOUBLES (GR_CH + 2)

struct TileSpan {          // element range of one (row, chunk) tile inside logp_in, and the 16-byte aligned copy range
    long long start;       // element index of j = chunk*GR_CH (may lie outside [0,count) for a rank's first/last row)
    long long lo, hi;      // valid elements [lo, hi)
    long long a0, a1;      // copied elements [a0, a1), both even; elements in [a1, hi) are loaded directly
    int len;               // number of j's in this chunk
};
__device__ __forceinline__ TileSpan tile_span(const GridTiling& T, long long first, long long count, long long t) {
    TileSpan s;
    const long long row = t / T.nchunks;
    const int chunk = (int)(t - row * T.nchunks);
    s.start = T.ML + (T.b_lo + row) * T.r2 + (long long)chunk * GR_CH - first;
    const long long rem = T.r2 - (long long)chunk * GR_CH;
    s.len = (int)(rem < GR_CH ? rem : GR_CH);
    s.lo = s.start > 0 ? s.start : 0;
    s.hi = s.start + s.len < count ? s.start + s.len : count;
    if (s.hi < s.lo) s.hi = s.lo;
    s.a0 = s.lo & ~1LL;
    s.a1 = (s.hi + 1) & ~1LL;
    if (s.a1 > count) s.a1 = s.hi & ~1LL;          // never read past the caller's buffer
    if (s.a1 < s.a0) s.a1 = s.a0;
    return s;
}

template <int KT, bool FUSE>
__global__ void __launch_bounds__(GR_THREADS) k_grid_pass2_tma(const __grid_constant__ DevSweep S,
                                                               const __grid_constant__ GridTiling T, long long first,
                                                               long long count, const double* __restrict__ logp_in,
                                                               const GridAux aux, const peuqse_best* __restrict__ best_in,
                                                               double log_rel, double* __restrict__ partial,
                                                               double* __restrict__ acc_out, unsigned int* __restrict__ ticket) {
    extern __shared__ __align__(16) unsigned char s_raw[];
    double* s_buf = (double*)s_raw;                                              // [P2_STAGES][P2_STAGE_DOUBLES]
    int* s_klist = (int*)(s_buf + P2_STAGES * P2_STAGE_DOUBLES);                 // [P2_KLIST] local tile slot | warp mask << 24
    unsigned int* s_col = (unsigned int*)(s_klist + P2_KLIST);                   // FUSE: [nchunks*GR_CH] column weight sums
    __shared__ __align__(8) unsigned long long s_full[P2_STAGES];
    __shared__ int s_nk;
    __shared__ double s_red[GR_WARPS];
    __shared__ bool s_last;
    constexpr int L = KT - 1;
    const long long best_index = best_in->index;
    const bool none = best_index < 0;
    const double thr = none ? pq_pos_inf() : best_in->logp + log_rel;   // IP.py:3477
    double shift[KT];
    grid_point<KT>(S, (unsigned long long)(none ? 0 : best_index), shift, KT);   // moments about the MAP point
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const long long ntiles = (T.b_hi - T.b_lo) * T.nchunks;
    if (threadIdx.x == 0) {
        for (int s = 0; s < P2_STAGES; ++s) mbar_init(&s_full[s], 1);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    if (FUSE)
        for (int a = threadIdx.x; a < T.nchunks * GR_CH; a += blockDim.x) s_col[a] = 0u;
    __syncthreads();
    GridAcc<KT> A;
    A.init();
    const long long nt_blk = none ? 0 : (ntiles - blockIdx.x + gridDim.x - 1) / gridDim.x;   // tiles blockIdx.x + i*gridDim.x
    unsigned int it = 0;                                             // stages consumed so far (ring position / phase)
    for (long long base = 0; base < nt_blk; base += P2_KLIST) {
        // (1) list the tiles of this batch that hold retained points
        if (threadIdx.x == 0) s_nk = 0;
        __syncthreads();
        for (int i = threadIdx.x; i < P2_KLIST && base + i < nt_blk; i += blockDim.x) {
            const long long t = blockIdx.x + (base + i) * gridDim.x;
            const uint4 k0 = __ldg((const uint4*)(aux.tilemax + (size_t)t * GR_WARPS));
            const uint4 k1 = __ldg((const uint4*)(aux.tilemax + (size_t)t * GR_WARPS) + 1);
            const unsigned int ks[8] = {k0.x, k0.y, k0.z, k0.w, k1.x, k1.y, k1.z, k1.w};
            int mask = 0;
#pragma unroll
            for (int w = 0; w < 8; ++w)
                if (key_value(ks[w]) > thr) mask |= 1 << w;
            if (mask) s_klist[atomicAdd(&s_nk, 1)] = i | (mask << 24);
        }
        __syncthreads();
        const int nk = s_nk;
        // (2) prologue: P2_STAGES-1 copies in flight
        if (threadIdx.x == 0) {
            for (int q = 0; q < P2_STAGES - 1 && q < nk; ++q) {
                const long long t = blockIdx.x + (base + (s_klist[q] & 0xffffff)) * gridDim.x;
                const TileSpan sp = tile_span(T, first, count, t);
                const int st = (it + q) % P2_STAGES;
                const uint32_t bytes = (uint32_t)((sp.a1 - sp.a0) * 8);
                mbar_arrive_expect_tx(&s_full[st], bytes);
                if (bytes) tma_load_1d(s_buf + st * P2_STAGE_DOUBLES, logp_in + sp.a0, bytes, &s_full[st]);
            }
        }
        // (3) consume
        for (int q = 0; q < nk; ++q, ++it) {
            const int st = it % P2_STAGES;
            if (threadIdx.x == 0 && q + P2_STAGES - 1 < nk) {        // refill the stage freed at the end of the previous iteration
                const int qn = q + P2_STAGES - 1;
                const long long t = blockIdx.x + (base + (s_klist[qn] & 0xffffff)) * gridDim.x;
                const TileSpan sp = tile_span(T, first, count, t);
                const int stn = (it + P2_STAGES - 1) % P2_STAGES;
                const uint32_t bytes = (uint32_t)((sp.a1 - sp.a0) * 8);
                mbar_arrive_expect_tx(&s_full[stn], bytes);
                if (bytes) tma_load_1d(s_buf + stn * P2_STAGE_DOUBLES, logp_in + sp.a0, bytes, &s_full[stn]);
            }
            const int ent = s_klist[q];
            const long long t = blockIdx.x + (base + (ent & 0xffffff)) * gridDim.x;
            const bool mine = (ent >> (24 + warp)) & 1;              // this warp's tile maximum is above the threshold
            while (!mbar_try_wait(&s_full[st], (it / P2_STAGES) & 1)) {}
            if (mine) {
                const TileSpan sp = tile_span(T, first, count, t);
                const long long row = t / T.nchunks;
                const int chunk = (int)(t - row * T.nchunks);
                const int jt = chunk * GR_CH + GR_JL0;
                const double* buf = s_buf + st * P2_STAGE_DOUBLES;
                double v[GR_JPT];
#pragma unroll
                for (int m = 0; m < GR_JPT; ++m) {
                    const long long e = sp.start + GR_JL0 + m * GR_MSTRIDE;
                    v[m] = pq_neg_inf();
                    if (GR_JL0 + m * GR_MSTRIDE < sp.len && e >= sp.lo && e < sp.hi)
                        v[m] = (e < sp.a1) ? buf[e - sp.a0] : __ldcs(logp_in + e);
                }
                double roww = 0.0, rowwL = 0.0;                      // sum_m w, sum_m w*(theta_L - shift_L)
#pragma unroll
                for (int m = 0; m < GR_JPT; ++m) {
                    if (v[m] > thr) {                               // IP.py:3481
                        const double reps = exp(v[m] - thr);        // IP.py:3485
                        const double wt = trunc(reps);              // IP.py:3486
                        const double thl = last_axis_value(S, L, jt + m * GR_MSTRIDE);
                        const double dl = thl - shift[L];
                        const double wdl = wt * dl;
                        A.wr = fma(wt, reps, A.wr);
                        A.nk += 1.0;
                        roww += wt;
                        rowwL += wdl;
                        A.s2[tri_index(KT, L, L)] = fma(wdl, dl, A.s2[tri_index(KT, L, L)]);
                        A.mn[L] = fmin(A.mn[L], thl);
                        A.mx[L] = fmax(A.mx[L], thl);
                        if (FUSE) s_col[jt + m * GR_MSTRIDE] += (unsigned int)wt;   // j always belongs to this thread
                    }
                }
                if (roww > 0.0) {                                    // row part of the moments, once per row
                    A.w += roww;
                    A.s1[L] += rowwL;
                    if (KT > 1) {
                        double th[KT];
                        grid_point<KT>(S, (unsigned long long)(T.b_lo + row), th, L);
#pragma unroll
                        for (int d = 0; d < L; ++d) {
                            const double dd = th[d] - shift[d];
                            const double wd = roww * dd;
                            A.s1[d] += wd;
#pragma unroll
                            for (int d2 = d; d2 < L; ++d2) A.s2[tri_index(KT, d, d2)] = fma(wd, th[d2] - shift[d2], A.s2[tri_index(KT, d, d2)]);
                            A.s2[tri_index(KT, d, L)] = fma(dd, rowwL, A.s2[tri_index(KT, d, L)]);
                            A.mn[d] = fmin(A.mn[d], th[d]);
                            A.mx[d] = fmax(A.mx[d], th[d]);
                        }
                    }
                }
                if (FUSE && KT > 1) {
                    const double rw = warp_sum(roww);                // integers < 2^53: exact in fp64
                    if (lane == 0 && rw > 0.0) atomicAdd(aux.rowsum + row, rw);
                }
            }
            __syncthreads();                                         // stage st may be refilled from now on
        }
    }
    for (long long g = T.slab_first + (long long)blockIdx.x * GR_THREADS + threadIdx.x; g < T.slab_first + T.slab_count && !none;
         g += (long long)gridDim.x * GR_THREADS) {
        const double v = __ldcs(logp_in + (g - first));
        if (v > thr) {
            double th[KT];
            grid_point<KT>(S, (unsigned long long)g, th, KT);
            A.add(v, thr, th, shift);
        }
    }
    if (FUSE) {
        __syncthreads();
        for (long long j = threadIdx.x; j < T.r2; j += blockDim.x) {
            const unsigned int cw = s_col[j];
            if (cw) atomicAdd(aux.colsum + j, (double)cw);
        }
    }
    double* mine_p = partial + (size_t)blockIdx.x * PQ_ACC_MAX;
    double r;
    r = blk_sum(A.w, s_red);  if (threadIdx.x == 0) mine_p[PEUQSE_ACC_W] = r;
    r = blk_sum(A.wr, s_red); if (threadIdx.x == 0) mine_p[PEUQSE_ACC_WEXP] = none ? 0.0 : r * exp(thr);
    r = blk_sum(A.nk, s_red); if (threadIdx.x == 0) mine_p[PEUQSE_ACC_NKEPT] = r;
    {
        int e = 0;
#pragma unroll
        for (int d = 0; d < KT; ++d) {
            r = blk_sum(A.s1[d], s_red);   if (threadIdx.x == 0) mine_p[PEUQSE_ACC_S1 + d] = r;
            r = blk_min(A.mn[d], s_red);   if (threadIdx.x == 0) mine_p[PEUQSE_ACC_MIN(KT) + d] = r;
            r = -blk_min(-A.mx[d], s_red); if (threadIdx.x == 0) mine_p[PEUQSE_ACC_MAX(KT) + d] = r;
#pragma unroll
            for (int d2 = d; d2 < KT; ++d2) {
                r = blk_sum(A.s2[e], s_red);
                if (threadIdx.x == 0) { mine_p[PEUQSE_ACC_S2(KT) + d * KT + d2] = r; mine_p[PEUQSE_ACC_S2(KT) + d2 * KT + d] = r; }
                ++e;
            }
        }
    }
    if (threadIdx.x == 0) {
        __threadfence();
        const unsigned int tkt = atomicAdd(ticket, 1u);
        s_last = (tkt == gridDim.x - 1);
    }
    __syncthreads();
    if (s_last) {
        __threadfence();
        constexpr int nacc = PEUQSE_ACC_DOUBLES(KT), mn0 = PEUQSE_ACC_MIN(KT), mx0 = PEUQSE_ACC_MAX(KT);
        for (int a = threadIdx.x; a < nacc; a += blockDim.x) {
            double rr = acc_out[a];
            for (int bb = 0; bb < (int)gridDim.x; ++bb) {          // fixed block order: deterministic
                const double x = __ldcg(partial + (size_t)bb * PQ_ACC_MAX + a);
                if (a >= mx0) rr = fmax(rr, x);
                else if (a >= mn0) rr = fmin(rr, x);
                else rr += x;
            }
            acc_out[a] = rr;
        }
        if (threadIdx.x == 0) *ticket = 0;
    }
}

// ---------------------------------------------------------------------------------------------
// pass 2, TMA version 2 (the one the C2 step uses): per retained point only exp, trunc and four accumulations remain;
// everything separable comes from the per-column sums (shared memory -> global) and per-row sums (global atomics):
//     W = sum_j colsum_j,  S1_L = sum_j colsum_j dL_j,  S2_LL = sum_j colsum_j dL_j^2,  min/max_L = support of colsum,
//     S1_d = sum_b rowsum_b dd_b,  S2_dd' = sum_b rowsum_b dd_b dd'_b,  S2_dL = sum_b dd_b rowsumL_b   (d, d' < L)
// which the single-CTA k_grid_finish evaluates (it also handles the slab).
// ---------------------------------------------------------------------------------------------
// exp(x) for 0 <= x < 700: the usual reduction x = n ln2 + r, |r| <= ln2/2, degree-13 Taylor polynomial in Horner form,
// scaling through the exponent field; <= 1 ulp like libdevice's exp but without its range / denormal handling.
__constant__ double c_exp_pos[16] = {1.6059043836821613e-10 /* 1/13! */, 2.08767569878681e-09, 2.505210838544172e-08,
                                     2.755731922398589e-07, 2.7557319223985893e-06, 2.48015873015873e-05,
                                     1.984126984126984e-04, 1.388888888888889e-03, 8.333333333333333e-03,
                                     4.1666666666666664e-02, 1.6666666666666666e-01, 0.5,
                                     1.4426950408889634 /* log2(e) */, -6.93147180369123816490e-01 /* -ln2 hi */,
                                     -1.90821492927058770002e-10 /* -ln2 lo */, 0.0};
// (the coefficients sit in the constant bank / uniform registers: as immediates each costs two UMOVs per use)
__device__ __forceinline__ double pq_exp_pos(double x) {
    const double t = fma(x, c_exp_pos[12], 6755399441055744.0);
    const int n = __double2loint(t);
    const double fn = t - 6755399441055744.0;
    double r = fma(fn, c_exp_pos[13], x);
    r = fma(fn, c_exp_pos[14], r);
    double p = c_exp_pos[0];
#pragma unroll
    for (int i = 1; i < 12; ++i) p = fma(p, r, c_exp_pos[i]);
    p = fma(p, r, 1.0);
    p = fma(p, r, 1.0);
    return __hiloint2double(__double2hiint(p) + (n << 20), __double2loint(p));
}
// the same value within 1 ulp from a 64-entry table of 2^(j/64) in shared memory (eval.cuh: c_exp_tab64) and a degree-5
// polynomial on |r| <= ln2/128: 10 fp64 instructions instead of 18 (pass 2 spends most of its fp64 work in this exp)
__constant__ double c_exp_pos_tab[8] = {92.332482616893656 /* 64/ln2 */, -0.01083042469326756 /* -(ln2/64) hi */,
                                        -2.9815858269852933e-12 /* lo */, 8.3333333333333332e-03, 4.1666666666666664e-02,
                                        1.6666666666666666e-01, 0.5, 1.0};
__device__ __forceinline__ double pq_exp_pos_tab(const double* __restrict__ tab, double x) {
    const double t = fma(x, c_exp_pos_tab[0], 6755399441055744.0);
    const int k = __double2loint(t);
    const double fn = t - 6755399441055744.0;
    double r = fma(fn, c_exp_pos_tab[1], x);
    r = fma(fn, c_exp_pos_tab[2], r);
    double q = fma(r, c_exp_pos_tab[3], c_exp_pos_tab[4]);
    q = fma(r, q, c_exp_pos_tab[5]);
    q = fma(r, q, c_exp_pos_tab[6]);
    q = fma(r, q, c_exp_pos_tab[7]);
    const double T = tab[k & 63];
    const double v = fma(T, r * q, T);
    return __hiloint2double(__double2hiint(v) + ((k >> 6) << 20), __double2loint(v));
}

// ---------------------------------------------------------------------------------------------
// pass 2, warp-autonomous version (the one the C2 step uses).  Its predecessor moved whole 2048-point tiles through one
// CTA-wide ring: one __syncthreads per tile (ncu, profiles/r2_ncu_c2_pass2_before_after.md: 22 % of the warp samples waiting
// at that barrier).  Here every WARP is its own consumer: it walks a strided sequence of warp tiles, tests their pass-1
// maxima 32 at a time (one candidate per lane, one ballot), and streams only the retained ones through its private
// 4-stage ring of 2 KB bulk copies (cp.async.bulk + mbarrier) -- no CTA-level synchronisation inside the loop, copies of up
// to four tiles in flight per warp.  Per retained point: exp, trunc, the weight sums of its row and column; the column
// of a point is no longer owned by one thread, so the per-CTA column sums are shared-memory reductions (RED.shared).
// The last-axis offset of a column is rebuilt from one int-to-double conversion per warp tile (the eight columns of a
// lane are 16 half-steps apart: exact additions), not from one per point.  Integer weights are exact in any order, the
// products accumulate per lane in the order the warp meets its tiles (fixed assignment: deterministic).
// ---------------------------------------------------------------------------------------------
#define P2W_STAGES 4
#define P2W_STAGE_DOUBLES 260          // 256 points + alignment slack (copies start and end on 16-byte boundaries)

template <int KT>
__global__ void __launch_bounds__(GR_THREADS, 2) k_grid_pass2_w(const __grid_constant__ DevSweep S,
                                                                const __grid_constant__ GridTiling T, long long first,
                                                                long long count, const double* __restrict__ logp_in,
                                                                const GridAux aux, const peuqse_best* __restrict__ best_in,
                                                                double log_rel, double* __restrict__ partial,
                                                                double* __restrict__ acc_out, unsigned int* __restrict__ ticket) {
    extern __shared__ __align__(16) unsigned char s_raw[];
    double* s_buf = (double*)s_raw;                                              // [GR_WARPS][P2W_STAGES][P2W_STAGE_DOUBLES]
    unsigned int* s_col = (unsigned int*)(s_buf + GR_WARPS * P2W_STAGES * P2W_STAGE_DOUBLES);   // [nchunks*GR_CH] column weight sums
    __shared__ __align__(8) unsigned long long s_full[GR_WARPS][P2W_STAGES];
    __shared__ long long s_mstart[GR_WARPS][P2W_STAGES];                         // per stage: first element of the warp tile ...
    __shared__ int2 s_mrc[GR_WARPS][P2W_STAGES];                                 // ... its row and first column
    __shared__ double s_red[GR_WARPS];
    __shared__ double s_tab[64];
    __shared__ bool s_last;
    constexpr int L = KT - 1;
    const long long best_index = best_in->index;
    const bool none = best_index < 0;
    const double thr = none ? pq_pos_inf() : best_in->logp + log_rel;   // IP.py:3477
    double shift[KT];
    grid_point<KT>(S, (unsigned long long)(none ? 0 : best_index), shift, KT);   // moments about the MAP point
    const double shiftL = shift[L], cL = S.center[L], hL = S.interval[L];
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    if (threadIdx.x < 64) s_tab[threadIdx.x] = c_exp_tab64[threadIdx.x];
    const int wpr = T.nchunks * GR_WARPS;                                        // warp tiles per row
    const long long nwt = none ? 0 : (T.b_hi - T.b_lo) * wpr;                    // warp tiles of the range
    if (lane == 0) {
        for (int s = 0; s < P2W_STAGES; ++s) mbar_init(&s_full[warp][s], 1);
    }
    if (threadIdx.x == 0) asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    for (int a = threadIdx.x; a < T.nchunks * GR_CH; a += blockDim.x) s_col[a] = 0u;
    __syncthreads();
    double* wbuf = s_buf + (size_t)warp * P2W_STAGES * P2W_STAGE_DOUBLES;
    const double shl = (lane & 1) ? hL : -hL;                      // c + (h i) or c - (h i), CombinationGeneratorModule.py:59,71
    const long long r2 = T.r2, base_el = T.ML + T.b_lo * T.r2 - first;
    double a_wr = 0.0;
    int a_nk = 0;
    unsigned int issued = 0, consumed = 0;
    const long long gw = (long long)blockIdx.x * GR_WARPS + warp, nw = (long long)gridDim.x * GR_WARPS;

    // element range of a warp tile inside logp_in: [lo, hi) valid points, [a0, a1) the 16-byte aligned copy range
    auto span = [&](long long start, int col, long long& lo, long long& hi, long long& a0, long long& a1) {
        const long long rem = r2 - col;
        const int len = (int)(rem < 256 ? (rem > 0 ? rem : 0) : 256);
        lo = start > 0 ? start : 0;
        hi = start + len < count ? start + len : count;
        if (hi < lo) hi = lo;
        a0 = lo & ~1LL;
        a1 = (hi + 1) & ~1LL;
        if (a1 > count) a1 = hi & ~1LL;          // never read past the caller's buffer
        if (a1 < a0) a1 = a0;
        return len;
    };

    auto consume = [&]() {
        const int st = (int)(consumed % P2W_STAGES);
        const long long start = s_mstart[warp][st];
        const int2 rc = s_mrc[warp][st];
        long long lo, hi, a0, a1;
        const int len = span(start, rc.y, lo, hi, a0, a1);
        while (!mbar_try_wait(&s_full[warp][st], (consumed / P2W_STAGES) & 1)) {}
        const double* buf = wbuf + st * P2W_STAGE_DOUBLES;
        const int j0 = rc.y + lane;                                 // column of this lane's first point
        const double d0 = (double)((j0 >> 1) + 1);                  // its half-step count; the lane's other points: + 16 m
        // last-axis moment of the tile's points about the MAP point: sum wt ((cL + shl i) - shiftL) with i = d0 + 16 m taken as
        // (cL - shiftL) sum wt + shl (d0 sum wt + 16 sum m wt): the sums over m are of integers (exact), one FMA per point
        double roww = 0.0, rowm = 0.0;
        const bool whole = len == 256 && lo == start && hi == start + 256 && a1 >= hi;
        if (whole) {
            const double* bufw = buf + (start - a0) + lane;
            double x[GR_JPT], reps[GR_JPT];
            bool kp[GR_JPT];
#pragma unroll
            for (int m = 0; m < GR_JPT; ++m) {
                const double v = bufw[m * 32];
                kp[m] = v > thr;                                   // IP.py:3481
                x[m] = kp[m] ? (v - thr) : 0.0;
            }
#pragma unroll
            for (int m = 0; m < GR_JPT; ++m) reps[m] = pq_exp_pos_tab(s_tab, x[m]);          // IP.py:3485
#pragma unroll
            for (int m = 0; m < GR_JPT; ++m) {
                const double wt_ = kp[m] ? trunc(reps[m]) : 0.0;                  // IP.py:3486
                a_wr = fma(wt_, reps[m], a_wr);
                a_nk += kp[m] ? 1 : 0;
                roww += wt_;
                rowm = fma(wt_, (double)m, rowm);
                if (kp[m]) atomicAdd(&s_col[j0 + m * 32], (unsigned int)wt_);
            }
        } else {
#pragma unroll
            for (int m = 0; m < GR_JPT; ++m) {
                const long long e = start + lane + m * 32;
                double v = pq_neg_inf();
                if (lane + m * 32 < len && e >= lo && e < hi) v = (e < a1) ? buf[e - a0] : __ldcs(logp_in + e);
                if (v > thr) {
                    const double r_ = pq_exp_pos_tab(s_tab, v - thr);
                    const double wt_ = trunc(r_);
                    a_wr = fma(wt_, r_, a_wr);
                    a_nk += 1;
                    roww += wt_;
                    rowm = fma(wt_, (double)m, rowm);
                    atomicAdd(&s_col[j0 + m * 32], (unsigned int)wt_);
                }
            }
        }
        const double rw = warp_sum(roww);                           // integers < 2^53: exact in fp64
        if (rw > 0.0) {
            const double rowi = fma(d0, roww, 16.0 * rowm);         // sum wt i: integers, exact below 2^53
            const double rwl = warp_sum(fma(shl, rowi, (cL - shiftL) * roww));
            if (lane == 0) {
                atomicAdd(aux.rowsum + rc.x, rw);
                atomicAdd(aux.rowsumL + rc.x, rwl);
            }
        }
        __syncwarp();                                               // every lane has read the stage: it may be refilled
        ++consumed;
    };

    for (long long base = 0; gw + base * nw < nwt; base += 32) {
        // one candidate warp tile per lane: retained iff the upper bound of its pass-1 maximum is above the threshold;
        // the lane also resolves its candidate's row and first column (one division per lane per 32 candidates)
        const long long mywt = gw + (base + lane) * nw;
        bool keep = false;
        int myrow = 0, mycol = 0;
        if (mywt < nwt) {
            keep = key_value(__ldg(aux.tilemax + mywt)) > thr;
            myrow = (nwt < 0x7fffffffLL) ? (int)((unsigned int)mywt / (unsigned int)wpr) : (int)(mywt / wpr);
            mycol = (int)(mywt - (long long)myrow * wpr) * 256;     // (chunk * 8 + w) * 256 = chunk * GR_CH + w * 256
        }
        unsigned int mask = __ballot_sync(0xffffffffu, keep);
        while (mask) {
            if (issued - consumed == P2W_STAGES) consume();
            const int b = __ffs(mask) - 1;
            mask &= mask - 1;
            const int row = __shfl_sync(0xffffffffu, myrow, b), col = __shfl_sync(0xffffffffu, mycol, b);
            if (lane == 0) {
                const long long start = base_el + (long long)row * r2 + col;
                long long lo, hi, a0, a1;
                span(start, col, lo, hi, a0, a1);
                const int st = (int)(issued % P2W_STAGES);
                const uint32_t bytes = (uint32_t)((a1 - a0) * 8);
                s_mstart[warp][st] = start;
                s_mrc[warp][st] = make_int2(row, col);
                mbar_arrive_expect_tx(&s_full[warp][st], bytes);
                if (bytes) tma_load_1d(wbuf + st * P2W_STAGE_DOUBLES, logp_in + a0, bytes, &s_full[warp][st]);
            }
            __syncwarp();
            ++issued;
        }
    }
    while (consumed < issued) consume();
    __syncthreads();
    for (long long j = threadIdx.x; j < T.r2; j += blockDim.x) {
        const unsigned int cw = s_col[j];
        if (cw) atomicAdd(aux.colsum + j, (double)cw);
    }
    double* mine_p = partial + (size_t)blockIdx.x * PQ_ACC_MAX;
    double r;
    r = blk_sum(a_wr, s_red); if (threadIdx.x == 0) mine_p[0] = none ? 0.0 : r * exp(thr);
    r = blk_sum((double)a_nk, s_red); if (threadIdx.x == 0) mine_p[1] = r;
    if (threadIdx.x == 0) {
        __threadfence();
        const unsigned int tkt = atomicAdd(ticket, 1u);
        s_last = (tkt == gridDim.x - 1);
    }
    __syncthreads();
    if (s_last) {                                                    // fixed assignment + fixed tree: deterministic
        __threadfence();
        double r0 = 0.0, r1 = 0.0;
        for (int bb = threadIdx.x; bb < (int)gridDim.x; bb += blockDim.x) {
            r0 += __ldcg(partial + (size_t)bb * PQ_ACC_MAX);
            r1 += __ldcg(partial + (size_t)bb * PQ_ACC_MAX + 1);
        }
        r0 = blk_sum(r0, s_red);
        r1 = blk_sum(r1, s_red);
        if (threadIdx.x == 0) {
            acc_out[PEUQSE_ACC_WEXP] += r0;
            acc_out[PEUQSE_ACC_NKEPT] += r1;
            *ticket = 0;
        }
    }
}

// separable moments and min/max from the row / column sums + the slab, accumulated into acc_out (a few CTAs; partials
// are combined by the last CTA in fixed order, so the result is deterministic)
template <int KT>
__global__ void __launch_bounds__(GR_THREADS) k_grid_finish(const __grid_constant__ DevSweep S, const __grid_constant__ GridTiling T,
                                                            long long first, const double* __restrict__ logp_in, const GridAux aux,
                                                            const peuqse_best* __restrict__ best_in, double log_rel,
                                                            double* __restrict__ partial, double* __restrict__ acc_out,
                                                            unsigned int* __restrict__ ticket) {
    __shared__ double s_red[GR_WARPS];
    __shared__ bool s_last;
    constexpr int L = KT - 1;
    const long long best_index = best_in->index;
    if (best_index < 0) return;
    const double thr = best_in->logp + log_rel;
    double shift[KT];
    grid_point<KT>(S, (unsigned long long)best_index, shift, KT);
    GridAcc<KT> A;
    A.init();
    const long long stride = (long long)gridDim.x * blockDim.x, t0 = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    for (long long j = t0; j < T.r2; j += stride) {
        const double cw = aux.colsum[j];
        if (cw > 0.0) {
            const double thl = last_axis_value(S, L, j);
            const double dl = thl - shift[L];
            const double wd = cw * dl;
            A.w += cw;
            A.s1[L] += wd;
            A.s2[tri_index(KT, L, L)] = fma(wd, dl, A.s2[tri_index(KT, L, L)]);
            A.mn[L] = fmin(A.mn[L], thl);
            A.mx[L] = fmax(A.mx[L], thl);
        }
    }
    if (KT > 1) {
        for (long long row = t0; row < T.b_hi - T.b_lo; row += stride) {
            const double rw = aux.rowsum[row];
            if (rw > 0.0) {
                const double rwl = aux.rowsumL[row];
                double th[KT];
                grid_point<KT>(S, (unsigned long long)(T.b_lo + row), th, L);
#pragma unroll
                for (int d = 0; d < L; ++d) {
                    const double dd = th[d] - shift[d];
                    const double wd = rw * dd;
                    A.s1[d] += wd;
#pragma unroll
                    for (int d2 = d; d2 < L; ++d2) A.s2[tri_index(KT, d, d2)] = fma(wd, th[d2] - shift[d2], A.s2[tri_index(KT, d, d2)]);
                    A.s2[tri_index(KT, d, L)] = fma(dd, rwl, A.s2[tri_index(KT, d, L)]);
                    A.mn[d] = fmin(A.mn[d], th[d]);
                    A.mx[d] = fmax(A.mx[d], th[d]);
                }
            }
        }
    }
    for (long long g = T.slab_first + t0; g < T.slab_first + T.slab_count; g += stride) {
        const double v = __ldcs(logp_in + (g - first));
        if (v > thr) {
            double th[KT];
            grid_point<KT>(S, (unsigned long long)g, th, KT);
            A.add(v, thr, th, shift);
        }
    }
    double* mine = partial + (size_t)blockIdx.x * PQ_ACC_MAX;
    double r;
    r = blk_sum(A.w, s_red);  if (threadIdx.x == 0) mine[PEUQSE_ACC_W] = r;
    r = blk_sum(A.wr, s_red); if (threadIdx.x == 0) mine[PEUQSE_ACC_WEXP] = r * exp(thr);
    r = blk_sum(A.nk, s_red); if (threadIdx.x == 0) mine[PEUQSE_ACC_NKEPT] = r;
    {
        int e = 0;
#pragma unroll
        for (int d = 0; d < KT; ++d) {
            r = blk_sum(A.s1[d], s_red);   if (threadIdx.x == 0) mine[PEUQSE_ACC_S1 + d] = r;
            r = blk_min(A.mn[d], s_red);   if (threadIdx.x == 0) mine[PEUQSE_ACC_MIN(KT) + d] = r;
            r = -blk_min(-A.mx[d], s_red); if (threadIdx.x == 0) mine[PEUQSE_ACC_MAX(KT) + d] = r;
#pragma unroll
            for (int d2 = d; d2 < KT; ++d2) {
                r = blk_sum(A.s2[e], s_red);
                if (threadIdx.x == 0) { mine[PEUQSE_ACC_S2(KT) + d * KT + d2] = r; mine[PEUQSE_ACC_S2(KT) + d2 * KT + d] = r; }
                ++e;
            }
        }
    }
    if (threadIdx.x == 0) {
        __threadfence();
        const unsigned int tkt = atomicAdd(ticket, 1u);
        s_last = (tkt == gridDim.x - 1);
    }
    __syncthreads();
    if (s_last) {
        __threadfence();
        constexpr int nacc = PEUQSE_ACC_DOUBLES(KT), mn0 = PEUQSE_ACC_MIN(KT), mx0 = PEUQSE_ACC_MAX(KT);
        for (int a = threadIdx.x; a < nacc; a += blockDim.x) {
            double x[16];
#pragma unroll
            for (int bb = 0; bb < 16; ++bb) x[bb] = (bb < (int)gridDim.x) ? __ldcg(partial + (size_t)bb * PQ_ACC_MAX + a) : 0.0;
            double rr = acc_out[a];
#pragma unroll
            for (int bb = 0; bb < 16; ++bb) {
                if (bb < (int)gridDim.x) {
                    if (a >= mx0) rr = fmax(rr, x[bb]);
                    else if (a >= mn0) rr = fmin(rr, x[bb]);
                    else rr += x[bb];
                }
            }
            acc_out[a] = rr;
        }
        if (threadIdx.x == 0) *ticket = 0;
    }
}

// ---------------------------------------------------------------------------------------------
// pass 3
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

// adds w to bins[bin] (bin < 0: nothing); lanes of a warp that hit the same bin are summed first, so that the 64-bit
// shared-memory atomic (a CAS loop) sees one update per distinct bin per warp instead of 32 colliding ones.
// Must be called by all 32 lanes.
__device__ __forceinline__ void warp_bin_add(unsigned long long* bins, int bin, double w) {
    unsigned int todo = __ballot_sync(0xffffffffu, bin >= 0 && w > 0.0);
    const int lane = threadIdx.x & 31;
    while (todo) {
        const int leader = __ffs(todo) - 1;
        const int b0 = __shfl_sync(0xffffffffu, bin, leader);
        const bool mine = (bin == b0) && ((todo >> lane) & 1u);
        const double sum = warp_sum(mine ? w : 0.0);                 // integer-valued doubles: exact
        if (lane == leader) atomicAdd(bins + b0, (unsigned long long)sum);
        todo &= ~__ballot_sync(0xffffffffu, mine);
    }
}

// (a) bins the row / column weight sums left by the fused pass 2 (and clears them); reads only the slab of logP
template <int KT>
__global__ void __launch_bounds__(GR_THREADS) k_grid_bin(const __grid_constant__ DevSweep S, const __grid_constant__ GridTiling T,
                                                         long long first, const double* __restrict__ logp_in, const GridAux aux,
                                                         const peuqse_best* __restrict__ best_in, double log_rel,
                                                         const double* __restrict__ edges, int nbins,
                                                         unsigned long long* __restrict__ hist) {
    extern __shared__ unsigned long long sh[];                      // [KT][nbins]
    constexpr int L = KT - 1;
    if (best_in->index < 0) return;
    const double thr = best_in->logp + log_rel;
    for (int a = threadIdx.x; a < KT * nbins; a += blockDim.x) sh[a] = 0ULL;
    __syncthreads();
    const long long stride = (long long)gridDim.x * blockDim.x, t0 = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    // every loop runs the same number of iterations in all lanes of a warp (warp_bin_add is a warp collective)
    for (long long j0 = (long long)blockIdx.x * blockDim.x; j0 < T.r2; j0 += stride) {
        const long long j = j0 + threadIdx.x;
        const double cw = (j < T.r2) ? aux.colsum[j] : 0.0;
        const int bin = (cw > 0.0) ? hist_bin(edges + L * (nbins + 1), nbins, last_axis_value(S, L, j)) : -1;
        warp_bin_add(sh + L * nbins, bin, cw);
    }
    if (KT > 1) {
        const long long rows = T.b_hi - T.b_lo;
        for (long long r0 = (long long)blockIdx.x * blockDim.x; r0 < rows; r0 += stride) {
            const long long row = r0 + threadIdx.x;
            const double rw = (row < rows) ? aux.rowsum[row] : 0.0;
            double th[KT];
            if (rw > 0.0) grid_point<KT>(S, (unsigned long long)(T.b_lo + row), th, L);
#pragma unroll
            for (int d = 0; d < L; ++d) {
                const int bin = (rw > 0.0) ? hist_bin(edges + d * (nbins + 1), nbins, th[d]) : -1;
                warp_bin_add(sh + d * nbins, bin, rw);
            }
        }
    }
    for (long long g0 = T.slab_first + (long long)blockIdx.x * blockDim.x; g0 < T.slab_first + T.slab_count; g0 += stride) {
        const long long g = g0 + threadIdx.x;
        const double v = (g < T.slab_first + T.slab_count) ? __ldcs(logp_in + (g - first)) : pq_neg_inf();
        const double w = (v > thr) ? trunc(exp(v - thr)) : 0.0;
        double th[KT];
        if (w > 0.0) grid_point<KT>(S, (unsigned long long)g, th, KT);
#pragma unroll
        for (int d = 0; d < KT; ++d) {
            const int bin = (w > 0.0) ? hist_bin(edges + d * (nbins + 1), nbins, th[d]) : -1;
            warp_bin_add(sh + d * nbins, bin, w);
        }
    }
    __syncthreads();
    for (int a = threadIdx.x; a < KT * nbins; a += blockDim.x)
        if (sh[a]) atomicAdd(hist + a, sh[a]);
}

// (b) stand-alone version (no fused sums available): reads the retained tiles of logP again
template <int KT>
__global__ void __launch_bounds__(GR_THREADS) k_grid_hist(const __grid_constant__ DevSweep S, const __grid_constant__ GridTiling T,
                                                          long long first, long long count, const double* __restrict__ logp_in,
                                                          const GridAux aux, const peuqse_best* __restrict__ best_in,
                                                          double log_rel, const double* __restrict__ edges, int nbins,
                                                          unsigned long long* __restrict__ hist) {
    extern __shared__ unsigned long long sh[];                      // [KT][nbins] bins, then colsum[nchunks*GR_CH] doubles
    constexpr int L = KT - 1;
    if (best_in->index < 0) return;
    const double thr = best_in->logp + log_rel;
    double* colsum = (double*)(sh + KT * nbins);
    const int ncol = T.nchunks * GR_CH;
    for (int a = threadIdx.x; a < KT * nbins; a += blockDim.x) sh[a] = 0ULL;
    for (int a = threadIdx.x; a < ncol; a += blockDim.x) colsum[a] = 0.0;
    __syncthreads();
    const int warp = threadIdx.x >> 5;
    const long long last = first + count;
    const long long ntiles = (T.b_hi - T.b_lo) * T.nchunks;
    for (long long t = blockIdx.x; t < ntiles; t += gridDim.x) {
        if (!(key_value(__ldg(aux.tilemax + (size_t)t * GR_WARPS + warp)) > thr)) continue;
        const long long row = t / T.nchunks;
        const long long b = T.b_lo + row;
        const int chunk = (int)(t - row * T.nchunks);
        const int jt = chunk * GR_CH + GR_JL0;
        const long long gbase = T.ML + b * T.r2 + jt;
        double v[GR_JPT];
#pragma unroll
        for (int m = 0; m < GR_JPT; ++m) {
            const long long g = gbase + m * GR_MSTRIDE;
            v[m] = pq_neg_inf();
            if (jt + m * GR_MSTRIDE < T.r2 && g >= first && g < last) v[m] = __ldcs(logp_in + (g - first));
        }
        double roww = 0.0;
#pragma unroll
        for (int m = 0; m < GR_JPT; ++m) {
            if (v[m] > thr) {
                const double w = trunc(exp(v[m] - thr));
                colsum[jt + m * GR_MSTRIDE] += w;
                roww += w;
            }
        }
        if (KT > 1) {
            roww = warp_sum(roww);
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
    __syncthreads();
    for (long long j = threadIdx.x; j < T.r2; j += blockDim.x) {
        const double cw = colsum[j];
        if (cw > 0.0) {
            const int bin = hist_bin(edges + L * (nbins + 1), nbins, last_axis_value(S, L, j));
            if (bin >= 0) atomicAdd(sh + L * nbins + bin, (unsigned long long)cw);
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
        T->dyn = 0;
        return true;
    }
    T->b_lo = (g0 - ML) / r2;
    T->b_hi = (end - ML + r2 - 1) / r2;
    const long long rows = T->b_hi - T->b_lo;
    const long long units = (rows + GR_RB - 1) / GR_RB;  // pass-1 work units per chunk
    long long nrg = resident_blocks / T->nchunks;       // one resident wave: no tail
    if (nrg > units) nrg = units;
    if (nrg < 1) nrg = 1;
    if (nrg * T->nchunks > PQ_SWEEP_MAX_BLOCKS) nrg = PQ_SWEEP_MAX_BLOCKS / T->nchunks;
    T->dyn = T->nchunks <= PQ_GRID_MAX_CTR ? 1 : 0;
    T->nrg = (int)nrg;
    return true;
}

template <class K>
static int resident_blocks(const peuqse_problem* p, K kern, size_t smem) {
    int per_sm = 1;
    if (cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, kern, GR_THREADS, smem) != cudaSuccess || per_sm < 1) per_sm = 1;
    return per_sm * p->sm_count;
}

static int gcd_int(int a, int b) { while (b) { int t = a % b; a = b; b = t; } return a; }
// number of CTAs for the flat tile loops: at most one resident wave, coprime to nchunks, not more than the tiles
static int flat_grid(const GridTiling& T, int resident) {
    long long ntiles = (T.b_hi - T.b_lo) * T.nchunks;
    long long g = resident;
    if (g > ntiles) g = ntiles;
    if (g > PQ_SWEEP_MAX_BLOCKS) g = PQ_SWEEP_MAX_BLOCKS;
    if (g < 1) g = 1;
    while (g > 1 && gcd_int((int)g, T.nchunks) != 1) --g;
    return (int)g;
}

static bool fast_model_ok(const peuqse_problem* p) {
    const DevProblem& d = p->d;
    if (d.fast_n <= 0) return false;
    if (d.model_id == PEUQSE_MODEL_QUADRATIC) return true;
    if (d.model_id == PEUQSE_MODEL_LINEAR) return d.k == 2 || d.k == 3 || d.k == 4 || d.k == 6;
    return false;
}

static bool grid_k_ok(int k) { return k == 1 || k == 2 || k == 3 || k == 4 || k == 6; }

static inline size_t align_up(size_t x, size_t a) { return (x + a - 1) / a * a; }

// layout of the auxiliary arrays for this range; returns their total size (0 when the specialised kernels do not apply)
static long long grid_aux_layout(const peuqse_problem* p, const DevSweep& S, long long first, long long count, char* base,
                                 GridAux* aux) {
    GridTiling T;
    if (!p || !fast_model_ok(p) || !grid_k_ok(S.k) || !grid_tiling(p, S, first, count, 1, &T)) return 0;
    const size_t rows = (size_t)(T.b_hi - T.b_lo);
    size_t off = 0;
    const size_t o_tile = off; off = align_up(off + (rows * T.nchunks * GR_WARPS + 1) * sizeof(unsigned int), 256);
    const size_t o_col = off;  off = align_up(off + (size_t)(T.r2 + 1) * sizeof(double), 256);
    const size_t o_row = off;  off = align_up(off + (rows + 1) * sizeof(double), 256);
    const size_t o_rowL = off; off = align_up(off + (rows + 1) * sizeof(double), 256);
    const size_t o_val = off;  off += 256;
    if (aux && base) {
        aux->tilemax = (unsigned int*)(base + o_tile);
        aux->colsum = (double*)(base + o_col);
        aux->rowsum = (double*)(base + o_row);
        aux->rowsumL = (double*)(base + o_rowL);
        aux->valid = (int*)(base + o_val);
    }
    return (long long)off;
}

long long pq_grid_aux_bytes(const peuqse_problem* p, const DevSweep& S, long long first, long long count) {
    return grid_aux_layout(p, S, first, count, nullptr, nullptr);
}

template <int MODEL, int KT, int NT, bool EXACT>
static int launch_grid_pass1(const peuqse_problem* p, const DevSweep& S, long long first, long long count, double* logp,
                             unsigned int* tilemax, peuqse_best* best, void* work, cudaStream_t st) {
    const DevProblem& d = p->d;
    const bool simple = d.prior_mode == PEUQSE_PRIOR_DIAG && d.prior_identity && d.kp == d.k && KT < 8;
    unsigned int* ticket = (unsigned int*)work;
    Best* partial = (Best*)((char*)work + PQ_SWEEP_TICKET_BYTES);
    GridTiling T;
#define GP1_LAUNCH(SIMPLE, STORE)                                                                                            \
    if (bounds) GP1_LAUNCH2(SIMPLE, STORE, true) else GP1_LAUNCH2(SIMPLE, STORE, false)
#define GP1_LAUNCH2(SIMPLE, STORE, BOUNDS)                                                                                   \
    {                                                                                                                         \
        auto kern = k_grid_pass1<MODEL, KT, NT, EXACT, SIMPLE, STORE, BOUNDS>;                                                \
        const int resident = resident_blocks(p, kern, 0);                                                                    \
        grid_tiling(p, S, first, count, resident, &T);                                                                        \
        long long nblk = (long long)T.nchunks * T.nrg;                                                                        \
        if (T.dyn) {                        /* any grid works: fill the GPU exactly, but not more CTAs than units */          \
            const long long units = ((T.b_hi - T.b_lo + GR_RB - 1) / GR_RB) * T.nchunks;                                      \
            nblk = resident < PQ_SWEEP_MAX_BLOCKS ? resident : PQ_SWEEP_MAX_BLOCKS;                                            \
            if (nblk > units) nblk = units;                                                                                   \
            if (nblk < 1) nblk = 1;                                                                                           \
        }                                                                                                                     \
        kern<<<(int)nblk, GR_THREADS, 0, st>>>(p->d, S, T, first, count, logp, tilemax, best, partial, ticket);                \
    }
    const bool store = logp != nullptr && tilemax != nullptr;
    const bool bounds = d.has_lb || d.has_ub;
    if (simple && store) GP1_LAUNCH(true, true)
    else if (simple) GP1_LAUNCH(true, false)
    else if (store) GP1_LAUNCH(false, true)
    else GP1_LAUNCH(false, false)
#undef GP1_LAUNCH
#undef GP1_LAUNCH2
    PQ_CUDA_OK(cudaGetLastError());
    return PEUQSE_OK;
}

// *handled = true when the specialised kernel covered the whole range [first, first+count)
int pq_grid_pass1(const peuqse_problem* p, const DevSweep& S, long long first, long long count, double* logp, void* aux_base,
                  peuqse_best* best, void* work, cudaStream_t st, bool* handled) {
    *handled = false;
    GridAux aux = {nullptr, nullptr, nullptr, nullptr, nullptr};
    if (!fast_model_ok(p) || grid_aux_layout(p, S, first, count, (char*)aux_base, &aux) == 0) return PEUQSE_OK;
    if (logp && !aux_base) return PEUQSE_OK;           // storing logP without room for the tile maxima: generic path
    *handled = true;
    const int n = p->d.fast_n;
    if (p->d.model_id == PEUQSE_MODEL_QUADRATIC) {
        switch (n) {
#define GQ(N) case N: return launch_grid_pass1<PEUQSE_MODEL_QUADRATIC, 2, N, true>(p, S, first, count, logp, aux.tilemax, best, work, st);
#ifdef PQ_GRID_MINIMAL                      /* development builds: only the C2 instantiation */
            GQ(3)
#else
            GQ(1) GQ(2) GQ(3) GQ(4) GQ(5) GQ(6) GQ(7) GQ(8)
#endif
#undef GQ
        }
#ifndef PQ_GRID_MINIMAL
    } else {
#define GL(K)                                                                                                                       \
    case K:                                                                                                                         \
        return n <= 4 ? launch_grid_pass1<PEUQSE_MODEL_LINEAR, K, 4, false>(p, S, first, count, logp, aux.tilemax, best, work, st)    \
                      : launch_grid_pass1<PEUQSE_MODEL_LINEAR, K, 8, false>(p, S, first, count, logp, aux.tilemax, best, work, st);
        switch (p->d.k) { GL(2) GL(3) GL(4) GL(6) }
#undef GL
#endif
    }
    return pq_set_error(PEUQSE_ERR_INVALID, "grid fast path dispatch");
}

// whether pass 2 of this range runs fused (leaves row/column weight sums for the binning kernel)
bool pq_grid_pass2_fuses(const peuqse_problem* p, const DevSweep& S, long long first, long long count) {
    GridTiling T;
    if (grid_aux_layout(p, S, first, count, nullptr, nullptr) == 0) return false;
    grid_tiling(p, S, first, count, 1, &T);
    return (size_t)T.nchunks * GR_CH * sizeof(double) + 1024 <= (size_t)p->max_smem_optin;
}

int pq_grid_pass2(const peuqse_problem* p, const DevSweep& S, long long first, long long count, const double* logp,
                  void* aux_base, const peuqse_best* best, double log_rel, double* acc, long long* kidx, double* klogp,
                  long long kcap, long long* kcount, void* work, cudaStream_t st, bool* handled) {
    *handled = false;
    GridAux aux = {nullptr, nullptr, nullptr, nullptr, nullptr};
    GridTiling T;
    if (!logp || !aux_base || grid_aux_layout(p, S, first, count, (char*)aux_base, &aux) == 0) return PEUQSE_OK;
    grid_tiling(p, S, first, count, 1, &T);
    *handled = true;
    unsigned int* ticket = (unsigned int*)work;
    double* partial = (double*)((char*)work + PQ_SWEEP_TICKET_BYTES);
    const size_t smem = (size_t)T.nchunks * GR_CH * sizeof(double);
    const bool fuse = pq_grid_pass2_fuses(p, S, first, count);
    const size_t smem_tma = (size_t)P2_STAGES * P2_STAGE_DOUBLES * sizeof(double) + P2_KLIST * sizeof(int) +
                            (fuse ? (size_t)T.nchunks * GR_CH * sizeof(unsigned int) : 0);
    const bool use_tma = (((uintptr_t)logp) & 15) == 0 && kidx == nullptr && smem_tma + 1024 <= (size_t)p->max_smem_optin;
#define G2T(K, F)                                                                                                                \
    {                                                                                                                            \
        cudaError_t e = pq_allow_smem((const void*)k_grid_pass2_tma<K, F>, smem_tma);                                            \
        if (e != cudaSuccess) return pq_set_cuda_error(e, "cudaFuncSetAttribute", __LINE__);                                      \
        const int g = flat_grid(T, resident_blocks(p, k_grid_pass2_tma<K, F>, smem_tma));                                        \
        k_grid_pass2_tma<K, F><<<g, GR_THREADS, smem_tma, st>>>(S, T, first, count, logp, aux, best, log_rel, partial, acc,       \
                                                                 ticket);                                                         \
    }
#define G2(K)                                                                                                                    \
    case K: {                                                                                                                    \
        if (fuse) {                                                                                                              \
            cudaError_t e = cudaMemsetAsync(aux.colsum, 0, (size_t)((char*)aux.valid - (char*)aux.colsum), st);                   \
            if (e != cudaSuccess) return pq_set_cuda_error(e, "cudaMemsetAsync", __LINE__);                                       \
        }                                                                                                                        \
        if (use_tma && fuse) {                                                                                                   \
            const size_t sm2 = (size_t)GR_WARPS * P2W_STAGES * P2W_STAGE_DOUBLES * sizeof(double) +                               \
                               (size_t)T.nchunks * GR_CH * sizeof(unsigned int);                                                   \
            cudaError_t e = pq_allow_smem((const void*)k_grid_pass2_w<K>, sm2);                                                   \
            if (e != cudaSuccess) return pq_set_cuda_error(e, "cudaFuncSetAttribute", __LINE__);                                  \
            const int g = resident_blocks(p, k_grid_pass2_w<K>, sm2);                                                             \
            k_grid_pass2_w<K><<<g, GR_THREADS, sm2, st>>>(S, T, first, count, logp, aux, best, log_rel, partial, acc, ticket);     \
            k_grid_finish<K><<<16, GR_THREADS, 0, st>>>(S, T, first, logp, aux, best, log_rel, partial, acc, ticket);             \
        } else if (use_tma) {                                                                                                    \
            G2T(K, false)                                                                                                        \
        } else if (fuse) {                                                                                                       \
            cudaError_t e = pq_allow_smem((const void*)k_grid_pass2<K, true>, smem);                                             \
            if (e != cudaSuccess) return pq_set_cuda_error(e, "cudaFuncSetAttribute", __LINE__);                                  \
            const int g = flat_grid(T, resident_blocks(p, k_grid_pass2<K, true>, smem));                                         \
            k_grid_pass2<K, true><<<g, GR_THREADS, smem, st>>>(S, T, first, count, logp, aux, best, log_rel, partial, acc, ticket, \
                                                               kidx, klogp, kcap, (unsigned long long*)kcount);                   \
            e = cudaGetLastError();                                                                                              \
            if (e != cudaSuccess) return pq_set_cuda_error(e, "k_grid_pass2<fused> launch", (int)smem);                          \
        } else {                                                                                                                 \
            const int g = flat_grid(T, resident_blocks(p, k_grid_pass2<K, false>, 0));                                           \
            k_grid_pass2<K, false><<<g, GR_THREADS, 0, st>>>(S, T, first, count, logp, aux, best, log_rel, partial, acc, ticket,   \
                                                             kidx, klogp, kcap, (unsigned long long*)kcount);                     \
        }                                                                                                                        \
    } break;
    switch (S.k) { G2(1) G2(2) G2(3) G2(4) G2(6) }
#undef G2T
#undef G2
    PQ_CUDA_OK(cudaGetLastError());
    return PEUQSE_OK;
}

// use_fused_sums: the caller ran pass 2 of the same range immediately before (row / column sums are in d_work)
int pq_grid_hist(const peuqse_problem* p, const DevSweep& S, long long first, long long count, const double* logp,
                 void* aux_base, const peuqse_best* best, double log_rel, const double* edges, int nbins,
                 unsigned long long* hist, cudaStream_t st, bool* handled, bool use_fused_sums) {
    *handled = false;
    GridAux aux = {nullptr, nullptr, nullptr, nullptr, nullptr};
    GridTiling T;
    if (!logp || !aux_base || grid_aux_layout(p, S, first, count, (char*)aux_base, &aux) == 0) return PEUQSE_OK;
    grid_tiling(p, S, first, count, 1, &T);
    const size_t smem_bins = (size_t)S.k * nbins * sizeof(unsigned long long);
    if (use_fused_sums && pq_grid_pass2_fuses(p, S, first, count)) {
        *handled = true;
#define GB(K) case K: k_grid_bin<K><<<40, GR_THREADS, smem_bins, st>>>(S, T, first, logp, aux, best, log_rel, edges, nbins, hist); break;
        switch (S.k) { GB(1) GB(2) GB(3) GB(4) GB(6) }
#undef GB
        PQ_CUDA_OK(cudaGetLastError());
        return PEUQSE_OK;
    }
    const size_t smem = smem_bins + (size_t)T.nchunks * GR_CH * sizeof(double);
    if (smem > (size_t)p->max_smem_optin) return PEUQSE_OK;                           // last axis too long: generic kernel
    *handled = true;
#define G3(K)                                                                                                                    \
    case K: {                                                                                                                    \
        cudaError_t e = pq_allow_smem((const void*)k_grid_hist<K>, smem);                                                        \
        if (e != cudaSuccess) return pq_set_cuda_error(e, "cudaFuncSetAttribute", __LINE__);                                      \
        const int g = flat_grid(T, resident_blocks(p, k_grid_hist<K>, smem));                                                     \
        k_grid_hist<K><<<g, GR_THREADS, smem, st>>>(S, T, first, count, logp, aux, best, log_rel, edges, nbins, hist);             \
    } break;
    switch (S.k) { G3(1) G3(2) G3(3) G3(4) G3(6) }
#undef G3
    PQ_CUDA_OK(cudaGetLastError());
    return PEUQSE_OK;
}
