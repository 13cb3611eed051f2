// Post-burn-in statistics of device-resident samples: calculatePostBurnInStatistics
// (IP.py:1861-1937) with its tail filter (arrayThresholdFilter, IP.py:3635-3664) and calculateInfoGain (IP.py:1770-1847),
// as streaming reductions over samples[N][k], logP[N], logPrior[N] that already live in HBM (the sampler kernels'
// outputs), so that neither the 8(k+2) N bytes nor a numpy pass over them ever cross to the host.
//
//   begin : A   sum and sum of squares of log(-logP) about a pivot (the first row's value), max logP (+ NaN flag),
//               sum theta                                                   (the unfiltered quantities of IP.py:1894-1900)
//   finish: CUT threshold of the chosen benchmark -> one rule for every mode: a row is REMOVED iff logP < cut
//           B1  over kept rows: n, sum theta, sum theta^2, first maximum of logP, sum exp(logP), min / max theta
//           MID mu_AP, evidence (IP.py:1931), joint histogram range of prior and posterior samples
//           B2  sum (theta - mu_AP)^2, log-ratio information gain terms, rank of the MAP row among the kept rows,
//               100-bin histograms (np.histogram semantics) of the kept samples; the prior samples likewise
//           FIN stdap, info gains, packed result
//   compact (optional): order-preserving copy of the kept rows.
// All sums are deterministic: every CTA owns a contiguous row range, per-CTA partials are combined in CTA order by the
// last CTA to finish.  HBM-bound: 8 B/row (begin) + 2 x 8(k+2) B/row (finish) [+ 16(k+2) B/row for compact].
// B1 / B2 stream contiguous 256-row tiles (samples, logP, logPrior) through a 3-stage cp.async.bulk (TMA) ring into
// shared memory: one elected thread issues the copies, every thread then reads its own row from the staged tile.
//
// 'auto' benchmark: the reference removes rows with log(-logP) > mean + coef*std.  log is monotone, so that set is
// {-logP > x0} with x0 the largest double whose log is <= the threshold; CUT finds x0 by stepping ulps around
// exp(threshold) and the passes compare logP itself (no log per row and pass).  NaN thresholds remove nothing and rows
// with logP >= 0 or NaN are kept, exactly like `filterKey > thresholdValue` on NaN keys (IP.py:3650).
#include "common.cuh"
#include "eval.cuh"
#include "launch.cuh"
#include "tma.cuh"

#define ST_THREADS 256
#define ST_MAXG 1024
#define ST_K PEUQSE_MAX_K
#define ST_SLOTS (8 + 4 * ST_K)            // per-CTA partial vector
#define ST_MAXBINS 128

struct StatsWork {                          // layout of d_work
    unsigned int ticket[64];
    double res[ST_SLOTS];                   // combined partials of the current pass
    double bestv;
    long long besti;
    double keep[16];                        // values carried between passes (see K_*)
    double partial[ST_MAXG][ST_SLOTS];
    double pbv[ST_MAXG];
    long long pbi[ST_MAXG];
    long long counts[ST_MAXG];              // kept rows per CTA (B1), for compact
    unsigned long long hist_post[ST_K * ST_MAXBINS];
    unsigned long long hist_prior[ST_K * ST_MAXBINS];
    double edges[ST_K * (ST_MAXBINS + 1)];
};
enum { K_N = 3, K_LOGEV = 4 };
#define ST_STAGES 3
#define ST_TILE ST_THREADS                  // rows per staged tile

struct StatsGeom {
    long long n, rows_per_cta;
    int grid, k;
};

static StatsGeom stats_geom(long long n, int k, int sm_count) {
    StatsGeom g;
    g.n = n;
    g.k = k;
    long long want = (n + ST_THREADS - 1) / ST_THREADS;
    long long grid = (long long)sm_count * 4;
    if (grid > ST_MAXG) grid = ST_MAXG;
    if (grid > want) grid = want;
    if (grid < 1) grid = 1;
    long long rpc = (n + grid - 1) / grid;
    rpc = (rpc + ST_THREADS - 1) / ST_THREADS * ST_THREADS;
    if (rpc < ST_THREADS) rpc = ST_THREADS;
    g.rows_per_cta = rpc;
    g.grid = (int)((n + rpc - 1) / rpc);
    if (g.grid < 1) g.grid = 1;
    return g;
}

template <int KT, bool VEC2>
__device__ __forceinline__ void load_row(const double* __restrict__ s, long long i, int k, double (&th)[KT]) {
    if (VEC2) {
        const double2* r = reinterpret_cast<const double2*>(s + i * k);
#pragma unroll
        for (int d = 0; d < KT; d += 2)
            if (d < k) {
                const double2 v = __ldcs(r + (d >> 1));
                th[d] = v.x;
                if (d + 1 < KT) th[d + 1] = v.y;
            }
    } else {
        const double* r = s + i * k;
#pragma unroll
        for (int d = 0; d < KT; ++d) th[d] = (d < k) ? __ldcs(r + d) : 0.0;
    }
}

// Combine the CTA's values v[0..nv) (slots < nsum: sum, < nsum+nmin: min, else max) and an arg-max record into the
// per-CTA partials; the last CTA to arrive reduces them in CTA order into W->res / W->bestv,besti.  Returns true in
// the last CTA (all threads) after the combined values are visible.
template <int NV>
__device__ __forceinline__ bool cta_combine(StatsWork* W, double (&v)[NV], int nsum, int nmin, Best best, bool use_best) {
    __shared__ double s_v[ST_THREADS / 32][NV];
    __shared__ Best s_b[ST_THREADS / 32];
    __shared__ bool s_last;
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
#pragma unroll
    for (int a = 0; a < NV; ++a) {
        double x = v[a];
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) {
            const double y = __shfl_xor_sync(0xffffffffu, x, o);
            x = (a < nsum) ? (x + y) : ((a < nsum + nmin) ? fmin(x, y) : fmax(x, y));
        }
        if (lane == 0) s_v[warp][a] = x;
    }
    if (use_best) {
        best = warp_best(best);
        if (lane == 0) s_b[warp] = best;
    }
    __syncthreads();
    if (threadIdx.x < NV) {
        const int a = threadIdx.x;
        double x = s_v[0][a];
        for (int w = 1; w < ST_THREADS / 32; ++w) {
            const double y = s_v[w][a];
            x = (a < nsum) ? (x + y) : ((a < nsum + nmin) ? fmin(x, y) : fmax(x, y));
        }
        W->partial[blockIdx.x][a] = x;
    }
    if (use_best && threadIdx.x == 0) {
        Best b = s_b[0];
        for (int w = 1; w < ST_THREADS / 32; ++w) b = best_merge(b, s_b[w]);
        W->pbv[blockIdx.x] = b.v;
        W->pbi[blockIdx.x] = b.i;
    }
    __threadfence();
    __syncthreads();
    if (threadIdx.x == 0) s_last = (atomicAdd(&W->ticket[0], 1u) == gridDim.x - 1);
    __syncthreads();
    if (!s_last) return false;
    __threadfence();
    if (threadIdx.x < NV) {
        const int a = threadIdx.x;
        double x = __ldcg(&W->partial[0][a]);
        for (int b = 1; b < (int)gridDim.x; ++b) {
            const double y = __ldcg(&W->partial[b][a]);
            x = (a < nsum) ? (x + y) : ((a < nsum + nmin) ? fmin(x, y) : fmax(x, y));
        }
        W->res[a] = x;
    }
    if (use_best && threadIdx.x == 32) {
        Best b = best_none();
        for (int c = 0; c < (int)gridDim.x; ++c) b = best_merge(b, Best{__ldcg(&W->pbv[c]), __ldcg(&W->pbi[c])});
        W->bestv = b.v;
        W->besti = b.i;
    }
    if (threadIdx.x == 0) W->ticket[0] = 0u;
    __syncthreads();
    return true;
}

// ---------------------------------------------------------------------------------------------
// begin
// ---------------------------------------------------------------------------------------------
template <int KT, bool VEC2, bool THETA>
__global__ void __launch_bounds__(ST_THREADS) k_stats_a1(const double* __restrict__ samples, const double* __restrict__ logp,
                                                         StatsGeom G, StatsWork* W, double* __restrict__ out) {
    // slots: [0] sum (x - p), [1] sum (x - p)^2 with x = log(-logP) and p = x of row 0, [2] NaN logP count, [3..] sum theta
    double v[3 + KT];
#pragma unroll
    for (int a = 0; a < 3 + KT; ++a) v[a] = 0.0;
    Best best = best_none();
    const double pivot = log(-1.0 * logp[0]);
    const bool piv_ok = pivot == pivot && fabs(pivot) < pq_pos_inf();
    const long long r0 = (long long)blockIdx.x * G.rows_per_cta;
    const long long r1 = (r0 + G.rows_per_cta < G.n) ? r0 + G.rows_per_cta : G.n;
    for (long long i = r0 + threadIdx.x; i < r1; i += ST_THREADS) {
        const double lp = __ldcs(logp + i);
        const double e = log(-1.0 * lp) - (piv_ok ? pivot : 0.0);    // IP.py:1894
        v[0] += e;
        v[1] += e * e;
        if (lp != lp) v[2] += 1.0;
        if (lp > best.v) { best.v = lp; best.i = i; }
        if (THETA) {
            double th[KT];
            load_row<KT, VEC2>(samples, i, G.k, th);
#pragma unroll
            for (int d = 0; d < KT; ++d) v[3 + d] += th[d];
        }
    }
    if (cta_combine<3 + KT>(W, v, 3 + KT, 0, best, true)) {
        if (threadIdx.x == 0) {
            const double n = (double)G.n;
            const double m1 = W->res[0] / n;                           // mean - pivot
            double var = W->res[1] / n - m1 * m1;                      // mean |x - mean|^2 (np.std, IP.py:1896)
            if (var < 0.0) var = 0.0;
            out[PEUQSE_STATS_N_TOTAL] = n;
            out[PEUQSE_STATS_MEAN_LOGNEG] = (piv_ok ? pivot : 0.0) + m1;   // IP.py:1895
            out[PEUQSE_STATS_STD_LOGNEG] = sqrt(var);
            out[PEUQSE_STATS_MAX_LOGP_ALL] = (W->res[2] > 0.0) ? __longlong_as_double(0x7ff8000000000000LL) : W->bestv;   // np.max propagates NaN
        }
        if (THETA && threadIdx.x < G.k) out[PEUQSE_STATS_MEAN_ALL + threadIdx.x] = W->res[3 + threadIdx.x] / (double)G.n;   // IP.py:1899
    }
}

// ---------------------------------------------------------------------------------------------
// tile streaming for B1 / B2: stage buffer = [ST_TILE*k samples][ST_TILE logP][ST_TILE logPrior]
// ---------------------------------------------------------------------------------------------
struct TileRing {
    double* buf;                 // [ST_STAGES][stage_doubles]
    unsigned long long* full;    // [ST_STAGES]
    int stage_doubles, k;
    bool prior;
    __device__ __forceinline__ void issue(int st, const double* samples, const double* logp, const double* logprior, long long row0) {
        double* b = buf + (size_t)st * stage_doubles;
        const uint32_t bs = (uint32_t)(ST_TILE * k * sizeof(double)), bl = (uint32_t)(ST_TILE * sizeof(double));
        mbar_arrive_expect_tx(&full[st], bs + bl + (prior ? bl : 0u));
        tma_load_1d(b, samples + row0 * k, bs, &full[st]);
        tma_load_1d(b + ST_TILE * k, logp + row0, bl, &full[st]);
        if (prior) tma_load_1d(b + ST_TILE * k + ST_TILE, logprior + row0, bl, &full[st]);
    }
};

// Calls f(row, logP, logPrior, theta[KT]) for every row of this CTA's range: full 256-row tiles through the TMA ring
// (when `tma`), the ragged tail (and everything, without `tma`) with direct loads.
template <int KT, bool VEC2, class F>
__device__ __forceinline__ void for_each_row(const double* __restrict__ samples, const double* __restrict__ logp,
                                             const double* __restrict__ logprior, const StatsGeom& G, bool tma, unsigned char* s_raw,
                                             F f) {
    const long long r0 = (long long)blockIdx.x * G.rows_per_cta;
    const long long r1 = (r0 + G.rows_per_cta < G.n) ? r0 + G.rows_per_cta : G.n;
    long long done = r0;
    if (tma) {
        __shared__ __align__(8) unsigned long long s_full[ST_STAGES];
        TileRing ring;
        ring.buf = (double*)s_raw;
        ring.full = s_full;
        ring.k = G.k;
        ring.prior = logprior != nullptr;
        ring.stage_doubles = ST_TILE * (G.k + 2);
        const long long ntile = (r1 - r0) / ST_TILE;
        if (threadIdx.x == 0) {
            for (int q = 0; q < ST_STAGES; ++q) mbar_init(&s_full[q], 1);
            asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
        }
        __syncthreads();
        if (threadIdx.x == 0)
            for (int q = 0; q < ST_STAGES - 1 && q < ntile; ++q) ring.issue(q, samples, logp, logprior, r0 + (long long)q * ST_TILE);
        for (long long t = 0; t < ntile; ++t) {
            const int st = (int)(t % ST_STAGES);
            if (threadIdx.x == 0 && t + ST_STAGES - 1 < ntile)       // the stage consumed in the previous iteration
                ring.issue((int)((t + ST_STAGES - 1) % ST_STAGES), samples, logp, logprior, r0 + (t + ST_STAGES - 1) * ST_TILE);
            while (!mbar_try_wait(&s_full[st], (uint32_t)((t / ST_STAGES) & 1))) {}
            const double* b = ring.buf + (size_t)st * ring.stage_doubles;
            double th[KT];
#pragma unroll
            for (int d = 0; d < KT; ++d) th[d] = (d < G.k) ? b[threadIdx.x * G.k + d] : 0.0;
            const double lp = b[ST_TILE * G.k + threadIdx.x];
            const double lpr = ring.prior ? b[ST_TILE * G.k + ST_TILE + threadIdx.x] : 0.0;
            f(r0 + t * ST_TILE + threadIdx.x, lp, lpr, th);
            __syncthreads();                                       // stage st may be refilled from now on
        }
        done = r0 + ntile * ST_TILE;
    }
    for (long long i = done + threadIdx.x; i < r1; i += ST_THREADS) {
        double th[KT];
        load_row<KT, VEC2>(samples, i, G.k, th);
        f(i, __ldcs(logp + i), logprior ? __ldcs(logprior + i) : 0.0, th);
    }
}

// ---------------------------------------------------------------------------------------------
// finish
// ---------------------------------------------------------------------------------------------
__global__ void k_stats_cut(int filter, double coef, const double* __restrict__ mu_ap_logp, double* __restrict__ out,
                            StatsWork* W, int k, int bins) {
    if (threadIdx.x != 0) return;
    double thr = __longlong_as_double(0x7ff8000000000000LL), cut = pq_neg_inf();
    if (filter == PEUQSE_FILTER_AUTO) {
        thr = out[PEUQSE_STATS_MEAN_LOGNEG] + coef * out[PEUQSE_STATS_STD_LOGNEG];                 // IP.py:1915
        if (thr == thr && thr < pq_pos_inf()) {
            // x0 = largest double with log(x0) <= thr; removed rows are those with -logP > x0, i.e. logP < -x0
            double x = exp(thr);
            if (!(x < pq_pos_inf())) x = 1.7976931348623157e308;
            for (int it = 0; it < 64 && log(x) > thr; ++it) x = __longlong_as_double(__double_as_longlong(x) - 1);
            for (int it = 0; it < 64; ++it) {
                const double xn = __longlong_as_double(__double_as_longlong(x) + 1);
                if (!(xn < pq_pos_inf()) || log(xn) > thr) break;
                x = xn;
            }
            cut = -x;
        }
    } else if (filter == PEUQSE_FILTER_MAP) {
        thr = out[PEUQSE_STATS_MAX_LOGP_ALL] - coef;                                               // IP.py:1910
        cut = (thr == thr) ? thr : pq_neg_inf();
    } else if (filter == PEUQSE_FILTER_MU_AP) {
        thr = mu_ap_logp[0] - coef;                                                                // IP.py:1904
        cut = (thr == thr) ? thr : pq_neg_inf();
    }
    out[PEUQSE_STATS_THRESHOLD] = thr;
    out[PEUQSE_STATS_CUT] = cut;
    for (int a = 0; a < k * bins; ++a) { W->hist_post[a] = 0ULL; W->hist_prior[a] = 0ULL; }
}

// RANGE: also the per-parameter minimum / maximum of the kept rows (needed for the KL histograms' range only)
template <int KT, bool VEC2, bool RANGE>
__global__ void __launch_bounds__(ST_THREADS) k_stats_b1(const double* __restrict__ samples, const double* __restrict__ logp,
                                                         StatsGeom G, StatsWork* W, double* __restrict__ out, bool tma) {
    // slots: [0] n_kept, [1] sum exp(logP), [2..2+KT) sum theta, [2+KT..2+2KT) sum theta^2 | min[KT] | max[KT]
    extern __shared__ __align__(16) unsigned char s_raw[];
    constexpr int NV = 2 + 4 * KT;
    double v[NV];
#pragma unroll
    for (int a = 0; a < 2 + 2 * KT; ++a) v[a] = 0.0;
#pragma unroll
    for (int d = 0; d < KT; ++d) { v[2 + 2 * KT + d] = pq_pos_inf(); v[2 + 3 * KT + d] = pq_neg_inf(); }
    Best best = best_none();
    const double cut = out[PEUQSE_STATS_CUT];
    for_each_row<KT, VEC2>(samples, logp, nullptr, G, tma, s_raw, [&](long long i, double lp, double, const double (&th)[KT]) {
        if (!(lp < cut)) {
            v[0] += 1.0;
            v[1] += exp(lp);                                       // IP.py:1931
            if (lp > best.v) { best.v = lp; best.i = i; }          // first maximum in list order (IP.py:1925-1927)
#pragma unroll
            for (int d = 0; d < KT; ++d) {
                v[2 + d] += th[d];
                v[2 + KT + d] += th[d] * th[d];
                if (RANGE) {
                    v[2 + 2 * KT + d] = fmin(v[2 + 2 * KT + d], th[d]);
                    v[2 + 3 * KT + d] = fmax(v[2 + 3 * KT + d], th[d]);
                }
            }
        }
    });
    {   // kept rows of this CTA (for compact): block-level count
        __shared__ double s_cnt[ST_THREADS / 32];
        const double c = warp_sum(v[0]);
        if ((threadIdx.x & 31) == 0) s_cnt[threadIdx.x >> 5] = c;
        __syncthreads();
        if (threadIdx.x == 0) {
            double t = 0.0;
            for (int w = 0; w < ST_THREADS / 32; ++w) t += s_cnt[w];
            W->counts[blockIdx.x] = (long long)t;
        }
    }
    if (cta_combine<NV>(W, v, 2 + 2 * KT, KT, best, true)) {
        const double n = W->res[0];
        if (threadIdx.x < G.k) {
            const int d = threadIdx.x;
            out[PEUQSE_STATS_MU_AP + d] = W->res[2 + d] / n;                                         // IP.py:1923
            const double qnan = __longlong_as_double(0x7ff8000000000000LL);
            out[PEUQSE_STATS_MIN + d] = RANGE ? W->res[2 + 2 * KT + d] : qnan;
            out[PEUQSE_STATS_MAX + d] = RANGE ? W->res[2 + 3 * KT + d] : qnan;
        }
        if (threadIdx.x == 0) {
            double f2 = 0.0;
            for (int d = 0; d < G.k; ++d) f2 += W->res[2 + KT + d];
            const double ev = (W->res[1] / n) / sqrt(f2);                                            // IP.py:1931
            out[PEUQSE_STATS_N_KEPT] = n;
            out[PEUQSE_STATS_SUM_EXP] = W->res[1];
            out[PEUQSE_STATS_FROB2] = f2;
            out[PEUQSE_STATS_EVIDENCE] = ev;
            out[PEUQSE_STATS_MAP_LOGP] = W->bestv;
            const long long bi = (W->besti == 0x7fffffffffffffffLL) ? -1 : W->besti;
            out[PEUQSE_STATS_MAP_INDEX_ALL] = __longlong_as_double(bi);
            W->keep[K_N] = n;
            W->keep[K_LOGEV] = log(ev);                                                              // IP.py:1799
        }
    }
}

// min / max of the prior samples per parameter, then the joint edges np.histogram uses for
// plt.hist([prior, posterior], bins) (one common range over both data sets)
__global__ void __launch_bounds__(ST_THREADS) k_stats_edges(const double* __restrict__ prior, long long np, int k, int bins,
                                                           double* __restrict__ out, StatsWork* W) {
    __shared__ double s_lo[ST_THREADS / 32], s_hi[ST_THREADS / 32];
    const int d = blockIdx.x;
    double lo = pq_pos_inf(), hi = pq_neg_inf();
    for (long long i = threadIdx.x; i < np; i += ST_THREADS) {
        const double x = prior[i * k + d];
        lo = fmin(lo, x);
        hi = fmax(hi, x);
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
        lo = fmin(lo, __shfl_xor_sync(0xffffffffu, lo, o));
        hi = fmax(hi, __shfl_xor_sync(0xffffffffu, hi, o));
    }
    if ((threadIdx.x & 31) == 0) { s_lo[threadIdx.x >> 5] = lo; s_hi[threadIdx.x >> 5] = hi; }
    __syncthreads();
    if (threadIdx.x == 0) {
        for (int w = 1; w < ST_THREADS / 32; ++w) { lo = fmin(lo, s_lo[w]); hi = fmax(hi, s_hi[w]); }
        lo = fmin(s_lo[0], lo); hi = fmax(s_hi[0], hi);
        lo = fmin(lo, out[PEUQSE_STATS_MIN + d]);
        hi = fmax(hi, out[PEUQSE_STATS_MAX + d]);
        if (!(hi > lo)) { lo = lo - 0.5; hi = hi + 0.5; }          // np.histogram widens an empty range by +-0.5
        s_lo[0] = lo; s_hi[0] = hi;
        out[PEUQSE_STATS_JOINT_LO + d] = lo;
        out[PEUQSE_STATS_JOINT_HI + d] = hi;
    }
    __syncthreads();
    lo = s_lo[0]; hi = s_hi[0];
    const double step = (hi - lo) / (double)bins;                  // np.linspace(lo, hi, bins + 1)
    for (int i = threadIdx.x; i <= bins; i += ST_THREADS) W->edges[d * (bins + 1) + i] = (i == bins) ? hi : ((double)i * step + lo);
}

// np.histogram bin of x (numpy/lib/_histograms_impl.py), -1 outside [lo, hi]
__device__ __forceinline__ int st_hist_bin(const double* __restrict__ ed, int nbins, double x) {
    const double lo = ed[0], hi = ed[nbins];
    if (!(x >= lo && x <= hi)) return -1;
    int b = (int)(((x - lo) / (hi - lo)) * (double)nbins);
    if (b >= nbins) b = nbins - 1;
    if (b < 0) b = 0;
    if (x < ed[b]) --b;
    else if (b != nbins - 1 && x >= ed[b + 1]) ++b;
    return b;
}

template <int KT, bool VEC2>
__global__ void __launch_bounds__(ST_THREADS) k_stats_b2(const double* __restrict__ samples, const double* __restrict__ logp,
                                                         const double* __restrict__ logprior, StatsGeom G, StatsWork* W,
                                                         double* __restrict__ out, int bins, int copies, bool tma) {
    // slots: [0] sum log-ratio terms, [1] kept rows before the MAP row, [2..2+KT) sum (theta - mu_AP)^2
    extern __shared__ __align__(16) unsigned char s_raw[];
    const size_t ring_bytes = tma ? (size_t)ST_STAGES * ST_TILE * (G.k + 2) * sizeof(double) : 0;
    double* s_edges = (double*)(s_raw + ring_bytes);                              // [k][bins+1]
    unsigned int* s_hist = (unsigned int*)(s_edges + (size_t)G.k * (bins + 1));   // [k][bins][copies]: lane & (copies-1) picks a
    const int cmask = copies - 1, csel = (int)(threadIdx.x & 31) & cmask;         // private copy: fewer same-address conflicts
    if (bins > 0) {
        for (int a = threadIdx.x; a < G.k * (bins + 1); a += ST_THREADS) s_edges[a] = W->edges[a];
        for (int a = threadIdx.x; a < G.k * bins * copies; a += ST_THREADS) s_hist[a] = 0u;
        __syncthreads();
    }
    constexpr int NV = 2 + KT;
    double v[NV];
#pragma unroll
    for (int a = 0; a < NV; ++a) v[a] = 0.0;
    double mu[KT];
#pragma unroll
    for (int d = 0; d < KT; ++d) mu[d] = (d < G.k) ? out[PEUQSE_STATS_MU_AP + d] : 0.0;
    const double cut = out[PEUQSE_STATS_CUT], logev = W->keep[K_LOGEV];
    const long long istar = __double_as_longlong(out[PEUQSE_STATS_MAP_INDEX_ALL]);
    const bool has_prior = logprior != nullptr;
    for_each_row<KT, VEC2>(samples, logp, logprior, G, tma, s_raw, [&](long long i, double lp, double lpr, const double (&th)[KT]) {
        if (!(lp < cut)) {
            if (has_prior) {
                const double t = (lp - logev) - lpr;               // IP.py:1799-1800
                if (t == t && fabs(t) < pq_pos_inf()) v[0] += t;   // inf -> 0, NaN -> 0 (IP.py:1801-1802)
            }
            if (i < istar) v[1] += 1.0;
#pragma unroll
            for (int d = 0; d < KT; ++d) {
                const double e = th[d] - mu[d];
                v[2 + d] += e * e;                                 // np.std: mean(|x - mean|^2)
                if (bins > 0 && d < G.k) {
                    const int b = st_hist_bin(s_edges + d * (bins + 1), bins, th[d]);
                    if (b >= 0) atomicAdd(&s_hist[(d * bins + b) * copies + csel], 1u);
                }
            }
        }
    });
    if (bins > 0) {
        __syncthreads();
        for (int a = threadIdx.x; a < G.k * bins; a += ST_THREADS) {
            unsigned int c = 0u;
            for (int q = 0; q < copies; ++q) c += s_hist[a * copies + q];
            if (c) atomicAdd(&W->hist_post[a], (unsigned long long)c);
        }
    }
    if (cta_combine<NV>(W, v, NV, 0, best_none(), false)) {
        const double n = W->keep[K_N];
        if (threadIdx.x < G.k) out[PEUQSE_STATS_STDAP + threadIdx.x] = sqrt(W->res[2 + threadIdx.x] / n);   // IP.py:1924
        if (threadIdx.x == 0) {
            out[PEUQSE_STATS_INFO_GAIN_LOG_RATIO] = logprior ? W->res[0] / n : __longlong_as_double(0x7ff8000000000000LL);
            const long long bi = __double_as_longlong(out[PEUQSE_STATS_MAP_INDEX_ALL]);
            out[PEUQSE_STATS_MAP_INDEX] = __longlong_as_double(bi < 0 ? -1LL : (long long)W->res[1]);        // IP.py:1927
        }
    }
}

__global__ void __launch_bounds__(ST_THREADS) k_stats_prior_hist(const double* __restrict__ prior, long long np, int k, int bins,
                                                                StatsWork* W) {
    extern __shared__ __align__(16) unsigned char s_raw[];
    unsigned int* s_hist = (unsigned int*)s_raw;                    // [k][bins], flushed once per CTA
    for (int a = threadIdx.x; a < k * bins; a += ST_THREADS) s_hist[a] = 0u;
    __syncthreads();
    const long long total = np * k;
    for (long long e = (long long)blockIdx.x * ST_THREADS + threadIdx.x; e < total; e += (long long)gridDim.x * ST_THREADS) {
        const int d = (int)(e % k);
        const int b = st_hist_bin(W->edges + d * (bins + 1), bins, prior[e]);
        if (b >= 0) atomicAdd(&s_hist[d * bins + b], 1u);
    }
    __syncthreads();
    for (int a = threadIdx.x; a < k * bins; a += ST_THREADS)
        if (s_hist[a]) atomicAdd(&W->hist_prior[a], (unsigned long long)s_hist[a]);
}

// MAP row and KL-divergence information gain (IP.py:1772-1791): densities of np.histogram(density=True)
__global__ void __launch_bounds__(ST_THREADS) k_stats_final(const double* __restrict__ samples, int k, int bins, bool kl,
                                                           double* __restrict__ out, StatsWork* W) {
    const long long bi = __double_as_longlong(out[PEUQSE_STATS_MAP_INDEX_ALL]);
    if ((int)threadIdx.x < k)
        out[PEUQSE_STATS_MAP_THETA + threadIdx.x] = bi >= 0 ? samples[bi * k + threadIdx.x] : __longlong_as_double(0x7ff8000000000000LL);
    if (!kl) {
        if (threadIdx.x == 0) out[PEUQSE_STATS_INFO_GAIN_KL] = __longlong_as_double(0x7ff8000000000000LL);
        return;
    }
    __shared__ double s_each[ST_K];
    if ((int)threadIdx.x < k) {
        const int d = threadIdx.x;
        const double* ed = W->edges + d * (bins + 1);
        double n0 = 0.0, n1 = 0.0;
        for (int b = 0; b < bins; ++b) { n0 += (double)W->hist_prior[d * bins + b]; n1 += (double)W->hist_post[d * bins + b]; }
        double acc = 0.0;
        for (int b = 0; b < bins; ++b) {                           // sequential like np.sum over <= 128 finite terms (pairwise kicks in above 8:
            const double db = ed[b + 1] - ed[b];                   // the difference is far below the histogram's own resolution)
            const double d0 = (double)W->hist_prior[d * bins + b] / db / n0;   // counts / db / counts.sum()
            const double d1 = (double)W->hist_post[d * bins + b] / db / n1;
            const double t = d1 * log(d1 / d0);
            if (t == t && fabs(t) < pq_pos_inf()) acc += t;        // np.isfinite filter (IP.py:1786)
        }
        s_each[d] = acc;
        out[PEUQSE_STATS_KL_EACH + d] = acc;
    }
    __syncthreads();
    if (threadIdx.x == 0) {
        double t = 0.0;
        for (int d = 0; d < k; ++d) t = t + s_each[d];             // IP.py:1789
        out[PEUQSE_STATS_INFO_GAIN_KL] = t;
    }
}

// ---------------------------------------------------------------------------------------------
// compact: kept rows copied in order (CTA offsets = exclusive scan of the B1 counts)
// ---------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(ST_THREADS) k_stats_compact(const double* __restrict__ samples, const double* __restrict__ logp,
                                                             const double* __restrict__ logprior, StatsGeom G, StatsWork* W,
                                                             const double* __restrict__ out, double* __restrict__ o_samples,
                                                             double* __restrict__ o_logp, double* __restrict__ o_logprior) {
    __shared__ long long s_base;
    __shared__ int s_w[ST_THREADS / 32];
    if (threadIdx.x == 0) {
        long long b = 0;
        for (int c = 0; c < (int)blockIdx.x; ++c) b += W->counts[c];
        s_base = b;
    }
    __syncthreads();
    long long base = s_base;
    const double cut = out[PEUQSE_STATS_CUT];
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const long long r0 = (long long)blockIdx.x * G.rows_per_cta;
    const long long r1 = (r0 + G.rows_per_cta < G.n) ? r0 + G.rows_per_cta : G.n;
    for (long long t0 = r0; t0 < r1; t0 += ST_THREADS) {
        const long long i = t0 + threadIdx.x;
        const bool in = i < r1;
        const double lp = in ? logp[i] : 0.0;
        const bool keep = in && !(lp < cut);
        const unsigned int m = __ballot_sync(0xffffffffu, keep);
        if (lane == 0) s_w[warp] = __popc(m);
        __syncthreads();
        int before = 0, tot = 0;
        for (int w = 0; w < ST_THREADS / 32; ++w) {
            if (w < warp) before += s_w[w];
            tot += s_w[w];
        }
        if (keep) {
            const long long o = base + before + __popc(m & ((1u << lane) - 1u));
            o_logp[o] = lp;
            if (o_logprior && logprior) o_logprior[o] = logprior[i];
            for (int d = 0; d < G.k; ++d) o_samples[o * G.k + d] = samples[i * G.k + d];
        }
        base += tot;
        __syncthreads();
    }
}

// ---------------------------------------------------------------------------------------------
// C ABI
// ---------------------------------------------------------------------------------------------
static int stats_sm_count() {
    int dev = 0, sms = 148;
    if (cudaGetDevice(&dev) == cudaSuccess) cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
    return sms > 0 ? sms : 148;
}

static int stats_check(const peuqse_stats_spec* s, const void* d_samples, const void* d_logp, int64_t n, const void* d_out,
                       const void* d_work) {
    if (!s || !d_samples || !d_logp || !d_out || !d_work) return pq_set_error(PEUQSE_ERR_INVALID, "null argument");
    if (s->k < 1 || s->k > PQ_MAXK) return pq_set_error(PEUQSE_ERR_INVALID, "k out of range");
    if (n < 1) return pq_set_error(PEUQSE_ERR_INVALID, "statistics need at least one row");
    if (s->filter < PEUQSE_FILTER_NONE || s->filter > PEUQSE_FILTER_MU_AP) return pq_set_error(PEUQSE_ERR_INVALID, "bad filter mode");
    if (s->kl_bins < 0 || s->kl_bins > ST_MAXBINS) return pq_set_error(PEUQSE_ERR_INVALID, "kl_bins out of range");
    return PEUQSE_OK;
}

extern "C" int64_t peuqse_stats_work_bytes(void) { return (int64_t)sizeof(StatsWork); }

extern "C" int peuqse_stats_begin(const peuqse_stats_spec* s, const double* d_samples, const double* d_logp, int64_t n,
                                  double* d_out, void* d_work, void* stream) {
    int rc = stats_check(s, d_samples, d_logp, n, d_out, d_work);
    if (rc) return rc;
    cudaStream_t st = (cudaStream_t)stream;
    StatsWork* W = (StatsWork*)d_work;
    const StatsGeom G = stats_geom(n, s->k, stats_sm_count());
    PQ_CUDA_OK(cudaMemsetAsync(W->ticket, 0, sizeof(W->ticket), st));
    PQ_CUDA_OK(cudaMemsetAsync(d_out, 0, PEUQSE_STATS_DOUBLES * sizeof(double), st));
    const bool vec2 = (s->k % 2 == 0) && ((((uintptr_t)d_samples) & 15) == 0);
    const bool theta = s->filter == PEUQSE_FILTER_MU_AP;
    const int kb = pq_k_bucket(s->k);
#define SA1(KT, ...)                                                                                          \
    {                                                                                                         \
        if (theta) {                                                                                          \
            if (vec2) k_stats_a1<KT, true, true><<<G.grid, ST_THREADS, 0, st>>>(d_samples, d_logp, G, W, d_out);   \
            else k_stats_a1<KT, false, true><<<G.grid, ST_THREADS, 0, st>>>(d_samples, d_logp, G, W, d_out);       \
        } else k_stats_a1<KT, false, false><<<G.grid, ST_THREADS, 0, st>>>(d_samples, d_logp, G, W, d_out);        \
    }
    PQ_DISPATCH_K(kb, SA1, 0)
#undef SA1
    PQ_CUDA_OK(cudaGetLastError());
    return PEUQSE_OK;
}

extern "C" int peuqse_stats_finish(const peuqse_stats_spec* s, const double* d_samples, const double* d_logp,
                                   const double* d_logprior, int64_t n, const double* d_mu_ap_logp,
                                   const double* d_prior_samples, int64_t n_prior, double* d_out, void* d_work, void* stream) {
    int rc = stats_check(s, d_samples, d_logp, n, d_out, d_work);
    if (rc) return rc;
    if (s->filter == PEUQSE_FILTER_MU_AP && !d_mu_ap_logp) return pq_set_error(PEUQSE_ERR_INVALID, "the mu_AP benchmark needs d_mu_ap_logp");
    cudaStream_t st = (cudaStream_t)stream;
    StatsWork* W = (StatsWork*)d_work;
    const StatsGeom G = stats_geom(n, s->k, stats_sm_count());
    const bool kl = d_prior_samples != nullptr && n_prior > 0 && s->kl_bins > 0;
    const int bins = kl ? s->kl_bins : 0;
    const bool vec2 = (s->k % 2 == 0) && ((((uintptr_t)d_samples) & 15) == 0);
    const int kb = pq_k_bucket(s->k);
    k_stats_cut<<<1, 32, 0, st>>>(s->filter, s->coef, d_mu_ap_logp, d_out, W, s->k, bins);
    // TMA ring: 16-byte aligned sources (tile starts are multiples of 256 rows) and room for three stages
    const size_t ring = (size_t)ST_STAGES * ST_TILE * (s->k + 2) * sizeof(double);
    int dev = 0, max_optin = 48 * 1024;
    cudaGetDevice(&dev);
    cudaDeviceGetAttribute(&max_optin, cudaDevAttrMaxSharedMemoryPerBlockOptin, dev);
    int copies = 8;                                     // private histogram copies per lane group, as many as fit
    auto hist_bytes = [&](int c) { return (size_t)s->k * (bins + 1) * sizeof(double) + (size_t)s->k * bins * c * sizeof(unsigned int); };
    while (copies > 1 && ring + hist_bytes(copies) + 2048 > (size_t)max_optin / 2) copies >>= 1;
    const size_t smem_hist = kl ? hist_bytes(copies) : 0;
    const bool aligned = ((((uintptr_t)d_samples) | ((uintptr_t)d_logp) | ((uintptr_t)d_logprior)) & 15) == 0;
    const bool tma = aligned && ring + smem_hist + 2048 <= (size_t)max_optin;
    const size_t smem1 = tma ? ring : 0, smem2 = (tma ? ring : 0) + smem_hist;
#define SB1L(KT, V, R)                                                                                                         \
    {                                                                                                                         \
        cudaError_t e = pq_allow_smem((const void*)k_stats_b1<KT, V, R>, smem1);                                             \
        if (e != cudaSuccess) return pq_set_cuda_error(e, "cudaFuncSetAttribute", __LINE__);                                  \
        k_stats_b1<KT, V, R><<<G.grid, ST_THREADS, smem1, st>>>(d_samples, d_logp, G, W, d_out, tma);                          \
    }
#define SB1(KT, ...)                                                                                                          \
    {                                                                                                                         \
        if (vec2 && kl) SB1L(KT, true, true) else if (vec2) SB1L(KT, true, false) else if (kl) SB1L(KT, false, true) else SB1L(KT, false, false) \
    }
    PQ_DISPATCH_K(kb, SB1, 0)
#undef SB1
#undef SB1L
    if (kl) k_stats_edges<<<s->k, ST_THREADS, 0, st>>>(d_prior_samples, n_prior, s->k, bins, d_out, W);
#define SB2L(KT, V)                                                                                                           \
    {                                                                                                                         \
        cudaError_t e = pq_allow_smem((const void*)k_stats_b2<KT, V>, smem2);                                                \
        if (e != cudaSuccess) return pq_set_cuda_error(e, "cudaFuncSetAttribute", __LINE__);                                  \
        k_stats_b2<KT, V><<<G.grid, ST_THREADS, smem2, st>>>(d_samples, d_logp, d_logprior, G, W, d_out, bins, copies, tma);  \
    }
#define SB2(KT, ...) { if (vec2) SB2L(KT, true) else SB2L(KT, false) }
    PQ_DISPATCH_K(kb, SB2, 0)
#undef SB2
#undef SB2L
    if (kl) {
        const long long tot = n_prior * s->k;
        const int g = (int)std::min<long long>((tot + 8 * ST_THREADS - 1) / (8 * ST_THREADS), 148);
        k_stats_prior_hist<<<g, ST_THREADS, (size_t)s->k * bins * sizeof(unsigned int), st>>>(d_prior_samples, n_prior, s->k, bins, W);
    }
    k_stats_final<<<1, ST_THREADS, 0, st>>>(d_samples, s->k, bins, kl, d_out, W);
    PQ_CUDA_OK(cudaGetLastError());
    return PEUQSE_OK;
}

extern "C" int peuqse_stats_compact(const peuqse_stats_spec* s, const double* d_samples, const double* d_logp,
                                    const double* d_logprior, int64_t n, const double* d_out, double* d_out_samples,
                                    double* d_out_logp, double* d_out_logprior, void* d_work, void* stream) {
    int rc = stats_check(s, d_samples, d_logp, n, d_out, d_work);
    if (rc) return rc;
    if (!d_out_samples || !d_out_logp) return pq_set_error(PEUQSE_ERR_INVALID, "compact needs output buffers");
    const StatsGeom G = stats_geom(n, s->k, stats_sm_count());
    k_stats_compact<<<G.grid, ST_THREADS, 0, (cudaStream_t)stream>>>(d_samples, d_logp, d_logprior, G, (StatsWork*)d_work, d_out,
                                                                   d_out_samples, d_out_logp, d_out_logprior);
    PQ_CUDA_OK(cudaGetLastError());
    return PEUQSE_OK;
}
