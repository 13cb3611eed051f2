// Design-of-experiments scenario batches: doeGetInfoGainMatrix and
// doeParameterModulationPermutationsScanner (IP.py:1363-1572) run one whole Metropolis-Hastings chain per
// (parameter modulation, condition) pair, each on synthetic observations generated under that pair
// (populateResponsesWithSyntheticData, IP.py:1326-1354), and keep the information gain of every run.
// Here the S pairs are ONE batch: scenario s owns a row of model constants (the independent variables the reference
// pushes into the simulator's namespace) and a row of observed values; the MH kernels of mh.cu walk S x chains chains at
// once (peuqse_mh_run_scenarios) and k_scen_stats reduces every scenario's post-burn-in rows (calculatePostBurnInStatistics
// IP.py:1861-1937 with the 'auto' / 'MAP' tail filter, calculateInfoGain's log-ratio form IP.py:1793-1803) in one launch,
// so the host reads back S x (8 + 3k) doubles for the whole info-gain matrix.
#include <vector>
#include "common.cuh"
#include "eval.cuh"
#include "launch.cuh"

namespace {

// y_s[j] = f(theta_s; constants of scenario s), device row order; optionally also in the caller's natural order
template <int MODEL, int KT>
__global__ void __launch_bounds__(128) k_scen_synth(const __grid_constant__ DevProblem P, const DevScen SC, long long S,
                                                    const double* __restrict__ theta, long long theta_stride,
                                                    double* __restrict__ y, double* __restrict__ y_natural) {
    const long long e = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    const long long total = S * P.n_total;
    if (e >= total) return;
    const long long s = e / P.n_total;
    const int j = (int)(e - s * P.n_total);
    ObsView O;
    O.y = nullptr; O.w = P.w; O.cval = P.cval;
    O.mc = SC.mc + s * SC.n_mconst;
    double th[KT];
#pragma unroll
    for (int d = 0; d < KT; ++d) th[d] = (d < P.k) ? theta[s * theta_stride + d] : 0.0;
    const double v = Sim<MODEL, KT>::at(P, O, th, nullptr, j);
    y[s * P.n_total + j] = v;
    if (y_natural) y_natural[s * P.n_total + P.src[j]] = v;
}

// natural-order observed rows -> device row order
__global__ void k_scen_set_y(const __grid_constant__ DevProblem P, long long S, const double* __restrict__ y_natural,
                             double* __restrict__ y) {
    const long long e = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (e >= S * P.n_total) return;
    const long long s = e / P.n_total;
    const int j = (int)(e - s * P.n_total);
    y[e] = y_natural[s * P.n_total + P.src[j]];
}

// getLogP / getLogPrior / getLogLikelihood of theta[S][B][k] under each scenario's observations (thread shape)
template <int MODEL, int KT>
__global__ void __launch_bounds__(128) k_scen_logp(const __grid_constant__ DevProblem P, const DevScen SC, long long S, long long B,
                                                   const double* __restrict__ theta, double* __restrict__ logp,
                                                   double* __restrict__ lprior, double* __restrict__ llik) {
    const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= S * B) return;
    const long long s = i / B;
    ObsView O;
    O.y = SC.y + s * SC.n_total; O.w = P.w; O.cval = P.cval;
    O.mc = SC.mc + s * SC.n_mconst;
    double th[KT];
#pragma unroll
    for (int d = 0; d < KT; ++d) th[d] = (d < P.k) ? theta[i * P.k + d] : 0.0;
    const double lp = eval_prior<KT>(P, th);
    const double ll = eval_loglik_thread<MODEL, KT>(P, O, th, nullptr);
    if (logp) logp[i] = combine_logp(lp, ll);
    if (lprior) lprior[i] = lp;
    if (llik) llik[i] = ll;
}

// the same for many observations: one warp per (scenario, point), lanes stride over the observations
template <int MODEL, int KT>
__global__ void __launch_bounds__(128) k_scen_logp_warp(const __grid_constant__ DevProblem P, const DevScen SC, long long S, long long B,
                                                        const double* __restrict__ theta, double* __restrict__ logp,
                                                        double* __restrict__ lprior, double* __restrict__ llik) {
    const long long i = (long long)blockIdx.x * 4 + (threadIdx.x >> 5);
    const int lane = threadIdx.x & 31;
    if (i >= S * B) return;
    const long long s = i / B;
    ObsView O;
    O.y = SC.y + s * SC.n_total; O.w = P.w; O.cval = P.cval;
    O.mc = SC.mc + s * SC.n_mconst;
    double th[KT];
#pragma unroll
    for (int d = 0; d < KT; ++d) th[d] = (d < P.k) ? theta[i * P.k + d] : 0.0;
    const double lp = eval_prior<KT>(P, th);
    const double ll = eval_loglik_warp<MODEL, KT>(P, O, th, nullptr, lane);
    if (lane == 0) {
        if (logp) logp[i] = combine_logp(lp, ll);
        if (lprior) lprior[i] = lp;
        if (llik) llik[i] = ll;
    }
}

// ---------------------------------------------------------------------------------------------------
// per-scenario post-burn-in statistics
// ---------------------------------------------------------------------------------------------------
#define SS_THREADS 256
#define SS_WARPS (SS_THREADS / 32)

// rows of a scenario in the reference's stacking order (chain after chain, IP.py:1282-1319): i = chain*rows + r
struct ScenRows {
    long long C, per, rows, first_row;      // chains in the launch, chains per scenario, post-burn-in rows per chain, first such row
};

// Two ways of handing a scenario's rows to threads.  The MH kernels store row-major over chains (samples[r][c][k],
// coalesced for the sampler), so a scenario's rows are C*8 bytes apart.
//   ColumnTiles: a CTA owns 32 ADJACENT chains (32/per scenarios); lane = chain, the 8 warps take interleaved rows, so every
//                warp load is one contiguous 256-byte run of a row; per-scenario totals = 8 row slices (shared memory, in
//                order) then the scenario's `per` adjacent lanes (shuffle butterfly).  Needs per in {1,2,4,8,16,32}.
//   CtaPerScenario: any `per`; one CTA per scenario with strided reads.
struct ColumnTiles {
    const ScenRows& R;
    double (*red)[32];
    long long (*redi)[32];
    int lane, w;
    long long c, s, ch;
    bool live;
    __device__ ColumnTiles(const ScenRows& R_, double (*red_)[32], long long (*redi_)[32]) : R(R_), red(red_), redi(redi_) {
        lane = threadIdx.x & 31; w = threadIdx.x >> 5;
        c = (long long)blockIdx.x * 32 + lane;
        live = c < R.C;
        s = c / R.per; ch = c - s * R.per;
    }
    __device__ __forceinline__ long long begin() const { return live ? w : R.rows; }
    __device__ __forceinline__ long long end() const { return R.rows; }
    __device__ __forceinline__ long long step() const { return SS_WARPS; }
    __device__ __forceinline__ long long index(long long r) const { return ch * R.rows + r; }
    __device__ __forceinline__ long long mem(long long r) const { return (R.first_row + r) * R.C + c; }
    __device__ __forceinline__ long long mem_of_index(long long i) const {
        const long long h = i / R.rows;
        return (R.first_row + (i - h * R.rows)) * R.C + (s * R.per + h);
    }
    __device__ __forceinline__ bool writer() const { return live && w == 0 && ch == 0; }
    __device__ __forceinline__ double sum(double v) const {
        __syncthreads();
        red[w][lane] = v;
        __syncthreads();
        double t = 0.0;
#pragma unroll
        for (int q = 0; q < SS_WARPS; ++q) t += red[q][lane];
        for (int o = (int)R.per >> 1; o > 0; o >>= 1) t += __shfl_xor_sync(0xffffffffu, t, o);
        return t;
    }
    __device__ __forceinline__ Best best(Best b) const {
        __syncthreads();
        red[w][lane] = b.v; redi[w][lane] = b.i;
        __syncthreads();
        Best t = best_none();
#pragma unroll
        for (int q = 0; q < SS_WARPS; ++q) t = best_merge(t, Best{red[q][lane], redi[q][lane]});
        for (int o = (int)R.per >> 1; o > 0; o >>= 1) {
            Best y;
            y.v = __shfl_xor_sync(0xffffffffu, t.v, o);
            y.i = __shfl_xor_sync(0xffffffffu, t.i, o);
            t = best_merge(t, y);
        }
        return t;
    }
};

struct CtaPerScenario {
    const ScenRows& R;
    double (*red)[32];
    long long (*redi)[32];
    long long s;
    __device__ CtaPerScenario(const ScenRows& R_, double (*red_)[32], long long (*redi_)[32]) : R(R_), red(red_), redi(redi_) {
        s = blockIdx.x;
    }
    __device__ __forceinline__ long long begin() const { return threadIdx.x; }
    __device__ __forceinline__ long long end() const { return R.per * R.rows; }
    __device__ __forceinline__ long long step() const { return SS_THREADS; }
    __device__ __forceinline__ long long index(long long i) const { return i; }
    __device__ __forceinline__ long long mem(long long i) const { return mem_of_index(i); }
    __device__ __forceinline__ long long mem_of_index(long long i) const {
        const long long h = i / R.rows;
        return (R.first_row + (i - h * R.rows)) * R.C + (s * R.per + h);
    }
    __device__ __forceinline__ bool writer() const { return threadIdx.x == 0; }
    __device__ __forceinline__ double sum(double v) const {
        v = warp_sum(v);
        __syncthreads();
        if ((threadIdx.x & 31) == 0) red[0][threadIdx.x >> 5] = v;
        __syncthreads();
        double t = 0.0;
#pragma unroll
        for (int q = 0; q < SS_WARPS; ++q) t += red[0][q];
        return t;
    }
    __device__ __forceinline__ Best best(Best b) const {
        b = warp_best(b);
        __syncthreads();
        if ((threadIdx.x & 31) == 0) { red[0][threadIdx.x >> 5] = b.v; redi[0][threadIdx.x >> 5] = b.i; }
        __syncthreads();
        Best t = best_none();
#pragma unroll
        for (int q = 0; q < SS_WARPS; ++q) t = best_merge(t, Best{red[0][q], redi[0][q]});
        return t;
    }
};

// U rows per trip with every load issued before the first use: a thread's rows are independent but each trip ends in
// transcendental chains (log / exp), so without this a warp has a single 256-byte load in flight and the kernel runs at the
// memory latency (measured: 1.3 TB/s with one row per trip).
template <int U, class MAP, class F>
__device__ __forceinline__ void for_rows_logp(const MAP& T, const double* __restrict__ logp, F f) {
    const long long b1 = T.end(), bs = T.step();
    for (long long r = T.begin(); r < b1; r += U * bs) {
        double lp[U];
#pragma unroll
        for (int u = 0; u < U; ++u) lp[u] = (r + u * bs < b1) ? logp[T.mem(r + u * bs)] : 0.0;
#pragma unroll
        for (int u = 0; u < U; ++u)
            if (r + u * bs < b1) f(r + u * bs, lp[u]);
    }
}

template <int KT, int U, class MAP, class F>
__device__ __forceinline__ void for_rows_all(const MAP& T, int k, const double* __restrict__ samples, const double* __restrict__ logp,
                                             const double* __restrict__ lprior, F f) {
    const long long b1 = T.end(), bs = T.step();
    for (long long r = T.begin(); r < b1; r += U * bs) {
        double lp[U], pr[U], th[U][KT];
#pragma unroll
        for (int u = 0; u < U; ++u) {
            const bool ok = r + u * bs < b1;
            const long long row = ok ? T.mem(r + u * bs) : 0;
            lp[u] = ok ? logp[row] : 0.0;
            pr[u] = (ok && lprior) ? lprior[row] : 0.0;
#pragma unroll
            for (int d = 0; d < KT; ++d) th[u][d] = (ok && d < k) ? samples[row * k + d] : 0.0;
        }
#pragma unroll
        for (int u = 0; u < U; ++u)
            if (r + u * bs < b1) f(r + u * bs, lp[u], pr[u], th[u]);
    }
}

template <int KT, class MAP>
__global__ void __launch_bounds__(SS_THREADS) k_scen_stats(const double* __restrict__ samples, const double* __restrict__ logp,
                                                           const double* __restrict__ lprior, const ScenRows R, int k, int filter,
                                                           double coef, double* __restrict__ out) {
    constexpr int U = KT <= 4 ? 4 : (KT <= 8 ? 2 : 1);
    __shared__ double red[SS_WARPS][32];
    __shared__ long long redi[SS_WARPS][32];
    const MAP T(R, red, redi);
    const double n = (double)(R.per * R.rows);
    // ---- threshold of the tail filter (IP.py:1887-1918), turned into ONE rule: a row is removed iff logP < cut ----
    double thr = __longlong_as_double(0x7ff8000000000000LL);     // the reference's filteringThreshold
    double cut = pq_neg_inf();
    if (filter == PEUQSE_FILTER_AUTO) {
        // mean and standard deviation of log(-logP) (IP.py:1894-1896) in one pass about a pivot (the scenario's first row)
        const double pivot = log(-1.0 * logp[T.mem_of_index(0)]);
        double a = 0.0, a2 = 0.0;
        for_rows_logp<8>(T, logp, [&](long long, double lp) { const double d = log(-1.0 * lp) - pivot; a += d; a2 = fma(d, d, a2); });
        const double m1 = T.sum(a) / n, m2 = T.sum(a2) / n;
        const double sd = sqrt(fmax(m2 - m1 * m1, 0.0));                                               // np.std, ddof 0
        thr = (pivot + m1) + coef * sd;
        // rows with log(-logP) > thr go (IP.py:1915, 3650).  log is monotone: x0 = the largest double with log(x0) <= thr,
        // removed <=> -logP > x0 <=> logP < -x0.  A NaN threshold (some logP >= 0) removes nothing, NaN rows are kept.
        double c = 0.0;
        if (T.writer()) {
            c = pq_neg_inf();
            if (thr == thr && thr < pq_pos_inf()) {
                double x = exp(thr);
                if (!(x < pq_pos_inf())) x = 1.7976931348623157e308;
                for (int it = 0; it < 64 && log(x) > thr; ++it) x = __longlong_as_double(__double_as_longlong(x) - 1);
                for (int it = 0; it < 64; ++it) {
                    const double xn = __longlong_as_double(__double_as_longlong(x) + 1);
                    if (!(xn < pq_pos_inf()) || log(xn) > thr) break;
                    x = xn;
                }
                c = -x;
            }
        }
        cut = T.sum(c);                                          // one writer per scenario: broadcast
    } else if (filter == PEUQSE_FILTER_MAP) {
        // np.max propagates NaN: a NaN anywhere makes the threshold NaN and nothing is removed
        double m = pq_neg_inf(), nan = 0.0;
        for_rows_logp<8>(T, logp, [&](long long, double v) { if (v != v) nan = 1.0; m = fmax(m, v); });
        const Best b = T.best(Best{m, 0});
        if (!(T.sum(nan) > 0.0)) { thr = b.v - coef; cut = thr; }                                      // IP.py:1897, 1910
    }
    // ---- pass B1 over kept rows: count, sum theta, sum theta^2, first maximum, sum exp(logP); branch-free so that the
    //      U rows' exp chains interleave ----
    double cnt = 0.0, sexp = 0.0, frob = 0.0, s1[KT];
#pragma unroll
    for (int d = 0; d < KT; ++d) s1[d] = 0.0;
    Best best = best_none();
    for_rows_all<KT, U>(T, k, samples, logp, nullptr, [&](long long r, double lp, double, const double (&th)[KT]) {
        const bool kept = !(lp < cut);
        const double e = exp(lp);
        cnt += kept ? 1.0 : 0.0;
        sexp += kept ? e : 0.0;
        if (kept && lp > best.v) { best.v = lp; best.i = T.index(r); }     // indices ascend per thread: first maximum
#pragma unroll
        for (int d = 0; d < KT; ++d)
            if (d < k) { const double v = kept ? th[d] : 0.0; s1[d] += v; frob = fma(v, v, frob); }
    });
    const double nk = T.sum(cnt);
    const double sumexp = T.sum(sexp);
    const double fro2 = T.sum(frob);
    double mu[KT];
#pragma unroll
    for (int d = 0; d < KT; ++d) mu[d] = T.sum(s1[d]) / nk;                                            // IP.py:1923
    best = T.best(best);
    const double evidence = (sumexp / nk) / sqrt(fro2);                                                // IP.py:1931
    const double logev = log(evidence);
    // ---- pass B2: sum (theta - mu)^2, log-ratio terms, rank of the MAP row among the kept rows ----
    double s2[KT], ratio = 0.0, before = 0.0;
#pragma unroll
    for (int d = 0; d < KT; ++d) s2[d] = 0.0;
    for_rows_all<KT, U>(T, k, samples, logp, lprior, [&](long long r, double lp, double pr, const double (&th)[KT]) {
        const bool kept = !(lp < cut);
        before += (kept && T.index(r) < best.i) ? 1.0 : 0.0;
#pragma unroll
        for (int d = 0; d < KT; ++d)
            if (d < k) { const double dv = kept ? th[d] - mu[d] : 0.0; s2[d] = fma(dv, dv, s2[d]); }
        if (lprior) {
            double q = (lp - logev) - pr;                                                              // IP.py:1799-1800
            if (!kept || isinf(q) || q != q) q = 0.0;                                                  // IP.py:1801-1802
            ratio += q;
        }
    });
    const double nbefore = T.sum(before);
    const double ig = T.sum(ratio) / nk;                                                               // IP.py:1803
    double sd[KT];
#pragma unroll
    for (int d = 0; d < KT; ++d) sd[d] = sqrt(T.sum(s2[d]) / nk);                                      // IP.py:1924
    if (T.writer()) {
        double* o = out + T.s * PEUQSE_SCEN_STATS_DOUBLES(k);
        o[PEUQSE_SCEN_STATS_N_TOTAL] = n;
        o[PEUQSE_SCEN_STATS_N_KEPT] = nk;
        o[PEUQSE_SCEN_STATS_THRESHOLD] = thr;
        o[PEUQSE_SCEN_STATS_MAP_LOGP] = best.v;
        o[PEUQSE_SCEN_STATS_MAP_INDEX] = nbefore;
        o[PEUQSE_SCEN_STATS_EVIDENCE] = evidence;
        o[PEUQSE_SCEN_STATS_INFO_GAIN_LOG_RATIO] = lprior ? ig : __longlong_as_double(0x7ff8000000000000LL);
        o[PEUQSE_SCEN_STATS_CUT] = cut;
        const long long maprow = nk > 0.0 ? T.mem_of_index(best.i) : 0;
#pragma unroll
        for (int d = 0; d < KT; ++d)
            if (d < k) {
                o[PEUQSE_SCEN_STATS_MU_AP + d] = mu[d];
                o[PEUQSE_SCEN_STATS_STDAP(k) + d] = sd[d];
                o[PEUQSE_SCEN_STATS_MAP_THETA(k) + d] = samples[maprow * k + d];
            }
    }
}

// ---------------------------------------------------------------------------------------------------
// per-scenario KL-divergence information gain (calculateInfoGain, IP.py:1772-1791): one CTA per scenario bins the kept
// posterior rows and the (shared) prior samples on that scenario's joint edges, np.histogram semantics, density=True
// ---------------------------------------------------------------------------------------------------
#define SK_MAXBINS 128

// np.histogram bin of x (numpy/lib/_histograms_impl.py), -1 outside [lo, hi]
__device__ __forceinline__ int sk_hist_bin(const double* __restrict__ ed, int nbins, double x) {
    const double lo = ed[0], hi = ed[nbins];
    if (!(x >= lo && x <= hi)) return -1;
    int b = (int)(((x - lo) / (hi - lo)) * (double)nbins);
    if (b >= nbins) b = nbins - 1;
    if (b < 0) b = 0;
    if (x < ed[b]) --b;
    else if (b != nbins - 1 && x >= ed[b + 1]) ++b;
    return b;
}

__global__ void __launch_bounds__(SS_THREADS) k_scen_kl(const double* __restrict__ samples, const double* __restrict__ logp,
                                                        const ScenRows R, int k, int bins, const double* __restrict__ prior,
                                                        long long n_prior, const double* __restrict__ stats, int stats_stride,
                                                        double* __restrict__ out) {
    extern __shared__ __align__(16) unsigned char sk_raw[];
    double* s_edges = (double*)sk_raw;                                   // [k][bins + 1]
    unsigned int* s_h0 = (unsigned int*)(s_edges + (size_t)k * (bins + 1));   // prior counts [k][bins]
    unsigned int* s_h1 = s_h0 + (size_t)k * bins;                        // posterior counts
    __shared__ double s_lo[PQ_MAXK][SS_WARPS], s_hi[PQ_MAXK][SS_WARPS];
    __shared__ double s_each[PQ_MAXK];
    const long long s = blockIdx.x;
    const long long n = R.per * R.rows;
    const int t = threadIdx.x, lane = t & 31, w = t >> 5;
    const double cut = stats[s * stats_stride + PEUQSE_SCEN_STATS_CUT];
    auto mem = [&](long long i) {
        const long long h = i / R.rows;
        return (R.first_row + (i - h * R.rows)) * R.C + (s * R.per + h);
    };
    for (int a = t; a < 2 * k * bins; a += SS_THREADS) s_h0[a] = 0u;
    // joint range of the kept posterior rows and the prior samples, per parameter
    for (int d = 0; d < k; ++d) {
        double lo = pq_pos_inf(), hi = pq_neg_inf();
        for (long long i = t; i < n; i += SS_THREADS) {
            const long long row = mem(i);
            if (logp[row] < cut) continue;
            const double v = samples[row * k + d];
            lo = fmin(lo, v); hi = fmax(hi, v);
        }
        for (long long i = t; i < n_prior; i += SS_THREADS) {
            const double v = prior[i * k + d];
            lo = fmin(lo, v); hi = fmax(hi, v);
        }
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) {
            lo = fmin(lo, __shfl_xor_sync(0xffffffffu, lo, o));
            hi = fmax(hi, __shfl_xor_sync(0xffffffffu, hi, o));
        }
        if (lane == 0) { s_lo[d][w] = lo; s_hi[d][w] = hi; }
    }
    __syncthreads();
    for (int d = 0; d < k; ++d) {
        double lo = s_lo[d][0], hi = s_hi[d][0];
        for (int q = 1; q < SS_WARPS; ++q) { lo = fmin(lo, s_lo[d][q]); hi = fmax(hi, s_hi[d][q]); }
        if (!(hi > lo)) { lo = lo - 0.5; hi = hi + 0.5; }               // np.histogram widens an empty range by +-0.5
        const double step = (hi - lo) / (double)bins;                    // np.linspace(lo, hi, bins + 1)
        for (int i = t; i <= bins; i += SS_THREADS) s_edges[d * (bins + 1) + i] = (i == bins) ? hi : ((double)i * step + lo);
    }
    __syncthreads();
    for (long long i = t; i < n; i += SS_THREADS) {
        const long long row = mem(i);
        if (logp[row] < cut) continue;
        for (int d = 0; d < k; ++d) {
            const int b = sk_hist_bin(s_edges + d * (bins + 1), bins, samples[row * k + d]);
            if (b >= 0) atomicAdd(&s_h1[d * bins + b], 1u);
        }
    }
    for (long long e = t; e < n_prior * k; e += SS_THREADS) {
        const int d = (int)(e % k);
        const int b = sk_hist_bin(s_edges + d * (bins + 1), bins, prior[e]);
        if (b >= 0) atomicAdd(&s_h0[d * bins + b], 1u);
    }
    __syncthreads();
    if (t < k) {
        const int d = t;
        const double* ed = s_edges + d * (bins + 1);
        double n0 = 0.0, n1 = 0.0;
        for (int b = 0; b < bins; ++b) { n0 += (double)s_h0[d * bins + b]; n1 += (double)s_h1[d * bins + b]; }
        double acc = 0.0;
        for (int b = 0; b < bins; ++b) {
            const double db = ed[b + 1] - ed[b];
            const double d0 = (double)s_h0[d * bins + b] / db / n0;      // counts / db / counts.sum()
            const double d1 = (double)s_h1[d * bins + b] / db / n1;
            const double v = d1 * log(d1 / d0);                          // IP.py:1785
            if (v == v && fabs(v) < pq_pos_inf()) acc += v;              // np.isfinite filter (IP.py:1786-1787)
        }
        s_each[d] = acc;
        out[s * (1 + k) + 1 + d] = acc;
    }
    __syncthreads();
    if (t == 0) {
        double tot = 0.0;
        for (int d = 0; d < k; ++d) tot = tot + s_each[d];               // IP.py:1789
        out[s * (1 + k)] = tot;
    }
}

#define SCEN_MODEL_DISPATCH(CALL)                                                                \
    switch (p->d.model_id) {                                                                      \
        case PEUQSE_MODEL_QUADRATIC: { CALL(PEUQSE_MODEL_QUADRATIC, 2); } break;                  \
        case PEUQSE_MODEL_TPD2SITE: { CALL(PEUQSE_MODEL_TPD2SITE, 6); } break;                    \
        case PEUQSE_MODEL_LANGMUIR_LOG: { CALL(PEUQSE_MODEL_LANGMUIR_LOG, 2); } break;            \
        case PEUQSE_MODEL_LINEAR: {                                                               \
            switch (pq_k_bucket(p->d.k)) {                                                        \
                case 2: { CALL(PEUQSE_MODEL_LINEAR, 2); } break;                                  \
                case 3: { CALL(PEUQSE_MODEL_LINEAR, 3); } break;                                  \
                case 4: { CALL(PEUQSE_MODEL_LINEAR, 4); } break;                                  \
                case 6: { CALL(PEUQSE_MODEL_LINEAR, 6); } break;                                  \
                case 8: { CALL(PEUQSE_MODEL_LINEAR, 8); } break;                                  \
                case 10: { CALL(PEUQSE_MODEL_LINEAR, 10); } break;                                \
                case 12: { CALL(PEUQSE_MODEL_LINEAR, 12); } break;                                \
                case 16: { CALL(PEUQSE_MODEL_LINEAR, 16); } break;                                \
                default: return pq_set_error(PEUQSE_ERR_INVALID, "k out of range");               \
            }                                                                                     \
        } break;                                                                                  \
        default: return pq_set_error(PEUQSE_ERR_INVALID, "scenario batches need a built-in model"); \
    }

}  // namespace

extern "C" int peuqse_scenarios_create(const peuqse_problem* p, int64_t S, const double* h_model_const, peuqse_scenarios** out) {
    if (!p || !h_model_const || !out || S < 1) return pq_set_error(PEUQSE_ERR_INVALID, "bad argument");
    if (p->d.model_id == PEUQSE_MODEL_EXTERNAL) return pq_set_error(PEUQSE_ERR_INVALID, "scenario batches need a built-in model");
    if (p->d.any_full) return pq_set_error(PEUQSE_ERR_INVALID, "scenario batches do not support full-covariance responses");
    const size_t nt = (size_t)p->d.n_total;
    const size_t nm = (size_t)p->d.n_mconst;                          // device layout (tpd2site: reciprocals appended)
    const size_t nuser = p->d.model_id == PEUQSE_MODEL_TPD2SITE ? 3 * nt : nm;
    std::vector<double> mc((size_t)S * nm);
    for (int64_t s = 0; s < S; ++s) {
        const double* src = h_model_const + (size_t)s * nuser;
        double* dst = mc.data() + (size_t)s * nm;
        for (size_t i = 0; i < nuser; ++i) dst[i] = src[i];
        if (p->d.model_id == PEUQSE_MODEL_TPD2SITE)
            for (size_t i = 0; i < nt; ++i) dst[3 * nt + i] = 1.0 / src[i];
    }
    peuqse_scenarios* sc = new peuqse_scenarios();
    sc->S = S; sc->n_total = (int)nt; sc->n_mconst = (int)nm; sc->has_y = 0; sc->pad = 0; sc->d_y = nullptr; sc->d_mc = nullptr;
    cudaError_t e = cudaMalloc((void**)&sc->d_mc, mc.size() * sizeof(double));
    if (e == cudaSuccess) e = cudaMalloc((void**)&sc->d_y, (size_t)S * nt * sizeof(double));
    if (e == cudaSuccess) e = cudaMemcpy(sc->d_mc, mc.data(), mc.size() * sizeof(double), cudaMemcpyHostToDevice);
    if (e != cudaSuccess) {
        if (sc->d_mc) cudaFree(sc->d_mc);
        if (sc->d_y) cudaFree(sc->d_y);
        delete sc;
        return pq_set_cuda_error(e, "scenario allocation", __LINE__);
    }
    *out = sc;
    return PEUQSE_OK;
}

extern "C" int peuqse_scenarios_destroy(peuqse_scenarios* sc) {
    if (!sc) return PEUQSE_OK;
    if (sc->d_mc) cudaFree(sc->d_mc);
    if (sc->d_y) cudaFree(sc->d_y);
    delete sc;
    return PEUQSE_OK;
}

static int scen_check(const peuqse_problem* p, const peuqse_scenarios* sc) {
    if (!p || !sc) return pq_set_error(PEUQSE_ERR_INVALID, "null argument");
    if (sc->n_total != p->d.n_total || sc->n_mconst != p->d.n_mconst) return pq_set_error(PEUQSE_ERR_INVALID, "scenarios belong to another problem");
    return PEUQSE_OK;
}

extern "C" int peuqse_scenarios_synthesize(const peuqse_problem* p, peuqse_scenarios* sc, const double* d_theta_true,
                                           int64_t theta_stride, double* d_y_natural, void* stream) {
    int rc = scen_check(p, sc);
    if (rc) return rc;
    if (!d_theta_true || (theta_stride != 0 && theta_stride < p->d.k)) return pq_set_error(PEUQSE_ERR_INVALID, "bad theta_true");
    const DevScen SC{sc->d_y, sc->d_mc, 1, sc->n_total, sc->n_mconst};
    const long long total = sc->S * sc->n_total;
    const unsigned grid = (unsigned)((total + 127) / 128);
    cudaStream_t st = (cudaStream_t)stream;
#define SYNTH(MODEL, KT) k_scen_synth<MODEL, KT><<<grid, 128, 0, st>>>(p->d, SC, sc->S, d_theta_true, theta_stride, sc->d_y, d_y_natural)
    SCEN_MODEL_DISPATCH(SYNTH)
#undef SYNTH
    PQ_CUDA_OK(cudaGetLastError());
    sc->has_y = 1;
    return PEUQSE_OK;
}

extern "C" int peuqse_scenarios_set_observed(const peuqse_problem* p, peuqse_scenarios* sc, const double* d_y_natural, void* stream) {
    int rc = scen_check(p, sc);
    if (rc) return rc;
    if (!d_y_natural) return pq_set_error(PEUQSE_ERR_INVALID, "null argument");
    const long long total = sc->S * sc->n_total;
    k_scen_set_y<<<(unsigned)((total + 255) / 256), 256, 0, (cudaStream_t)stream>>>(p->d, sc->S, d_y_natural, sc->d_y);
    PQ_CUDA_OK(cudaGetLastError());
    sc->has_y = 1;
    return PEUQSE_OK;
}

extern "C" int peuqse_scenarios_logp(const peuqse_problem* p, const peuqse_scenarios* sc, const double* d_theta, int64_t B,
                                     double* d_logp, double* d_logprior, double* d_loglik, void* stream) {
    int rc = scen_check(p, sc);
    if (rc) return rc;
    if (!sc->has_y) return pq_set_error(PEUQSE_ERR_INVALID, "scenarios have no observed values yet");
    if (!d_theta || B < 0) return pq_set_error(PEUQSE_ERR_INVALID, "bad argument");
    if (B == 0) return PEUQSE_OK;
    const DevScen SC{sc->d_y, sc->d_mc, (long long)B, sc->n_total, sc->n_mconst};
    const long long total = sc->S * B;
    const bool warp_shape = p->d.n_total > PQ_THREAD_SHAPE_MAX_N;
    const unsigned grid = (unsigned)(warp_shape ? (total + 3) / 4 : (total + 127) / 128);
    cudaStream_t st = (cudaStream_t)stream;
#define SLOGP(MODEL, KT)                                                                                                      \
    if (warp_shape) k_scen_logp_warp<MODEL, KT><<<grid, 128, 0, st>>>(p->d, SC, sc->S, B, d_theta, d_logp, d_logprior, d_loglik); \
    else k_scen_logp<MODEL, KT><<<grid, 128, 0, st>>>(p->d, SC, sc->S, B, d_theta, d_logp, d_logprior, d_loglik)
    SCEN_MODEL_DISPATCH(SLOGP)
#undef SLOGP
    PQ_CUDA_OK(cudaGetLastError());
    return PEUQSE_OK;
}

extern "C" int peuqse_scenarios_stats(int32_t k, int32_t filter, double coef, int64_t n_scenarios, int64_t chains_per_scenario,
                                      int64_t n_rows, int64_t first_row, const double* d_samples, const double* d_logp,
                                      const double* d_logprior, double* d_out, void* stream) {
    if (!d_samples || !d_logp || !d_out) return pq_set_error(PEUQSE_ERR_INVALID, "null argument");
    if (k < 1 || k > PEUQSE_MAX_K || n_scenarios < 1 || chains_per_scenario < 1 || first_row < 0 || n_rows <= first_row)
        return pq_set_error(PEUQSE_ERR_INVALID, "bad sizes");
    if (filter != PEUQSE_FILTER_NONE && filter != PEUQSE_FILTER_AUTO && filter != PEUQSE_FILTER_MAP)
        return pq_set_error(PEUQSE_ERR_INVALID, "scenario statistics support the 'auto' and 'MAP' filter benchmarks");
    ScenRows R;
    R.C = n_scenarios * chains_per_scenario; R.per = chains_per_scenario; R.rows = n_rows - first_row; R.first_row = first_row;
    cudaStream_t st = (cudaStream_t)stream;
    const bool tiles = chains_per_scenario <= 32 && (chains_per_scenario & (chains_per_scenario - 1)) == 0;
    const unsigned col_grid = (unsigned)((R.C + 31) / 32);
#define SSTAT(KT, ...)                                                                                                              \
    if (tiles) k_scen_stats<KT, ColumnTiles><<<col_grid, SS_THREADS, 0, st>>>(d_samples, d_logp, d_logprior, R, k, filter, coef, d_out); \
    else k_scen_stats<KT, CtaPerScenario><<<(unsigned)n_scenarios, SS_THREADS, 0, st>>>(d_samples, d_logp, d_logprior, R, k, filter, coef, d_out)
    PQ_DISPATCH_K(pq_k_bucket(k), SSTAT, 0)
#undef SSTAT
    PQ_CUDA_OK(cudaGetLastError());
    return PEUQSE_OK;
}

extern "C" int peuqse_scenarios_kl(int32_t k, int32_t bins, int64_t n_scenarios, int64_t chains_per_scenario, int64_t n_rows,
                                   int64_t first_row, const double* d_samples, const double* d_logp, const double* d_prior_samples,
                                   int64_t n_prior, const double* d_stats, double* d_out, void* stream) {
    if (!d_samples || !d_logp || !d_prior_samples || !d_stats || !d_out) return pq_set_error(PEUQSE_ERR_INVALID, "null argument");
    if (k < 1 || k > PEUQSE_MAX_K || bins < 1 || bins > SK_MAXBINS || n_scenarios < 1 || chains_per_scenario < 1 || first_row < 0 ||
        n_rows <= first_row || n_prior < 1)
        return pq_set_error(PEUQSE_ERR_INVALID, "bad sizes");
    ScenRows R;
    R.C = n_scenarios * chains_per_scenario; R.per = chains_per_scenario; R.rows = n_rows - first_row; R.first_row = first_row;
    const size_t smem = (size_t)k * (bins + 1) * sizeof(double) + 2 * (size_t)k * bins * sizeof(unsigned int);
    k_scen_kl<<<(unsigned)n_scenarios, SS_THREADS, smem, (cudaStream_t)stream>>>(d_samples, d_logp, R, k, bins, d_prior_samples, n_prior,
                                                                                d_stats, PEUQSE_SCEN_STATS_DOUBLES(k), d_out);
    PQ_CUDA_OK(cudaGetLastError());
    return PEUQSE_OK;
}
