// Design-of-experiments scenario batches: doeGetInfoGainMatrix and
// doeParameterModulationPermutationsScanner (IP.py:1363-1572) run one whole Metropolis-Hastings chain per
// (parameter modulation, condition) pair, each on synthetic observations generated under that pair
// (populateResponsesWithSyntheticData, IP.py:1326-1354), and keep the information gain of every run.
// Here the S pairs are ONE batch: scenario s owns a row of model constants (the independent variables the reference
// pushes into the simulator's namespace) and a row of observed values; the MH kernels of mh.cu walk S x chains chains at
// once (peuqse_mh_run_scenarios) and k_scen_stats reduces every scenario's post-burn-in rows (calculatePostBurnInStatistics
// IP.py:1861-1937 with the 'auto' / 'MAP' tail filter, calculateInfoGain's log-ratio form IP.py:1793-1803) with one CTA
// per scenario, so the host reads back S x 56 doubles for the whole info-gain matrix.
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

// ---------------------------------------------------------------------------------------------------
// per-scenario post-burn-in statistics: one CTA per scenario
// ---------------------------------------------------------------------------------------------------
#define SS_THREADS 256

// deterministic block sum: lanes by shuffle butterfly, warps in order by thread 0; result broadcast to every thread
__device__ __forceinline__ double block_sum(double v, double* red) {
    v = warp_sum(v);
    __syncthreads();
    if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = v;
    __syncthreads();
    double t = 0.0;
#pragma unroll
    for (int w = 0; w < SS_THREADS / 32; ++w) t += red[w];
    return t;
}
__device__ __forceinline__ Best block_best(Best b, double* redv, long long* redi) {
    b = warp_best(b);
    __syncthreads();
    if ((threadIdx.x & 31) == 0) { redv[threadIdx.x >> 5] = b.v; redi[threadIdx.x >> 5] = b.i; }
    __syncthreads();
    Best t = best_none();
#pragma unroll
    for (int w = 0; w < SS_THREADS / 32; ++w) t = best_merge(t, Best{redv[w], redi[w]});
    return t;
}

// rows of scenario s in the reference's stacking order (chain after chain, IP.py:1282-1319): i -> (chain, stored row)
struct ScenRows {
    long long C, per, rows, first_row;      // chains in the launch, chains per scenario, post-burn-in rows per chain, first such row
    __device__ __forceinline__ long long at(long long s, long long i) const {
        const long long ch = i / rows, r = first_row + (i - ch * rows);
        return r * C + (s * per + ch);
    }
};

template <int KT>
__global__ void __launch_bounds__(SS_THREADS) k_scen_stats(const double* __restrict__ samples, const double* __restrict__ logp,
                                                           const double* __restrict__ lprior, const ScenRows R, int k, int filter,
                                                           double coef, double* __restrict__ out) {
    __shared__ double red[SS_THREADS / 32];
    __shared__ long long redi[SS_THREADS / 32];
    const long long s = blockIdx.x;
    const long long n = R.per * R.rows;
    const int t = threadIdx.x;
    // ---- threshold of the tail filter (IP.py:1887-1918): a row is removed iff `drop(logP)` ----
    double thr = 0.0;                        // the reference's filteringThreshold
    if (filter == PEUQSE_FILTER_AUTO) {
        double a = 0.0;
        for (long long i = t; i < n; i += SS_THREADS) a += log(-1.0 * logp[R.at(s, i)]);              // IP.py:1894
        const double mean = block_sum(a, red) / (double)n;
        a = 0.0;
        for (long long i = t; i < n; i += SS_THREADS) { const double d = log(-1.0 * logp[R.at(s, i)]) - mean; a += d * d; }
        const double sd = sqrt(block_sum(a, red) / (double)n);                                         // np.std, ddof 0
        thr = mean + coef * sd;                                                                        // IP.py:1895-1896
    } else if (filter == PEUQSE_FILTER_MAP) {
        // np.max propagates NaN: a NaN anywhere makes the threshold NaN and nothing is removed
        double m = pq_neg_inf(); bool nan = false;
        for (long long i = t; i < n; i += SS_THREADS) { const double v = logp[R.at(s, i)]; nan = nan || (v != v); m = fmax(m, v); }
        const Best b = block_best(Best{m, (long long)t}, red, redi);
        const int anynan = __syncthreads_or(nan ? 1 : 0);
        thr = anynan ? __longlong_as_double(0x7ff8000000000000LL) : b.v - coef;                        // IP.py:1897, 1910
    }
    auto dropped = [&](double lp) -> bool {
        if (filter == PEUQSE_FILTER_AUTO) return log(-1.0 * lp) > thr;                                 // remove 'above', IP.py:1915, 3650
        if (filter == PEUQSE_FILTER_MAP) return lp < thr;                                              // remove 'below', IP.py:1910
        return false;
    };
    // ---- pass B1 over kept rows: count, sum theta, sum theta^2, first maximum, sum exp(logP) ----
    double cnt = 0.0, sexp = 0.0, frob = 0.0, s1[KT];
#pragma unroll
    for (int d = 0; d < KT; ++d) s1[d] = 0.0;
    Best best = best_none();
    for (long long i = t; i < n; i += SS_THREADS) {
        const long long row = R.at(s, i);
        const double lp = logp[row];
        if (dropped(lp)) continue;
        cnt += 1.0;
        sexp += exp(lp);
        if (lp > best.v) { best.v = lp; best.i = i; }          // i ascending per thread: first maximum
#pragma unroll
        for (int d = 0; d < KT; ++d)
            if (d < k) { const double v = samples[row * k + d]; s1[d] += v; frob = fma(v, v, frob); }
    }
    const double nk = block_sum(cnt, red);
    const double sumexp = block_sum(sexp, red);
    const double fro2 = block_sum(frob, red);
    double mu[KT];
#pragma unroll
    for (int d = 0; d < KT; ++d) mu[d] = block_sum(s1[d], red) / nk;                                   // IP.py:1923
    best = block_best(best, red, redi);
    const double evidence = (sumexp / nk) / sqrt(fro2);                                                // IP.py:1931
    const double logev = log(evidence);
    // ---- pass B2: sum (theta - mu)^2, log-ratio terms, rank of the MAP row among the kept rows ----
    double s2[KT], ratio = 0.0, before = 0.0;
#pragma unroll
    for (int d = 0; d < KT; ++d) s2[d] = 0.0;
    for (long long i = t; i < n; i += SS_THREADS) {
        const long long row = R.at(s, i);
        const double lp = logp[row];
        if (dropped(lp)) continue;
        if (i < best.i) before += 1.0;
#pragma unroll
        for (int d = 0; d < KT; ++d)
            if (d < k) { const double dv = samples[row * k + d] - mu[d]; s2[d] = fma(dv, dv, s2[d]); }
        if (lprior) {
            double r = (lp - logev) - lprior[row];                                                     // IP.py:1799-1800
            if (isinf(r) || r != r) r = 0.0;                                                           // IP.py:1801-1802
            ratio += r;
        }
    }
    const double nbefore = block_sum(before, red);
    const double ig = block_sum(ratio, red) / nk;                                                      // IP.py:1803
    double sd[KT];
#pragma unroll
    for (int d = 0; d < KT; ++d) sd[d] = sqrt(block_sum(s2[d], red) / nk);                             // IP.py:1924
    if (t == 0) {
        double* o = out + s * PEUQSE_SCEN_STATS_DOUBLES;
        o[PEUQSE_SCEN_STATS_N_TOTAL] = (double)n;
        o[PEUQSE_SCEN_STATS_N_KEPT] = nk;
        o[PEUQSE_SCEN_STATS_THRESHOLD] = thr;
        o[PEUQSE_SCEN_STATS_MAP_LOGP] = best.v;
        o[PEUQSE_SCEN_STATS_MAP_INDEX] = nbefore;
        o[PEUQSE_SCEN_STATS_EVIDENCE] = evidence;
        o[PEUQSE_SCEN_STATS_INFO_GAIN_LOG_RATIO] = lprior ? ig : __longlong_as_double(0x7ff8000000000000LL);
        o[7] = 0.0;
        const long long maprow = nk > 0.0 ? R.at(s, best.i) : 0;
        for (int d = 0; d < 16; ++d) {
            o[PEUQSE_SCEN_STATS_MU_AP + d] = 0.0; o[PEUQSE_SCEN_STATS_STDAP + d] = 0.0; o[PEUQSE_SCEN_STATS_MAP_THETA + d] = 0.0;
        }
#pragma unroll
        for (int d = 0; d < KT; ++d)
            if (d < k) {
                o[PEUQSE_SCEN_STATS_MU_AP + d] = mu[d];
                o[PEUQSE_SCEN_STATS_STDAP + d] = sd[d];
                o[PEUQSE_SCEN_STATS_MAP_THETA + d] = samples[maprow * k + d];
            }
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
    if (p->d.n_total > PQ_THREAD_SHAPE_MAX_N) return pq_set_error(PEUQSE_ERR_INVALID, "scenario batches support n_total <= 128");
    if (B == 0) return PEUQSE_OK;
    const DevScen SC{sc->d_y, sc->d_mc, (long long)B, sc->n_total, sc->n_mconst};
    const long long total = sc->S * B;
    const unsigned grid = (unsigned)((total + 127) / 128);
    cudaStream_t st = (cudaStream_t)stream;
#define SLOGP(MODEL, KT) k_scen_logp<MODEL, KT><<<grid, 128, 0, st>>>(p->d, SC, sc->S, B, d_theta, d_logp, d_logprior, d_loglik)
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
#define SSTAT(KT, ...) k_scen_stats<KT><<<(unsigned)n_scenarios, SS_THREADS, 0, st>>>(d_samples, d_logp, d_logprior, R, k, filter, coef, d_out)
    PQ_DISPATCH_K(pq_k_bucket(k), SSTAT, 0)
#undef SSTAT
    PQ_CUDA_OK(cudaGetLastError());
    return PEUQSE_OK;
}
