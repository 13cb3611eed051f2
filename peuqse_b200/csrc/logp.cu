// Batched log-posterior from parameter vectors in memory (kernel K2 of SURVEY section 2):
// replaces getLogP / getLogPrior / getLogLikelihood called once per point from Python loops
// (IP.py:1647-1654, 2756-2794, 2850-2980).
#include "common.cuh"
#include "eval.cuh"
#include "launch.cuh"

// thread shape: one point per thread, theta rows read with stride k (each 128 B line is fully used by
// neighbouring threads and stays in L1), outputs written coalesced.
template <int MODEL, int KT>
__global__ void __launch_bounds__(256) k_logp_thread(const __grid_constant__ DevProblem P, const double* __restrict__ theta,
                                                     const double* __restrict__ sim, long long B,
                                                     double* __restrict__ logp, double* __restrict__ lprior,
                                                     double* __restrict__ llik, int use_smem) {
    extern __shared__ double smem[];
    const ObsView O = stage_obs(P, smem, use_smem);
    const long long stride = (long long)gridDim.x * blockDim.x;
    for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < B; i += stride) {
        double th[KT];
#pragma unroll
        for (int d = 0; d < KT; ++d) th[d] = (theta != nullptr && d < P.k) ? theta[i * P.k + d] : 0.0;
        const double lp = theta ? eval_prior<KT>(P, th) : 0.0;
        const double ll = eval_loglik_thread<MODEL, KT>(P, O, th, sim ? sim + i * P.n_total : nullptr);
        if (logp) logp[i] = combine_logp(lp, ll);
        if (lprior) lprior[i] = lp;
        if (llik) llik[i] = ll;
    }
}

// warp shape: one point per warp
template <int MODEL, int KT>
__global__ void __launch_bounds__(256) k_logp_warp(const __grid_constant__ DevProblem P, const double* __restrict__ theta,
                                                   const double* __restrict__ sim, long long B,
                                                   double* __restrict__ logp, double* __restrict__ lprior,
                                                   double* __restrict__ llik, int use_smem) {
    extern __shared__ double smem[];
    const ObsView O = stage_obs(P, smem, use_smem);
    const int lane = threadIdx.x & 31;
    const long long wstride = (long long)gridDim.x * (blockDim.x >> 5);
    for (long long i = (long long)blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5); i < B; i += wstride) {
        double th[KT];
#pragma unroll
        for (int d = 0; d < KT; ++d) th[d] = (theta != nullptr && d < P.k) ? theta[i * P.k + d] : 0.0;
        const double lp = theta ? eval_prior<KT>(P, th) : 0.0;
        const double ll = eval_loglik_warp<MODEL, KT>(P, O, th, sim ? sim + i * P.n_total : nullptr, lane);
        if (lane == 0) {
            if (logp) logp[i] = combine_logp(lp, ll);
            if (lprior) lprior[i] = lp;
            if (llik) llik[i] = ll;
        }
    }
}

// getSimulatedResponses (IP.py:2822-2848): d_sim[B][n_total] in natural order
template <int MODEL, int KT>
__global__ void __launch_bounds__(256) k_simulate(const __grid_constant__ DevProblem P, const double* __restrict__ theta,
                                                  long long B, double* __restrict__ sim, int use_smem) {
    extern __shared__ double smem[];
    const ObsView O = stage_obs(P, smem, use_smem);
    const long long total = B * P.n_total;
    const long long stride = (long long)gridDim.x * blockDim.x;
    for (long long e = (long long)blockIdx.x * blockDim.x + threadIdx.x; e < total; e += stride) {
        const long long i = e / P.n_total;
        const int j = (int)(e - i * P.n_total);
        double th[KT];
#pragma unroll
        for (int d = 0; d < KT; ++d) th[d] = (d < P.k) ? theta[i * P.k + d] : 0.0;
        sim[i * P.n_total + P.src[j]] = Sim<MODEL, KT>::at(P, O, th, nullptr, j);
    }
}

// getLogPrior alone (IP.py:2756-2794): for callers that supply the likelihood themselves (custom_logLikelihood, IP.py:2854-2857)
template <int KT>
__global__ void __launch_bounds__(256) k_logprior(const __grid_constant__ DevProblem P, const double* __restrict__ theta, long long B,
                                                  double* __restrict__ lprior) {
    const long long stride = (long long)gridDim.x * blockDim.x;
    for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < B; i += stride) {
        double th[KT];
#pragma unroll
        for (int d = 0; d < KT; ++d) th[d] = (d < P.k) ? theta[i * P.k + d] : 0.0;
        lprior[i] = eval_prior<KT>(P, th);
    }
}

namespace {

template <int MODEL, int KT>
int launch_logp(const peuqse_problem* p, const double* th, const double* sim, long long B, double* logp, double* lpr,
                double* llk, cudaStream_t st) {
    const bool warp_shape = p->d.n_total > PQ_THREAD_SHAPE_MAX_N;
    int use_smem;
    const size_t smem = pq_pick_smem(p, &use_smem);
    if (warp_shape) {
        auto kern = k_logp_warp<MODEL, KT>;
        PQ_CUDA_OK(pq_allow_smem((const void*)kern, smem));
        const long long blocks = (B + 7) / 8;
        const int grid = (int)std::min<long long>(blocks, (long long)p->sm_count * 8);
        kern<<<grid, 256, smem, st>>>(p->d, th, sim, B, logp, lpr, llk, use_smem);
    } else {
        auto kern = k_logp_thread<MODEL, KT>;
        PQ_CUDA_OK(pq_allow_smem((const void*)kern, smem));
        const long long blocks = (B + 255) / 256;
        const int grid = (int)std::min<long long>(blocks, (long long)p->sm_count * 8);
        kern<<<grid, 256, smem, st>>>(p->d, th, sim, B, logp, lpr, llk, use_smem);
    }
    PQ_CUDA_OK(cudaGetLastError());
    return PEUQSE_OK;
}

template <int MODEL, int KT>
int launch_sim(const peuqse_problem* p, const double* th, long long B, double* sim, cudaStream_t st) {
    int use_smem;
    const size_t smem = pq_pick_smem(p, &use_smem);
    auto kern = k_simulate<MODEL, KT>;
    PQ_CUDA_OK(pq_allow_smem((const void*)kern, smem));
    const long long blocks = (B * p->d.n_total + 255) / 256;
    const int grid = (int)std::min<long long>(blocks, (long long)p->sm_count * 8);
    kern<<<grid, 256, smem, st>>>(p->d, th, B, sim, use_smem);
    PQ_CUDA_OK(cudaGetLastError());
    return PEUQSE_OK;
}

}  // namespace

// prior + diagonal / per-value responses of caller-supplied simulated rows (full-covariance responses skipped)
int pq_logp_external(const peuqse_problem* p, const double* th, const double* sim, long long B, double* logp, double* lpr,
                     double* llk, cudaStream_t st) {
    const int kb = pq_k_bucket(p->d.k);
#define PQ_LOGP_EXT2(KT, ...) return launch_logp<PEUQSE_MODEL_EXTERNAL, KT>(p, th, sim, B, logp, lpr, llk, st)
    PQ_DISPATCH_K(kb, PQ_LOGP_EXT2, 0)
#undef PQ_LOGP_EXT2
    return PEUQSE_OK;
}

int pq_fullcov_logp(const peuqse_problem* p, const double* d_theta, const double* d_sim, long long B, double* d_logp,
                    double* d_logprior, double* d_loglik, cudaStream_t st);

extern "C" int peuqse_simulate_batch(const peuqse_problem* p, const double* d_theta, int64_t B, double* d_sim,
                                     void* stream) {
    if (!p || !d_theta || !d_sim || B < 0) return pq_set_error(PEUQSE_ERR_INVALID, "bad argument");
    if (B == 0) return PEUQSE_OK;
    cudaStream_t st = (cudaStream_t)stream;
    const int kb = pq_k_bucket(p->d.k);
    switch (p->d.model_id) {
        case PEUQSE_MODEL_QUADRATIC: return launch_sim<PEUQSE_MODEL_QUADRATIC, 2>(p, d_theta, B, d_sim, st);
        case PEUQSE_MODEL_TPD2SITE: return launch_sim<PEUQSE_MODEL_TPD2SITE, 6>(p, d_theta, B, d_sim, st);
        case PEUQSE_MODEL_LANGMUIR_LOG: return launch_sim<PEUQSE_MODEL_LANGMUIR_LOG, 2>(p, d_theta, B, d_sim, st);
        case PEUQSE_MODEL_LINEAR:
#define PQ_SIM_LIN(KT, ...) return launch_sim<PEUQSE_MODEL_LINEAR, KT>(p, d_theta, B, d_sim, st)
            PQ_DISPATCH_K(kb, PQ_SIM_LIN, 0)
#undef PQ_SIM_LIN
            break;
        default: break;
    }
    return pq_set_error(PEUQSE_ERR_INVALID, "simulate_batch needs a built-in model");
}

extern "C" int peuqse_logp_batch(const peuqse_problem* p, const double* d_theta, int64_t B, double* d_logp,
                                 double* d_logprior, double* d_loglik, void* stream) {
    if (!p || !d_theta || B < 0) return pq_set_error(PEUQSE_ERR_INVALID, "bad argument");
    if (B == 0) return PEUQSE_OK;
    cudaStream_t st = (cudaStream_t)stream;
    if (p->d.model_id == PEUQSE_MODEL_EXTERNAL)
        return pq_set_error(PEUQSE_ERR_INVALID, "logp_batch needs a built-in model; use peuqse_logp_from_sim");
    if (p->d.any_full) {
        bool all_full = true;
        for (int r = 0; r < p->d.n_resp; ++r) all_full = all_full && p->d.resp_mode[r] == PEUQSE_RESP_MVN_FULL;
        // linear model, every response dense: the residual kernels of the contraction simulate themselves (csrc/fullcov.cu)
        if (all_full && p->d.model_id == PEUQSE_MODEL_LINEAR) return pq_fullcov_logp(p, d_theta, nullptr, B, d_logp, d_logprior, d_loglik, st);
        // built-in simulate -> full-covariance contraction on the simulated block
        double* sim = nullptr;
        PQ_CUDA_OK(cudaMallocAsync((void**)&sim, (size_t)B * p->d.n_total * sizeof(double), st));
        int rc = peuqse_simulate_batch(p, d_theta, B, sim, stream);
        if (rc == PEUQSE_OK) rc = pq_fullcov_logp(p, d_theta, sim, B, d_logp, d_logprior, d_loglik, st);
        cudaFreeAsync(sim, st);
        return rc;
    }
    const int kb = pq_k_bucket(p->d.k);
    switch (p->d.model_id) {
        case PEUQSE_MODEL_QUADRATIC:
            return launch_logp<PEUQSE_MODEL_QUADRATIC, 2>(p, d_theta, nullptr, B, d_logp, d_logprior, d_loglik, st);
        case PEUQSE_MODEL_TPD2SITE:
            return launch_logp<PEUQSE_MODEL_TPD2SITE, 6>(p, d_theta, nullptr, B, d_logp, d_logprior, d_loglik, st);
        case PEUQSE_MODEL_LANGMUIR_LOG:
            return launch_logp<PEUQSE_MODEL_LANGMUIR_LOG, 2>(p, d_theta, nullptr, B, d_logp, d_logprior, d_loglik, st);
        case PEUQSE_MODEL_LINEAR:
#define PQ_LOGP_LIN(KT, ...) return launch_logp<PEUQSE_MODEL_LINEAR, KT>(p, d_theta, nullptr, B, d_logp, d_logprior, d_loglik, st)
            PQ_DISPATCH_K(kb, PQ_LOGP_LIN, 0)
#undef PQ_LOGP_LIN
            break;
        default: break;
    }
    return pq_set_error(PEUQSE_ERR_INVALID, "unknown model");
}

extern "C" int peuqse_logprior_batch(const peuqse_problem* p, const double* d_theta, int64_t B, double* d_logprior, void* stream) {
    if (!p || !d_theta || !d_logprior || B < 0) return pq_set_error(PEUQSE_ERR_INVALID, "bad argument");
    if (B == 0) return PEUQSE_OK;
    const int kb = pq_k_bucket(p->d.k);
    const int grid = (int)std::min<long long>((B + 255) / 256, (long long)p->sm_count * 8);
#define PQ_LPRIOR(KT, ...) k_logprior<KT><<<grid, 256, 0, (cudaStream_t)stream>>>(p->d, d_theta, B, d_logprior)
    PQ_DISPATCH_K(kb, PQ_LPRIOR, 0)
#undef PQ_LPRIOR
    PQ_CUDA_OK(cudaGetLastError());
    return PEUQSE_OK;
}

extern "C" int peuqse_logp_from_sim(const peuqse_problem* p, const double* d_theta, const double* d_sim, int64_t B,
                                    double* d_logp, double* d_logprior, double* d_loglik, void* stream) {
    if (!p || !d_sim || B < 0) return pq_set_error(PEUQSE_ERR_INVALID, "bad argument");
    if (!d_theta && (d_logp || d_logprior)) return pq_set_error(PEUQSE_ERR_INVALID, "logp/logprior need d_theta");
    if (B == 0) return PEUQSE_OK;
    cudaStream_t st = (cudaStream_t)stream;
    if (p->d.any_full) return pq_fullcov_logp(p, d_theta, d_sim, B, d_logp, d_logprior, d_loglik, st);
    const int kb = pq_k_bucket(p->d.k);
#define PQ_LOGP_EXT(KT, ...) return launch_logp<PEUQSE_MODEL_EXTERNAL, KT>(p, d_theta, d_sim, B, d_logp, d_logprior, d_loglik, st)
    PQ_DISPATCH_K(kb, PQ_LOGP_EXT, 0)
#undef PQ_LOGP_EXT
    return PEUQSE_OK;
}

// ---------------------------------------------------------------------------------------------
// Likelihood whose variances change with the point: model['responses_simulation_uncertainties'] as a FUNCTION of the parameters
// (IP.py:563-572, 2883-2896).  The simulation's standard deviations are squared and added to the observed variances
// (IP.py:2925-2927), so nothing of the density can be hoisted: per point and response
//     -0.5 (n_r log 2pi + sum_j log v_j + sum_j (y_j - sim_j)^2 / v_j),  v_j = var_obs_j + sd_j^2          (one logpdf, IP.py:2956)
// or, for a response longer than responses_observed_max_covmat_size, the sum of one scalar density per value with -inf replaced by
// -1E100 (IP.py:2961-2975).  A NaN among a point's simulated values makes its likelihood -inf (IP.py:2866-2882).
// One warp per point, lanes strided over the observations; sums in a fixed tree (deterministic, independent of the batch).
// ---------------------------------------------------------------------------------------------
#define PQ_SIMVAR_MAX_RESP 64
struct SimVarResp {
    int n_resp;
    int off[PQ_SIMVAR_MAX_RESP + 1];
    int per_value[PQ_SIMVAR_MAX_RESP];
};

__global__ void __launch_bounds__(128) k_loglik_simvar(const __grid_constant__ SimVarResp R, const double* __restrict__ sim,
                                                       const double* __restrict__ sd, const double* __restrict__ y,
                                                       const double* __restrict__ var_obs, long long B, double* __restrict__ out) {
    const long long b = ((long long)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    const int lane = threadIdx.x & 31;
    if (b >= B) return;
    const int n_total = R.off[R.n_resp];
    const double* s = sim + b * n_total;
    const double* e = sd + b * n_total;
    const double LOG2PI = 1.8378770664093453;
    double ll = 0.0;
    int bad = 0;
    for (int r = 0; r < R.n_resp; ++r) {
        const int lo = R.off[r], hi = R.off[r + 1];
        double a_log = 0.0, a_q = 0.0, a_pv = 0.0;
        for (int j = lo + lane; j < hi; j += 32) {
            const double sj = s[j], ej = e[j];
            bad |= (sj != sj);
            const double v = var_obs[j] + ej * ej;
            const double d = y[j] - sj;
            const double lg = log(v), q = d * d / v;
            if (R.per_value[r]) {
                double t = -0.5 * (LOG2PI + lg + q);
                if (pq_is_neg_inf(t)) t = -1E100;                      // IP.py:2973-2975
                a_pv += t;
            } else {
                a_log += lg;
                a_q += q;
            }
        }
        if (R.per_value[r]) ll += warp_sum(a_pv);
        else ll += -0.5 * ((double)(hi - lo) * LOG2PI + warp_sum(a_log) + warp_sum(a_q));
    }
    bad = __any_sync(0xffffffffu, bad);
    if (lane == 0) out[b] = bad ? pq_neg_inf() : ll;
}

extern "C" int peuqse_loglik_simvar(const double* d_sim, const double* d_sim_sd, const double* d_y, const double* d_var_obs,
                                    const int32_t* resp_n, const int32_t* per_value, int32_t n_resp, int64_t B, double* d_loglik,
                                    void* stream) {
    if (!d_sim || !d_sim_sd || !d_y || !d_var_obs || !resp_n || !per_value || !d_loglik || B < 0)
        return pq_set_error(PEUQSE_ERR_INVALID, "bad argument");
    if (n_resp < 1 || n_resp > PQ_SIMVAR_MAX_RESP) return pq_set_error(PEUQSE_ERR_INVALID, "1 <= n_resp <= 64");
    if (B == 0) return PEUQSE_OK;
    SimVarResp R;
    R.n_resp = n_resp;
    R.off[0] = 0;
    for (int r = 0; r < n_resp; ++r) {
        if (resp_n[r] < 1) return pq_set_error(PEUQSE_ERR_INVALID, "empty response");
        R.off[r + 1] = R.off[r] + resp_n[r];
        R.per_value[r] = per_value[r] ? 1 : 0;
    }
    const long long blocks = (B * 32 + 127) / 128;
    k_loglik_simvar<<<(unsigned)blocks, 128, 0, (cudaStream_t)stream>>>(R, d_sim, d_sim_sd, d_y, d_var_obs, B, d_loglik);
    PQ_CUDA_OK(cudaGetLastError());
    return PEUQSE_OK;
}

// ---------------------------------------------------------------------------------------------
// The simulated side of transform_responses for data_overcategory 'transient_kinetics' (IP.py:609-658, applied per evaluation at
// IP.py:2890): every pass replaces a response by its explicit-Euler integral over the abscissa, y[0] = 0, y[j+1] = y[j] + v[j] dt[j]
// (littleEulerGivenArray, IP.py:3501-3513).  A running sum in the reference's order -- one thread per (point, response), in place --
// so the transformed values are the reference's to the last bit (a parallel scan would reorder the additions).
// ---------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(128) k_transform_transient(const __grid_constant__ SimVarResp R, const double* __restrict__ dt,
                                                             long long B, double* __restrict__ sim) {
    const long long id = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (id >= B * R.n_resp) return;
    const long long b = id / R.n_resp;
    const int r = (int)(id - b * R.n_resp);
    const int lo = R.off[r], hi = R.off[r + 1];
    double* s = sim + b * R.off[R.n_resp];
    for (int pass = 0; pass < R.per_value[r]; ++pass) {           // per_value carries the number of passes here
        double acc = 0.0;
        for (int j = lo; j < hi; ++j) {
            const double v = s[j];
            s[j] = acc;
            acc = acc + v * dt[j];
        }
    }
}

extern "C" int peuqse_transform_transient(double* d_sim, const double* d_dt, const int32_t* resp_n, const int32_t* passes,
                                          int32_t n_resp, int64_t B, void* stream) {
    if (!d_sim || !d_dt || !resp_n || !passes || B < 0) return pq_set_error(PEUQSE_ERR_INVALID, "bad argument");
    if (n_resp < 1 || n_resp > PQ_SIMVAR_MAX_RESP) return pq_set_error(PEUQSE_ERR_INVALID, "1 <= n_resp <= 64");
    if (B == 0) return PEUQSE_OK;
    SimVarResp R;
    R.n_resp = n_resp;
    R.off[0] = 0;
    for (int r = 0; r < n_resp; ++r) {
        if (resp_n[r] < 1 || passes[r] < 0 || passes[r] > 2) return pq_set_error(PEUQSE_ERR_INVALID, "bad response description");
        R.off[r + 1] = R.off[r] + resp_n[r];
        R.per_value[r] = passes[r];
    }
    const long long blocks = (B * n_resp + 127) / 128;
    k_transform_transient<<<(unsigned)blocks, 128, 0, (cudaStream_t)stream>>>(R, d_dt, B, d_sim);
    PQ_CUDA_OK(cudaGetLastError());
    return PEUQSE_OK;
}
