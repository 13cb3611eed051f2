// Lock-step Metropolis-Hastings: the chains of doMetropolisHastings (IP.py:2563-2754) advanced one step at a time by two small
// kernels, with the log-prior and log-likelihood of every chain's proposal supplied from OUTSIDE between them.  This is the form
// for what cannot live inside the persistent kernels of mh.cu: a simulator that is a torch callable, custom_logPrior /
// custom_logLikelihood (IP.py:2757-2759, 2854-2857) and responses with a full covariance matrix (the batched contraction of
// fullcov.cu).  Draws, proposal arithmetic, accept test, 'MAP_finding' adaptation, stored rows and the per-chain statistics
// block are those of k_mh (same Philox addressing, same operation order): for a built-in model both forms walk the same
// trajectory, which is how the tests pin this one.
//   step i of every chain:  k_mhs_propose  ->  <lp[c], ll[c] of d_prop[c][0..k)>  ->  k_mhs_accept
// The step index is a kernel argument, or (argument < 0) a per-chain counter in the state, so that one captured step can be replayed
// as a CUDA graph.
#include "common.cuh"
#include "eval.cuh"
#include "mh_common.cuh"

namespace {

// per-chain scratch behind the statistics block (which itself carries the current point, its logP and the running sums)
enum { MHS_CUR_LP = 0, MHS_CUR_LL, MHS_CUR_EXP, MHS_U, MHS_MOD, MHS_MF_DYN, MHS_MF_N, MHS_MF_SX, MHS_MF_SY, MHS_MF_SXY, MHS_MF_Y0,
       MHS_STEP /* the chain's next step (counter mode: step argument < 0, e.g. inside a replayed CUDA graph) */, MHS_TH0 /* [k] */ };
__host__ __device__ inline int mhs_state_doubles(int k) { return MHS_TH0 + k; }

struct StepOut {
    double* samples;
    double* logp;
    double* lprior;
    double* llik;
    double* deltas;
    double* us;
};

// ChainStats::add_row of mh.cu on the statistics block in global memory (upper triangle of S2; mirrored by k_mhs_finish)
__device__ __forceinline__ void mhs_add_row(double* sb, const double* th0, int k, long long row, const double* th, double lp, double explp) {
    if (lp > sb[PEUQSE_MH_STAT_MAPLOGP]) {
        sb[PEUQSE_MH_STAT_MAPLOGP] = lp;
        sb[PEUQSE_MH_STAT_MAPROW] = (double)row;
        for (int d = 0; d < k; ++d) sb[PEUQSE_MH_STAT_MAPTHETA + d] = th[d];
    }
    sb[PEUQSE_MH_STAT_SUMEXP] += explp;
    for (int d = 0; d < k; ++d) {
        const double dd = th[d] - th0[d];
        sb[PEUQSE_MH_STAT_S1(k) + d] += dd;
        for (int d2 = d; d2 < k; ++d2) {
            double* s2 = sb + PEUQSE_MH_STAT_S2(k) + d * k + d2;
            *s2 = fma(dd, th[d2] - th0[d2], *s2);
        }
    }
}

__global__ void __launch_bounds__(128) k_mhs_begin(const __grid_constant__ DevMH M, const double* __restrict__ theta0,
                                                   const double* __restrict__ lp0, const double* __restrict__ ll0,
                                                   double* __restrict__ state, double* __restrict__ stats, const StepOut out) {
    const long long c = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (c >= M.n_chains) return;
    const int k = M.k;
    double* sb = stats + c * PEUQSE_MH_STAT_DOUBLES(k);
    double* sx = state + c * mhs_state_doubles(k);
    double th[PQ_MAXK];
    for (int e = 0; e < PEUQSE_MH_STAT_DOUBLES(k); ++e) sb[e] = 0.0;
    sb[PEUQSE_MH_STAT_MAPLOGP] = pq_neg_inf();
    sb[PEUQSE_MH_STAT_MAPROW] = -1.0;
    for (int d = 0; d < k; ++d) {
        th[d] = theta0[c * k + d];
        sx[MHS_TH0 + d] = th[d];
        sb[PEUQSE_MH_STAT_LAST(k) + d] = th[d];
    }
    const double cur_lp = lp0[c], cur_ll = ll0[c];
    const double cur_sum = cur_ll + cur_lp;                      // what the accept test uses (IP.py:2641)
    // row 0: samples[0] = start, logP[0] = getLogP(start); log_priors[0] = log_likelihoods[0] = 0 (IP.py:2610-2625)
    const double lpr = combine_logp(cur_lp, cur_ll);
    if (out.samples)
        for (int d = 0; d < k; ++d) out.samples[c * k + d] = th[d];
    if (out.logp) out.logp[c] = lpr;
    if (out.lprior) out.lprior[c] = 0.0;
    if (out.llik) out.llik[c] = 0.0;
    if (M.burn_in <= 0) mhs_add_row(sb, sx + MHS_TH0, k, 0, th, lpr, exp(lpr));
    sb[PEUQSE_MH_STAT_LASTLOGP] = cur_sum;
    sx[MHS_CUR_LP] = cur_lp; sx[MHS_CUR_LL] = cur_ll; sx[MHS_CUR_EXP] = exp(cur_sum);
    sx[MHS_U] = 0.0; sx[MHS_MOD] = 1.0; sx[MHS_STEP] = 1.0;
    MapFind mf;
    mf.init();
    sx[MHS_MF_DYN] = mf.dyn; sx[MHS_MF_N] = mf.n; sx[MHS_MF_SX] = mf.sx; sx[MHS_MF_SY] = mf.sy; sx[MHS_MF_SXY] = mf.sxy; sx[MHS_MF_Y0] = mf.y0;
}

__global__ void __launch_bounds__(128) k_mhs_propose(const __grid_constant__ DevMH M, long long i, double* __restrict__ state,
                                                     const double* __restrict__ stats, double* __restrict__ prop, const StepOut out) {
    const long long c = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (c >= M.n_chains) return;
    const int k = M.k;
    const double* sb = stats + c * PEUQSE_MH_STAT_DOUBLES(k);
    double* sx = state + c * mhs_state_doubles(k);
    if (i < 0) i = (long long)sx[MHS_STEP];
    if (i >= M.n_steps) return;                                   // a replay beyond the run's last step does nothing
    double z[PQ_MAXK], delta[PQ_MAXK], u, um;
    draws_thread<PQ_MAXK>(M, (unsigned long long)(M.chain_offset + c), (unsigned long long)(M.first_step + i), z, u, um);
    make_delta<PQ_MAXK>(M, z, delta);
    const double mod = um + 0.5;                                  // IP.py:2633
    if (M.map_finding) {
        const double sc = sqrt(sx[MHS_MF_DYN] * mod);
        for (int d = 0; d < k; ++d) delta[d] = sc * delta[d];
    }
    for (int d = 0; d < k; ++d) prop[c * k + d] = sb[PEUQSE_MH_STAT_LAST(k) + d] + delta[d];   // IP.py:2630
    sx[MHS_U] = u;
    sx[MHS_MOD] = mod;
    if (out.deltas)
        for (int d = 0; d < k; ++d) out.deltas[((i - 1) * M.n_chains + c) * k + d] = delta[d];
    if (out.us) {
        if (M.map_finding) { out.us[((i - 1) * M.n_chains + c) * 2] = u; out.us[((i - 1) * M.n_chains + c) * 2 + 1] = mod; }
        else out.us[(i - 1) * M.n_chains + c] = u;
    }
}

__global__ void __launch_bounds__(128) k_mhs_accept(const __grid_constant__ DevMH M, long long i, const double* __restrict__ prop,
                                                    const double* __restrict__ lp_prop, const double* __restrict__ ll_prop,
                                                    double* __restrict__ state, double* __restrict__ stats, const StepOut out) {
    const long long c = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (c >= M.n_chains) return;
    const int k = M.k;
    const long long C = M.n_chains;
    double* sb = stats + c * PEUQSE_MH_STAT_DOUBLES(k);
    double* sx = state + c * mhs_state_doubles(k);
    const bool counter = i < 0;
    if (counter) i = (long long)sx[MHS_STEP];
    if (i >= M.n_steps) return;
    if (counter) sx[MHS_STEP] = (double)(i + 1);
    double cur_sum = sb[PEUQSE_MH_STAT_LASTLOGP];
    const double lp_p = lp_prop[c], ll_p = ll_prop[c];
    const double sum_p = ll_p + lp_p;
    double la = sum_p - cur_sum;                                  // IP.py:2641
    if (M.modulate != 0.0) la = (1.0 / M.modulate) * la;          // IP.py:2649
    const bool acc = la > log(sx[MHS_U]);                         // IP.py:2651-2652; NaN -> reject
    double th[PQ_MAXK];
    if (acc) {
        for (int d = 0; d < k; ++d) { th[d] = prop[c * k + d]; sb[PEUQSE_MH_STAT_LAST(k) + d] = th[d]; }
        cur_sum = sum_p;
        sx[MHS_CUR_LP] = lp_p; sx[MHS_CUR_LL] = ll_p; sx[MHS_CUR_EXP] = exp(cur_sum);
        sb[PEUQSE_MH_STAT_LASTLOGP] = cur_sum;
        sb[PEUQSE_MH_STAT_NACC] += 1.0;
    } else {
        for (int d = 0; d < k; ++d) th[d] = sb[PEUQSE_MH_STAT_LAST(k) + d];
    }
    if (M.map_finding) {
        MapFind mf;
        mf.dyn = sx[MHS_MF_DYN]; mf.n = sx[MHS_MF_N]; mf.sx = sx[MHS_MF_SX]; mf.sy = sx[MHS_MF_SY]; mf.sxy = sx[MHS_MF_SXY]; mf.y0 = sx[MHS_MF_Y0];
        mf.after_step(i, sx[MHS_MOD], sum_p);
        sx[MHS_MF_DYN] = mf.dyn; sx[MHS_MF_N] = mf.n; sx[MHS_MF_SX] = mf.sx; sx[MHS_MF_SY] = mf.sy; sx[MHS_MF_SXY] = mf.sxy; sx[MHS_MF_Y0] = mf.y0;
    }
    if (i % M.thin == 0) {                                        // rows are written every thin-th step
        const long long row = i / M.thin;
        if (out.samples)
            for (int d = 0; d < k; ++d) out.samples[(row * C + c) * k + d] = th[d];
        if (out.logp) out.logp[row * C + c] = cur_sum;          // IP.py:2661, 2673
        if (out.lprior) out.lprior[row * C + c] = sx[MHS_CUR_LP];
        if (out.llik) out.llik[row * C + c] = sx[MHS_CUR_LL];
    }
    if (i >= M.burn_in) mhs_add_row(sb, sx + MHS_TH0, k, i, th, cur_sum, sx[MHS_CUR_EXP]);
}

__global__ void __launch_bounds__(128) k_mhs_finish(const __grid_constant__ DevMH M, const double* __restrict__ state, double* __restrict__ stats) {
    const long long c = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (c >= M.n_chains) return;
    const int k = M.k;
    double* sb = stats + c * PEUQSE_MH_STAT_DOUBLES(k);
    for (int d = 0; d < k; ++d)
        for (int d2 = d + 1; d2 < k; ++d2) sb[PEUQSE_MH_STAT_S2(k) + d2 * k + d] = sb[PEUQSE_MH_STAT_S2(k) + d * k + d2];
    sb[PEUQSE_MH_STAT_DYN(k)] = state[c * mhs_state_doubles(k) + MHS_MF_DYN];
}

int mhs_spec(const peuqse_mh_spec* s, DevMH& M) {
    if (!s) return pq_set_error(PEUQSE_ERR_INVALID, "null argument");
    if (s->k < 1 || s->k > 14) return pq_set_error(PEUQSE_ERR_INVALID, "mh supports 1 <= k <= 14 (7 Philox blocks of normals per step)");
    if (s->n_chains < 1 || s->n_steps < 1 || s->thin < 1) return pq_set_error(PEUQSE_ERR_INVALID, "bad mh sizes");
    if ((unsigned long long)(s->first_step + s->n_steps) * 8ULL > 0xffffffffULL)
        return pq_set_error(PEUQSE_ERR_INVALID, "first_step + n_steps exceeds the 2^29 steps a Philox stream addresses");
    M.n_chains = s->n_chains; M.n_steps = s->n_steps; M.burn_in = s->burn_in; M.first_step = s->first_step;
    M.chain_offset = s->chain_offset; M.seed = s->seed; M.modulate = s->modulate; M.k = s->k; M.thin = s->thin; M.map_finding = s->map_finding ? 1 : 0;
    M.chol_diag = 1;
    for (int i = 0; i < PQ_MAXK * PQ_MAXK; ++i) M.chol[i] = 0.0;
    for (int d = 0; d < s->k; ++d)
        for (int e = 0; e < s->k; ++e) {
            const double v = s->step_chol[d * s->k + e];
            M.chol[d * s->k + e] = v;
            if (d != e && v != 0.0) M.chol_diag = 0;
        }
    return PEUQSE_OK;
}

inline unsigned mhs_blocks(const DevMH& M) { return (unsigned)((M.n_chains + 127) / 128); }

}  // namespace

extern "C" int64_t peuqse_mhs_state_doubles(int32_t k) { return mhs_state_doubles(k); }

extern "C" int peuqse_mhs_begin(const peuqse_mh_spec* s, const double* d_theta0, const double* d_logprior0, const double* d_loglik0,
                                double* d_state, double* d_stats, double* d_samples, double* d_logp, double* d_logprior,
                                double* d_loglik, void* stream) {
    DevMH M;
    if (int rc = mhs_spec(s, M)) return rc;
    if (!d_theta0 || !d_logprior0 || !d_loglik0 || !d_state || !d_stats) return pq_set_error(PEUQSE_ERR_INVALID, "null argument");
    const StepOut out{d_samples, d_logp, d_logprior, d_loglik, nullptr, nullptr};
    k_mhs_begin<<<mhs_blocks(M), 128, 0, (cudaStream_t)stream>>>(M, d_theta0, d_logprior0, d_loglik0, d_state, d_stats, out);
    PQ_CUDA_OK(cudaGetLastError());
    return PEUQSE_OK;
}

extern "C" int peuqse_mhs_propose(const peuqse_mh_spec* s, int64_t step, double* d_state, const double* d_stats, double* d_prop,
                                  double* d_deltas, double* d_us, void* stream) {
    DevMH M;
    if (int rc = mhs_spec(s, M)) return rc;
    if (!d_state || !d_stats || !d_prop || step == 0 || step >= s->n_steps) return pq_set_error(PEUQSE_ERR_INVALID, "bad argument");
    const StepOut out{nullptr, nullptr, nullptr, nullptr, d_deltas, d_us};
    k_mhs_propose<<<mhs_blocks(M), 128, 0, (cudaStream_t)stream>>>(M, step, d_state, d_stats, d_prop, out);
    PQ_CUDA_OK(cudaGetLastError());
    return PEUQSE_OK;
}

extern "C" int peuqse_mhs_accept(const peuqse_mh_spec* s, int64_t step, const double* d_prop, const double* d_logprior_prop,
                                 const double* d_loglik_prop, double* d_state, double* d_stats, double* d_samples, double* d_logp,
                                 double* d_logprior, double* d_loglik, void* stream) {
    DevMH M;
    if (int rc = mhs_spec(s, M)) return rc;
    if (!d_prop || !d_logprior_prop || !d_loglik_prop || !d_state || !d_stats || step == 0 || step >= s->n_steps)
        return pq_set_error(PEUQSE_ERR_INVALID, "bad argument");
    const StepOut out{d_samples, d_logp, d_logprior, d_loglik, nullptr, nullptr};
    k_mhs_accept<<<mhs_blocks(M), 128, 0, (cudaStream_t)stream>>>(M, step, d_prop, d_logprior_prop, d_loglik_prop, d_state, d_stats, out);
    PQ_CUDA_OK(cudaGetLastError());
    return PEUQSE_OK;
}

extern "C" int peuqse_mhs_finish(const peuqse_mh_spec* s, const double* d_state, double* d_stats, void* stream) {
    DevMH M;
    if (int rc = mhs_spec(s, M)) return rc;
    if (!d_state || !d_stats) return pq_set_error(PEUQSE_ERR_INVALID, "null argument");
    k_mhs_finish<<<mhs_blocks(M), 128, 0, (cudaStream_t)stream>>>(M, d_state, d_stats);
    PQ_CUDA_OK(cudaGetLastError());
    return PEUQSE_OK;
}
