// Ensemble slice sampling (kernel K4 of SURVEY section 2): replaces zeus.EnsembleSampler.run_mcmc, which the reference
// calls at IP.py:2495-2498 with logprob_fn=self.getLogP (one Python call per walker per stepping-out / shrink round).
// Algorithm: Karamanis & Beutler, "Ensemble slice sampling" (arXiv:2002.06212), differential move, as shipped in
// zeus-mcmc (third-party, not vendored in the reference: parity unpinned, validated statistically).
//
// One ensemble iteration = shuffle the walkers, split them in two halves, and update each half against the other.
// Within a half-sweep every active walker only READS the complementary half, so walkers are independent: one thread
// (small n) or one warp (large n) owns a walker and runs its whole slice update -- stepping out left / right and
// shrinking -- with its own Philox stream, evaluating the log-posterior in registers.  Kernel launches per iteration:
// shuffle, half 0, half 1, tune+record.  Nothing returns to the host during the run.
#include "common.cuh"
#include "eval.cuh"
#include "launch.cuh"
#include "philox.cuh"
#include "ess_common.cuh"

template <int MODEL, int KT, bool WARP>
__device__ __forceinline__ double ess_logp(const DevProblem& P, const ObsView& O, const double (&th)[KT], int lane) {
    const double lp = eval_prior<KT>(P, th);
    const double ll = WARP ? eval_loglik_warp<MODEL, KT>(P, O, th, nullptr, lane) : eval_loglik_thread<MODEL, KT>(P, O, th, nullptr);
    return combine_logp(lp, ll);      // getLogP (IP.py:1647-1654)
}

template <int MODEL, int KT, bool WARP>
__global__ void __launch_bounds__(WARP ? 256 : 128) k_ess_half(const __grid_constant__ DevProblem P, const DevEss E, long long it,
                                                               int half, const int* __restrict__ perm, double* __restrict__ X,
                                                               double* __restrict__ Z, EssState* __restrict__ st, int use_smem) {
    extern __shared__ double smem[];
    const ObsView O = stage_obs(P, smem, use_smem);
    const int lane = threadIdx.x & 31;
    const long long H = E.W / 2;
    const long long ia = WARP ? ((long long)blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5))
                              : ((long long)blockIdx.x * blockDim.x + threadIdx.x);
    if (ia >= H) return;
    const bool writer = WARP ? (lane == 0) : true;
    const int k = P.k;
    const long long w = perm[half * H + ia];
    const int* comp = perm + (1 - half) * H;
    const unsigned long long idx = ess_idx(E, it, half, w);
    const double mu = st->mu;

    // direction: differential move from two distinct walkers of the complementary half
    const Philox4 r0 = philox_block(E.seed, PQ_STREAM_ESS, idx, 0u);
    long long i1 = (long long)(philox_u53(r0.x, r0.y) * (double)H);
    long long i2 = (long long)(philox_u53(r0.z, r0.w) * (double)(H - 1));
    if (i1 >= H) i1 = H - 1;
    if (i2 >= H - 1) i2 = H - 2;
    if (i2 >= i1) ++i2;
    const long long c1 = comp[i1], c2 = comp[i2];
    double x0[KT], dir[KT], th[KT];
#pragma unroll
    for (int d = 0; d < KT; ++d) {
        x0[d] = (d < k) ? X[w * k + d] : 0.0;
        dir[d] = (d < k) ? 2.0 * mu * (X[c1 * k + d] - X[c2 * k + d]) : 0.0;
    }
    const Philox4 r1 = philox_block(E.seed, PQ_STREAM_ESS, idx, 1u);
    const double z0 = Z[w] - (-log(1.0 - philox_u53(r1.x, r1.y)));          // slice level: logP - Exp(1)
    double L = -philox_u53(r1.z, r1.w), R = L + 1.0;                          // initial interval of unit length around 0
    const Philox4 r2 = philox_block(E.seed, PQ_STREAM_ESS, idx, 2u);
    double J = floor(ESS_MAXSTEPS * philox_u53(r2.x, r2.y));
    double K = (ESS_MAXSTEPS - 1.0) - J;
    unsigned long long nexp = 0, ncon = 0, neval = 0;
    int err = 0;
    // stepping out, left then right
    for (int rounds = 0; J >= 1.0; ++rounds) {
#pragma unroll
        for (int d = 0; d < KT; ++d) th[d] = x0[d] + L * dir[d];
        const double zl = ess_logp<MODEL, KT, WARP>(P, O, th, lane);
        ++neval;
        if (!(z0 < zl)) break;
        L -= 1.0; J -= 1.0; ++nexp;
        if (rounds > E.max_rounds) { err = 1; break; }
    }
    for (int rounds = 0; K >= 1.0; ++rounds) {
#pragma unroll
        for (int d = 0; d < KT; ++d) th[d] = x0[d] + R * dir[d];
        const double zr = ess_logp<MODEL, KT, WARP>(P, O, th, lane);
        ++neval;
        if (!(z0 < zr)) break;
        R += 1.0; K -= 1.0; ++nexp;
        if (rounds > E.max_rounds) { err = 1; break; }
    }
    // shrinking
    double znew = Z[w];
    bool moved = false;
    for (unsigned int r = 0; !err; ++r) {
        const Philox4 rr = philox_block(E.seed, PQ_STREAM_ESS, idx, 3u + (r >> 1));
        const double u = (r & 1u) ? philox_u53(rr.z, rr.w) : philox_u53(rr.x, rr.y);
        const double wd = L + u * (R - L);
#pragma unroll
        for (int d = 0; d < KT; ++d) th[d] = x0[d] + wd * dir[d];
        const double zc = ess_logp<MODEL, KT, WARP>(P, O, th, lane);
        ++neval;
        if (z0 < zc) { znew = zc; moved = true; break; }
        if (wd < 0.0) L = wd;
        else if (wd > 0.0) R = wd;
        ++ncon;
        if ((int)r > E.max_rounds) { err = 2; break; }
    }
    if (writer) {
        if (moved) {
#pragma unroll
            for (int d = 0; d < KT; ++d)
                if (d < k) X[w * k + d] = th[d];
            Z[w] = znew;
        }
        atomicAdd(&st->nexp, nexp);
        atomicAdd(&st->ncon, ncon);
        atomicAdd(&st->neval, neval);
        if (err) atomicMax(&st->error, err);
    }
}

extern "C" int64_t peuqse_ess_work_bytes(int64_t n_walkers, int32_t k) {
    (void)k;
    return 256 + (int64_t)n_walkers * (int64_t)(sizeof(int) + sizeof(unsigned long long)) + 512;
}

int pq_ess_run_rounds(const peuqse_problem* p, const peuqse_ess_spec* s, double* d_x, double* d_logp_x, double* d_chain,
                      double* d_chain_logp, double* d_info, cudaStream_t st);

namespace {
template <int MODEL, int KT>
int launch_half(const peuqse_problem* p, const DevEss& E, long long it, int half, const int* perm, double* X, double* Z,
                EssState* st, cudaStream_t s) {
    int use_smem;
    const size_t smem = pq_pick_smem(p, &use_smem);
    const long long H = E.W / 2;
    if (p->d.n_total > PQ_THREAD_SHAPE_MAX_N) {
        auto kern = k_ess_half<MODEL, KT, true>;
        PQ_CUDA_OK(pq_allow_smem((const void*)kern, smem));
        kern<<<(unsigned)((H + 7) / 8), 256, smem, s>>>(p->d, E, it, half, perm, X, Z, st, use_smem);
    } else {
        auto kern = k_ess_half<MODEL, KT, false>;
        PQ_CUDA_OK(pq_allow_smem((const void*)kern, smem));
        kern<<<(unsigned)((H + 127) / 128), 128, smem, s>>>(p->d, E, it, half, perm, X, Z, st, use_smem);
    }
    PQ_CUDA_OK(cudaGetLastError());
    return PEUQSE_OK;
}

int dispatch_half(const peuqse_problem* p, const DevEss& E, long long it, int half, const int* perm, double* X, double* Z,
                  EssState* st, cudaStream_t s) {
    const int kb = pq_k_bucket(p->d.k);
    switch (p->d.model_id) {
        case PEUQSE_MODEL_QUADRATIC: return launch_half<PEUQSE_MODEL_QUADRATIC, 2>(p, E, it, half, perm, X, Z, st, s);
        case PEUQSE_MODEL_TPD2SITE: return launch_half<PEUQSE_MODEL_TPD2SITE, 6>(p, E, it, half, perm, X, Z, st, s);
        case PEUQSE_MODEL_LINEAR:
#define EH_L(KT, ...) return launch_half<PEUQSE_MODEL_LINEAR, KT>(p, E, it, half, perm, X, Z, st, s)
            PQ_DISPATCH_K(kb, EH_L, 0)
#undef EH_L
            break;
        default: break;
    }
    return pq_set_error(PEUQSE_ERR_INVALID, "ess needs a built-in model");
}
}  // namespace

extern "C" int peuqse_ess_run(const peuqse_problem* p, const peuqse_ess_spec* s, double* d_x, double* d_logp_x, double* d_chain,
                              double* d_chain_logp, double* d_info, void* d_work, void* stream) {
    if (!p || !s || !d_x || !d_logp_x) return pq_set_error(PEUQSE_ERR_INVALID, "null argument");
    if (s->k != p->d.k) return pq_set_error(PEUQSE_ERR_INVALID, "ess k != problem k");
    if (s->n_walkers < 4 || (s->n_walkers & 1)) return pq_set_error(PEUQSE_ERR_INVALID, "ess needs an even number (>= 4) of walkers");
    if (s->n_iters < 0) return pq_set_error(PEUQSE_ERR_INVALID, "negative n_iters");
    if (p->d.model_id == PEUQSE_MODEL_EXTERNAL)
        return pq_set_error(PEUQSE_ERR_INVALID, "peuqse_ess_run needs a built-in model; drive the rounds of a caller-simulated "
                                                "problem with peuqse_essr_* and peuqse_logp_from_sim");
    if (s->driver < 0 || s->driver > 2) return pq_set_error(PEUQSE_ERR_INVALID, "ess driver must be 0 (auto), 1 or 2");
    if (s->driver == 1 && p->d.any_full)
        return pq_set_error(PEUQSE_ERR_INVALID, "full-covariance responses need the round-synchronous driver (driver 0 or 2)");
    cudaStream_t st = (cudaStream_t)stream;
    if (s->driver == 2 || p->d.any_full) return pq_ess_run_rounds(p, s, d_x, d_logp_x, d_chain, d_chain_logp, d_info, st);
    if (!d_work) return pq_set_error(PEUQSE_ERR_INVALID, "null d_work");
    EssState* state = (EssState*)d_work;
    int* perm = (int*)((char*)d_work + 256);
    unsigned long long* keys = (unsigned long long*)((char*)d_work + 256 + ((s->n_walkers * sizeof(int) + 255) / 256) * 256);
    DevEss E;
    E.W = s->n_walkers; E.first_iter = s->first_iter; E.seed = s->seed; E.k = s->k;
    E.max_rounds = s->max_rounds > 0 ? s->max_rounds : 10000;
    k_ess_init<<<1, 1, 0, st>>>(state, s->mu0 > 0 ? s->mu0 : 1.0, s->tune);
    const unsigned wb = (unsigned)((E.W + 127) / 128);
    for (long long it = 0; it < s->n_iters; ++it) {
        k_ess_shuffle<<<wb, 128, 0, st>>>(E, it, perm, keys);
        k_ess_rank<<<wb, 128, 0, st>>>(E, keys, perm);
        for (int half = 0; half < 2; ++half) {
            int rc = dispatch_half(p, E, it, half, perm, d_x, d_logp_x, state, st);
            if (rc) return rc;
        }
        k_ess_tune<<<1, 1, 0, st>>>(state);
        if (d_chain) PQ_CUDA_OK(cudaMemcpyAsync(d_chain + it * E.W * E.k, d_x, (size_t)E.W * E.k * sizeof(double), cudaMemcpyDeviceToDevice, st));
        if (d_chain_logp) PQ_CUDA_OK(cudaMemcpyAsync(d_chain_logp + it * E.W, d_logp_x, (size_t)E.W * sizeof(double), cudaMemcpyDeviceToDevice, st));
    }
    if (d_info) k_ess_info<<<1, 1, 0, st>>>(state, d_info);
    PQ_CUDA_OK(cudaGetLastError());
    return PEUQSE_OK;
}
