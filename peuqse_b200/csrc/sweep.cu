// Multistart sweeps (kernels K1 + K6 of SURVEY section 2): grid / uniform point generation fused with
// simulate -> prior -> likelihood and the reductions the reference does afterwards on the host.
// Replaces doMultiStart / doListOfPermutationsSearch's Python loop (IP.py:844-956), the list built by
// CombinationGeneratorModule.combinationGenerator (CombinationGeneratorModule.py:40-81),
// convertPermutationsToSamples (IP.py:3463-3491) and the moment part of calculatePostBurnInStatistics
// (IP.py:1923-1931).
#include "common.cuh"
#include "eval.cuh"
#include "launch.cuh"

#include "sweep_common.cuh"

// specialised grid kernels (sweep_grid.cu)
long long pq_grid_aux_bytes(const peuqse_problem* p, const DevSweep& S, long long first, long long count);
int pq_grid_pass1(const peuqse_problem* p, const DevSweep& S, long long first, long long count, double* logp, void* aux_base,
                  peuqse_best* best, void* work, cudaStream_t st, bool* handled);
int pq_grid_pass2(const peuqse_problem* p, const DevSweep& S, long long first, long long count, const double* logp,
                  void* aux_base, const peuqse_best* best, double log_rel, double* acc, long long* kidx, double* klogp,
                  long long kcap, long long* kcount, void* work, cudaStream_t st, bool* handled);
int pq_grid_hist(const peuqse_problem* p, const DevSweep& S, long long first, long long count, const double* logp,
                 void* aux_base, const peuqse_best* best, double log_rel, const double* edges, int nbins,
                 unsigned long long* hist, cudaStream_t st, bool* handled, bool use_fused_sums);

#define PQ_SWEEP_BASE_BYTES ((long long)PQ_SWEEP_TICKET_BYTES + (long long)PQ_SWEEP_MAX_BLOCKS * PQ_ACC_MAX * (long long)sizeof(double))
static inline void* aux_ptr(const peuqse_problem* p, const DevSweep& S, long long first, long long count, void* d_work) {
    return (d_work && pq_grid_aux_bytes(p, S, first, count) > 0) ? (void*)((char*)d_work + PQ_SWEEP_BASE_BYTES) : nullptr;
}

__global__ void k_best_init(peuqse_best* b) {
    b->logp = pq_neg_inf();
    b->index = -1;
}

// ---------------------------------------------------------------------------------------------
// pass 1: logP of every point (+ optional store) and the first strict maximum
// ---------------------------------------------------------------------------------------------
template <int MODEL, int KT>
__global__ void __launch_bounds__(256) k_sweep_pass1(const __grid_constant__ DevProblem P, const __grid_constant__ DevSweep S,
                                                     long long first, long long count, double* __restrict__ logp_out,
                                                     peuqse_best* __restrict__ best_out, Best* __restrict__ partial,
                                                     unsigned int* __restrict__ ticket, int use_smem) {
    extern __shared__ double smem[];
    __shared__ Best s_best[8];
    __shared__ bool s_last;
    const ObsView O = stage_obs(P, smem, use_smem);
    Best best = best_none();
    const long long stride = (long long)gridDim.x * blockDim.x;
    for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < count; i += stride) {
        const long long g = first + i;
        double th[KT];
        sweep_point<KT>(S, (unsigned long long)g, th);
        const double lp = eval_prior<KT>(P, th);
        double v = lp;
        if (!pq_is_neg_inf(lp)) v = eval_loglik_thread<MODEL, KT>(P, O, th, nullptr) + lp;   // IP.py:1649-1653
        if (logp_out) __stcs(logp_out + i, v);
        if (v > best.v) { best.v = v; best.i = g; }            // strict '>', increasing g: first maximum
    }
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
            s_last = (t == gridDim.x - 1);
        }
    }
    __syncthreads();
    if (s_last && threadIdx.x < 32) {
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

template <int KT>
__global__ void __launch_bounds__(256) k_sweep_points(const __grid_constant__ DevSweep S, long long first, long long count,
                                                      double* __restrict__ theta) {
    const long long stride = (long long)gridDim.x * blockDim.x;
    for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < count; i += stride) {
        double th[KT];
        sweep_point<KT>(S, (unsigned long long)(first + i), th);
#pragma unroll
        for (int d = 0; d < KT; ++d)
            if (d < S.k) theta[i * S.k + d] = th[d];
    }
}

// ---------------------------------------------------------------------------------------------
// pass 2: weights + moments (+ compaction of the retained points)
// ---------------------------------------------------------------------------------------------
template <int KT>
struct Pass2Acc {
    double w, wexp, nk;
    double s1[KT];
    double s2[KT * (KT + 1) / 2];
    double mn[KT], mx[KT];
};

__device__ __forceinline__ double block_sum(double v, double* s_red) {
    v = warp_sum(v);
    __syncthreads();
    if ((threadIdx.x & 31) == 0) s_red[threadIdx.x >> 5] = v;
    __syncthreads();
    double t = 0.0;
    if (threadIdx.x < 32) {
        t = (threadIdx.x < (blockDim.x >> 5)) ? s_red[threadIdx.x] : 0.0;
        t = warp_sum(t);
    }
    return t;   // valid in warp 0
}
__device__ __forceinline__ double block_min(double v, double* s_red) {
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

template <int MODEL, int KT>
__global__ void __launch_bounds__(256) k_sweep_pass2(const __grid_constant__ DevProblem P, const __grid_constant__ DevSweep S,
                                                     long long first, long long count, const double* __restrict__ logp_in,
                                                     const peuqse_best* __restrict__ best_in, double log_rel,
                                                     double* __restrict__ partial, double* __restrict__ acc_out,
                                                     unsigned int* __restrict__ ticket, long long* __restrict__ keep_idx,
                                                     double* __restrict__ keep_logp, long long keep_cap,
                                                     unsigned long long* __restrict__ keep_count, int use_smem) {
    extern __shared__ double smem[];
    __shared__ double s_red[8];
    __shared__ bool s_last;
    const ObsView O = stage_obs(P, smem, use_smem);
    const long long best_index = best_in->index;
    if (best_index < 0) return;                                // no finite point: nothing is retained
    const double thr = best_in->logp + log_rel;               // maxLogP + np.log(relativeFilteringThreshold), IP.py:3477
    double shift[KT];                                          // moments about the MAP point (no cancellation)
    sweep_point<KT>(S, (unsigned long long)best_index, shift);
    Pass2Acc<KT> A;
    A.w = A.wexp = A.nk = 0.0;
#pragma unroll
    for (int d = 0; d < KT; ++d) { A.s1[d] = 0.0; A.mn[d] = pq_pos_inf(); A.mx[d] = pq_neg_inf(); }
#pragma unroll
    for (int e = 0; e < KT * (KT + 1) / 2; ++e) A.s2[e] = 0.0;

    const long long stride = (long long)gridDim.x * blockDim.x;
    for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < count; i += stride) {
        const long long g = first + i;
        double th[KT];
        double v;
        if (logp_in) {
            v = __ldcs(logp_in + i);
            if (!(v > thr)) continue;                          // IP.py:3481
            sweep_point<KT>(S, (unsigned long long)g, th);
        } else {
            sweep_point<KT>(S, (unsigned long long)g, th);
            const double lp = eval_prior<KT>(P, th);
            v = lp;
            if (!pq_is_neg_inf(lp)) v = eval_loglik_thread<MODEL, KT>(P, O, th, nullptr) + lp;
            if (!(v > thr)) continue;
        }
        const double reps = exp(v - thr);                       // IP.py:3485
        const double w = trunc(reps);                           // int(...), IP.py:3486
        A.w += w;
        A.wexp = fma(w, exp(v), A.wexp);
        A.nk += 1.0;
        int e = 0;
#pragma unroll
        for (int d = 0; d < KT; ++d) {
            if (d < S.k) {
                const double dd = th[d] - shift[d];
                const double wd = w * dd;
                A.s1[d] += wd;
#pragma unroll
                for (int d2 = d; d2 < KT; ++d2) {
                    if (d2 < S.k) A.s2[e] = fma(wd, th[d2] - shift[d2], A.s2[e]);
                    ++e;
                }
                A.mn[d] = fmin(A.mn[d], th[d]);
                A.mx[d] = fmax(A.mx[d], th[d]);
            } else {
                e += KT - d;
            }
        }
        if (keep_idx) {
            const unsigned long long slot = atomicAdd(keep_count, 1ULL);
            if ((long long)slot < keep_cap) {
                keep_idx[slot] = g;
                keep_logp[slot] = v;
            }
        }
    }
    // block reduction -> partial[blockIdx][...] laid out like d_acc
    const int k = S.k;
    double* mine = partial + (size_t)blockIdx.x * PQ_ACC_MAX;
    double t;
    t = block_sum(A.w, s_red);    if (threadIdx.x == 0) mine[PEUQSE_ACC_W] = t;
    t = block_sum(A.wexp, s_red); if (threadIdx.x == 0) mine[PEUQSE_ACC_WEXP] = t;
    t = block_sum(A.nk, s_red);   if (threadIdx.x == 0) mine[PEUQSE_ACC_NKEPT] = t;
    {
        int e = 0;
#pragma unroll
        for (int d = 0; d < KT; ++d) {
            if (d < k) {
                t = block_sum(A.s1[d], s_red);
                if (threadIdx.x == 0) mine[PEUQSE_ACC_S1 + d] = t;
                t = block_min(A.mn[d], s_red);
                if (threadIdx.x == 0) mine[PEUQSE_ACC_MIN(k) + d] = t;
                t = -block_min(-A.mx[d], s_red);
                if (threadIdx.x == 0) mine[PEUQSE_ACC_MAX(k) + d] = t;
            }
#pragma unroll
            for (int d2 = d; d2 < KT; ++d2) {
                if (d < k && d2 < k) {
                    t = block_sum(A.s2[e], s_red);
                    if (threadIdx.x == 0) {
                        mine[PEUQSE_ACC_S2(k) + d * k + d2] = t;
                        mine[PEUQSE_ACC_S2(k) + d2 * k + d] = t;
                    }
                }
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
        const int nacc = PEUQSE_ACC_DOUBLES(k);
        const int mn0 = PEUQSE_ACC_MIN(k), mx0 = PEUQSE_ACC_MAX(k);
        for (int a = threadIdx.x; a < nacc; a += blockDim.x) {
            double r = acc_out[a];
            for (int b = 0; b < (int)gridDim.x; ++b) {           // fixed block order: deterministic
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
// pass 3: weighted marginal histograms (np.histogram semantics on given edges)
// ---------------------------------------------------------------------------------------------
template <int MODEL, int KT>
__global__ void __launch_bounds__(256) k_sweep_hist(const __grid_constant__ DevProblem P, const __grid_constant__ DevSweep S,
                                                    long long first, long long count, const double* __restrict__ logp_in,
                                                    const peuqse_best* __restrict__ best_in, double log_rel,
                                                    const double* __restrict__ edges, int nbins,
                                                    unsigned long long* __restrict__ hist, int use_smem) {
    extern __shared__ double smem[];
    const ObsView O = stage_obs(P, smem, use_smem);
    // block-private histogram after the staged constants
    unsigned long long* sh = (unsigned long long*)(smem + (use_smem ? pq_obs_doubles(P) : 0));
    const int k = S.k;
    if (best_in->index < 0) return;
    const double thr = best_in->logp + log_rel;
    for (int a = threadIdx.x; a < k * nbins; a += blockDim.x) sh[a] = 0ULL;
    __syncthreads();
    const long long stride = (long long)gridDim.x * blockDim.x;
    for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < count; i += stride) {
        const long long g = first + i;
        double th[KT];
        double v;
        if (logp_in) {
            v = __ldcs(logp_in + i);
            if (!(v > thr)) continue;
            sweep_point<KT>(S, (unsigned long long)g, th);
        } else {
            sweep_point<KT>(S, (unsigned long long)g, th);
            const double lp = eval_prior<KT>(P, th);
            v = lp;
            if (!pq_is_neg_inf(lp)) v = eval_loglik_thread<MODEL, KT>(P, O, th, nullptr) + lp;
            if (!(v > thr)) continue;
        }
        const unsigned long long w = (unsigned long long)trunc(exp(v - thr));
#pragma unroll
        for (int d = 0; d < KT; ++d) {
            if (d < k) {
                const double* ed = edges + d * (nbins + 1);
                const double lo = ed[0], hi = ed[nbins];
                const double x = th[d];
                if (x >= lo && x <= hi) {
                    // numpy/lib/_histograms_impl.py: index from the normalised position, then corrected
                    // against the actual edges; the last bin is closed on the right
                    int b = (int)(((x - lo) / (hi - lo)) * (double)nbins);
                    if (b >= nbins) b = nbins - 1;
                    if (b < 0) b = 0;
                    if (x < ed[b]) --b;
                    else if (b != nbins - 1 && x >= ed[b + 1]) ++b;
                    atomicAdd(sh + d * nbins + b, w);
                }
            }
        }
    }
    __syncthreads();
    for (int a = threadIdx.x; a < k * nbins; a += blockDim.x)
        if (sh[a]) atomicAdd(hist + a, sh[a]);
}

// ---------------------------------------------------------------------------------------------
// launchers
// ---------------------------------------------------------------------------------------------
namespace {

inline int sweep_grid(const peuqse_problem* p, long long count) {
    const long long blocks = (count + 255) / 256;
    return (int)std::max<long long>(1, std::min<long long>(blocks, std::min<long long>((long long)p->sm_count * 8, PQ_SWEEP_MAX_BLOCKS)));
}

template <int MODEL, int KT>
int launch_pass1(const peuqse_problem* p, const DevSweep& S, long long first, long long count, double* logp,
                 peuqse_best* best, void* work, cudaStream_t st) {
    int use_smem;
    const size_t smem = pq_pick_smem(p, &use_smem);
    auto kern = k_sweep_pass1<MODEL, KT>;
    PQ_CUDA_OK(pq_allow_smem((const void*)kern, smem));
    unsigned int* ticket = (unsigned int*)work;
    Best* partial = (Best*)((char*)work + PQ_SWEEP_TICKET_BYTES);
    kern<<<sweep_grid(p, count), 256, smem, st>>>(p->d, S, first, count, logp, best, partial, ticket, use_smem);
    PQ_CUDA_OK(cudaGetLastError());
    return PEUQSE_OK;
}

template <int MODEL, int KT>
int launch_pass2(const peuqse_problem* p, const DevSweep& S, long long first, long long count, const double* logp,
                 const peuqse_best* best, double log_rel, double* acc, long long* kidx, double* klogp, long long kcap, long long* kcount, void* work,
                 cudaStream_t st) {
    int use_smem;
    const size_t smem = pq_pick_smem(p, &use_smem);
    auto kern = k_sweep_pass2<MODEL, KT>;
    PQ_CUDA_OK(pq_allow_smem((const void*)kern, smem));
    unsigned int* ticket = (unsigned int*)work;
    double* partial = (double*)((char*)work + PQ_SWEEP_TICKET_BYTES);
    kern<<<sweep_grid(p, count), 256, smem, st>>>(p->d, S, first, count, logp, best, log_rel, partial, acc, ticket, kidx, klogp, kcap,
                                                  (unsigned long long*)kcount, use_smem);
    PQ_CUDA_OK(cudaGetLastError());
    return PEUQSE_OK;
}

template <int MODEL, int KT>
int launch_hist(const peuqse_problem* p, const DevSweep& S, long long first, long long count, const double* logp,
                const peuqse_best* best, double log_rel, const double* edges, int nbins, unsigned long long* hist, cudaStream_t st) {
    int use_smem;
    size_t smem = pq_pick_smem(p, &use_smem);
    smem += (size_t)S.k * nbins * sizeof(unsigned long long);
    auto kern = k_sweep_hist<MODEL, KT>;
    PQ_CUDA_OK(pq_allow_smem((const void*)kern, smem));
    kern<<<sweep_grid(p, count), 256, smem, st>>>(p->d, S, first, count, logp, best, log_rel, edges, nbins, hist, use_smem);
    PQ_CUDA_OK(cudaGetLastError());
    return PEUQSE_OK;
}

}  // namespace

extern "C" int64_t peuqse_sweep_work_bytes(const peuqse_problem* p, const peuqse_sweep_spec* spec, int64_t first, int64_t count) {
    long long aux = 0;
    if (p && spec) {
        DevSweep S;
        long long total;
        if (build_sweep(spec, &S, &total) == PEUQSE_OK) aux = pq_grid_aux_bytes(p, S, first, count);
    }
    return PQ_SWEEP_BASE_BYTES + aux;
}

#define PQ_SWEEP_DISPATCH(CALL_Q, CALL_T, CALL_L, CALL_E)                                         \
    switch (p->d.model_id) {                                                                       \
        case PEUQSE_MODEL_QUADRATIC: return CALL_Q;                                                \
        case PEUQSE_MODEL_TPD2SITE: return CALL_T;                                                 \
        case PEUQSE_MODEL_LINEAR: {                                                                \
            const int kb = pq_k_bucket(p->d.k);                                                    \
            PQ_DISPATCH_K(kb, CALL_L, 0)                                                           \
        } break;                                                                                   \
        case PEUQSE_MODEL_EXTERNAL: {                                                              \
            const int kb = pq_k_bucket(p->d.k);                                                    \
            PQ_DISPATCH_K(kb, CALL_E, 0)                                                           \
        } break;                                                                                   \
        default: break;                                                                            \
    }

static int check_sweep_args(const peuqse_problem* p, const peuqse_sweep_spec* spec, long long first, long long count) {
    if (!p || !spec) return pq_set_error(PEUQSE_ERR_INVALID, "null argument");
    if (spec->k != p->d.k) return pq_set_error(PEUQSE_ERR_INVALID, "sweep k != problem k");
    if (first < 0 || count < 0) return pq_set_error(PEUQSE_ERR_INVALID, "negative range");
    if (p->d.any_full) return pq_set_error(PEUQSE_ERR_INVALID, "sweeps do not support full-covariance responses; use logp_batch on sweep_points");
    return PEUQSE_OK;
}

extern "C" int peuqse_sweep_pass1(const peuqse_problem* p, const peuqse_sweep_spec* spec, int64_t first, int64_t count,
                                  double* d_logp, peuqse_best* d_best, void* d_work, void* stream) {
    int rc = check_sweep_args(p, spec, first, count);
    if (rc) return rc;
    if (!d_best || !d_work) return pq_set_error(PEUQSE_ERR_INVALID, "d_best and d_work are required");
    if (p->d.model_id == PEUQSE_MODEL_EXTERNAL) return pq_set_error(PEUQSE_ERR_INVALID, "pass 1 needs a built-in model");
    DevSweep S;
    long long total;
    rc = build_sweep(spec, &S, &total);
    if (rc) return rc;
    if (first + count > total) return pq_set_error(PEUQSE_ERR_INVALID, "range exceeds the number of sweep points");
    cudaStream_t st = (cudaStream_t)stream;
    {
        bool handled = false;
        rc = pq_grid_pass1(p, S, first, count, d_logp, aux_ptr(p, S, first, count, d_work), d_best, d_work, st, &handled);
        if (rc || handled) return rc;
    }
    PQ_CUDA_OK(cudaMemsetAsync(d_work, 0, PQ_SWEEP_TICKET_BYTES, st));
    k_best_init<<<1, 1, 0, st>>>(d_best);
    if (count == 0) return PEUQSE_OK;
#define P1_L(KT, ...) return launch_pass1<PEUQSE_MODEL_LINEAR, KT>(p, S, first, count, d_logp, d_best, d_work, st)
#define P1_E(KT, ...) return pq_set_error(PEUQSE_ERR_INVALID, "pass 1 needs a built-in model")
    PQ_SWEEP_DISPATCH((launch_pass1<PEUQSE_MODEL_QUADRATIC, 2>(p, S, first, count, d_logp, d_best, d_work, st)),
                      (launch_pass1<PEUQSE_MODEL_TPD2SITE, 6>(p, S, first, count, d_logp, d_best, d_work, st)), P1_L, P1_E)
#undef P1_L
#undef P1_E
    return pq_set_error(PEUQSE_ERR_INVALID, "unknown model");
}

extern "C" int peuqse_sweep_points(const peuqse_sweep_spec* spec, int64_t first, int64_t count, double* d_theta, void* stream) {
    DevSweep S;
    long long total;
    int rc = build_sweep(spec, &S, &total);
    if (rc) return rc;
    if (first < 0 || count < 0 || first + count > total) return pq_set_error(PEUQSE_ERR_INVALID, "bad range");
    if (count == 0) return PEUQSE_OK;
    cudaStream_t st = (cudaStream_t)stream;
    const int grid = (int)std::min<long long>((count + 255) / 256, 148 * 8);
    const int kb = pq_k_bucket(spec->k);
#define PTS(KT, ...) k_sweep_points<KT><<<grid, 256, 0, st>>>(S, first, count, d_theta)
    PQ_DISPATCH_K(kb, PTS, 0)
#undef PTS
    PQ_CUDA_OK(cudaGetLastError());
    return PEUQSE_OK;
}

extern "C" int peuqse_sweep_pass2(const peuqse_problem* p, const peuqse_sweep_spec* spec, int64_t first, int64_t count,
                                  const double* d_logp, const peuqse_best* d_best, double log_rel_threshold, double* d_acc,
                                  int64_t* d_keep_index,
                                  double* d_keep_logp, int64_t keep_capacity, int64_t* d_keep_count, void* d_work,
                                  void* stream) {
    int rc = check_sweep_args(p, spec, first, count);
    if (rc) return rc;
    if (!d_acc || !d_work || !d_best) return pq_set_error(PEUQSE_ERR_INVALID, "d_acc, d_best and d_work are required");
    if (d_keep_index && (!d_keep_logp || !d_keep_count)) return pq_set_error(PEUQSE_ERR_INVALID, "incomplete keep buffers");
    if (p->d.model_id == PEUQSE_MODEL_EXTERNAL && !d_logp) return pq_set_error(PEUQSE_ERR_INVALID, "external model needs d_logp");
    DevSweep S;
    long long total;
    rc = build_sweep(spec, &S, &total);
    if (rc) return rc;
    if (first + count > total) return pq_set_error(PEUQSE_ERR_INVALID, "range exceeds the number of sweep points");
    cudaStream_t st = (cudaStream_t)stream;
    {
        bool handled = false;
        rc = pq_grid_pass2(p, S, first, count, d_logp, aux_ptr(p, S, first, count, d_work), d_best, log_rel_threshold, d_acc,
                           (long long*)d_keep_index, d_keep_logp, keep_capacity, (long long*)d_keep_count, d_work, st, &handled);
        if (rc || handled) return rc;
    }
    PQ_CUDA_OK(cudaMemsetAsync(d_work, 0, PQ_SWEEP_TICKET_BYTES, st));
    if (count == 0) return PEUQSE_OK;
#define P2_ARGS p, S, first, count, d_logp, d_best, log_rel_threshold, d_acc, (long long*)d_keep_index, d_keep_logp, keep_capacity, (long long*)d_keep_count, d_work, st
#define P2_L(KT, ...) return launch_pass2<PEUQSE_MODEL_LINEAR, KT>(P2_ARGS)
#define P2_E(KT, ...) return launch_pass2<PEUQSE_MODEL_EXTERNAL, KT>(P2_ARGS)
    PQ_SWEEP_DISPATCH((launch_pass2<PEUQSE_MODEL_QUADRATIC, 2>(P2_ARGS)), (launch_pass2<PEUQSE_MODEL_TPD2SITE, 6>(P2_ARGS)), P2_L, P2_E)
#undef P2_L
#undef P2_E
    return pq_set_error(PEUQSE_ERR_INVALID, "unknown model");
}

extern "C" int peuqse_sweep_hist(const peuqse_problem* p, const peuqse_sweep_spec* spec, int64_t first, int64_t count,
                                 const double* d_logp, const peuqse_best* d_best, double log_rel_threshold,
                                 const double* d_edges, int32_t nbins, uint64_t* d_hist, void* d_work,
                                 int32_t reuse_pass2_sums, void* stream) {
    int rc = check_sweep_args(p, spec, first, count);
    if (rc) return rc;
    if (!d_edges || !d_hist || !d_best || nbins < 1 || nbins > 4096) return pq_set_error(PEUQSE_ERR_INVALID, "bad histogram arguments");
    if (p->d.model_id == PEUQSE_MODEL_EXTERNAL && !d_logp) return pq_set_error(PEUQSE_ERR_INVALID, "external model needs d_logp");
    DevSweep S;
    long long total;
    rc = build_sweep(spec, &S, &total);
    if (rc) return rc;
    if (first + count > total) return pq_set_error(PEUQSE_ERR_INVALID, "range exceeds the number of sweep points");
    cudaStream_t st = (cudaStream_t)stream;
    {
        bool handled = false;
        rc = pq_grid_hist(p, S, first, count, d_logp, aux_ptr(p, S, first, count, d_work), d_best, log_rel_threshold, d_edges,
                          nbins, (unsigned long long*)d_hist, st, &handled, reuse_pass2_sums != 0);
        if (rc || handled) return rc;
    }
    if (count == 0) return PEUQSE_OK;
#define P3_ARGS p, S, first, count, d_logp, d_best, log_rel_threshold, d_edges, nbins, (unsigned long long*)d_hist, st
#define P3_L(KT, ...) return launch_hist<PEUQSE_MODEL_LINEAR, KT>(P3_ARGS)
#define P3_E(KT, ...) return launch_hist<PEUQSE_MODEL_EXTERNAL, KT>(P3_ARGS)
    PQ_SWEEP_DISPATCH((launch_hist<PEUQSE_MODEL_QUADRATIC, 2>(P3_ARGS)), (launch_hist<PEUQSE_MODEL_TPD2SITE, 6>(P3_ARGS)), P3_L, P3_E)
#undef P3_L
#undef P3_E
    return pq_set_error(PEUQSE_ERR_INVALID, "unknown model");
}

// np.linspace(min_d, max_d, nbins+1) from the pass-2 accumulator (np.histogram's edges for range=None):
// y = arange(0, num)*step + start with step = (stop-start)/div, last element := stop (numpy/_core/function_base.py)
__global__ void k_sweep_edges(const double* __restrict__ acc, int k, int nbins, double* __restrict__ edges,
                              unsigned long long* __restrict__ hist_to_zero) {
    const int t = blockIdx.x * blockDim.x + threadIdx.x;
    if (hist_to_zero && t < k * nbins) hist_to_zero[t] = 0ULL;
    if (t >= k * (nbins + 1)) return;
    const int d = t / (nbins + 1), i = t - d * (nbins + 1);
    double lo = acc[PEUQSE_ACC_MIN(k) + d], hi = acc[PEUQSE_ACC_MAX(k) + d];
    if (!(hi > lo)) { lo = lo - 0.5; hi = hi + 0.5; }          // np.histogram widens an empty range by +-0.5
    const double step = (hi - lo) / (double)nbins;
    edges[t] = (i == nbins) ? hi : ((double)i * step + lo);
}

extern "C" int peuqse_sweep_edges(const double* d_acc, int32_t k, int32_t nbins, double* d_edges, uint64_t* d_hist_to_zero,
                                  void* stream) {
    if (!d_acc || !d_edges || k < 1 || k > PQ_MAXK || nbins < 1) return pq_set_error(PEUQSE_ERR_INVALID, "bad edges arguments");
    const int n = k * (nbins + 1);
    k_sweep_edges<<<(n + 127) / 128, 128, 0, (cudaStream_t)stream>>>(d_acc, k, nbins, d_edges, (unsigned long long*)d_hist_to_zero);
    PQ_CUDA_OK(cudaGetLastError());
    return PEUQSE_OK;
}

// theta of the best point, decoded on the device (keeps the whole multistart free of host round trips)
template <int KT>
__global__ void k_best_point(const __grid_constant__ DevSweep S, const peuqse_best* __restrict__ best, double* __restrict__ theta) {
    double th[KT];
    const long long g = best->index;
    if (g >= 0) sweep_point<KT>(S, (unsigned long long)g, th);
#pragma unroll
    for (int d = 0; d < KT; ++d)
        if (d < S.k) theta[d] = g >= 0 ? th[d] : __longlong_as_double(0x7ff8000000000000LL);
}

extern "C" int peuqse_sweep_best_point(const peuqse_sweep_spec* spec, const peuqse_best* d_best, double* d_theta, void* stream) {
    DevSweep S;
    long long total;
    int rc = build_sweep(spec, &S, &total);
    if (rc) return rc;
    if (!d_best || !d_theta) return pq_set_error(PEUQSE_ERR_INVALID, "null argument");
    cudaStream_t st = (cudaStream_t)stream;
    const int kb = pq_k_bucket(spec->k);
#define BPT(KT, ...) k_best_point<KT><<<1, 1, 0, st>>>(S, d_best, d_theta)
    PQ_DISPATCH_K(kb, BPT, 0)
#undef BPT
    PQ_CUDA_OK(cudaGetLastError());
    return PEUQSE_OK;
}

// ---------------------------------------------------------------------------------------------
// multi-GPU merges: the all-gathered per-rank records are combined by one tiny kernel each
// ---------------------------------------------------------------------------------------------
__global__ void k_merge_best(const peuqse_best* __restrict__ all, int n, peuqse_best* __restrict__ out) {
    Best b = best_none();
    for (int r = threadIdx.x; r < n; r += 32) {
        Best o;
        o.v = all[r].logp;
        o.i = all[r].index < 0 ? 0x7fffffffffffffffLL : all[r].index;
        if (all[r].index < 0) o.v = pq_neg_inf();
        b = best_merge(b, o);
    }
    b = warp_best(b);
    if (threadIdx.x == 0) {
        out->logp = b.v;
        out->index = (b.i == 0x7fffffffffffffffLL) ? -1 : b.i;
    }
}

__global__ void k_merge_acc(const double* __restrict__ all, int n, int k, double* __restrict__ out) {
    const int nacc = PEUQSE_ACC_DOUBLES(k), mn0 = PEUQSE_ACC_MIN(k), mx0 = PEUQSE_ACC_MAX(k);
    for (int a = threadIdx.x; a < nacc; a += blockDim.x) {
        double r = all[a];
        for (int q = 1; q < n; ++q) {                      // fixed rank order: identical result on every rank
            const double x = all[(size_t)q * nacc + a];
            if (a >= mx0) r = fmax(r, x);
            else if (a >= mn0) r = fmin(r, x);
            else r += x;
        }
        out[a] = r;
    }
}

extern "C" int peuqse_merge_best(const peuqse_best* d_all, int32_t n, peuqse_best* d_out, void* stream) {
    if (!d_all || !d_out || n < 1) return pq_set_error(PEUQSE_ERR_INVALID, "bad merge_best arguments");
    k_merge_best<<<1, 32, 0, (cudaStream_t)stream>>>(d_all, n, d_out);
    PQ_CUDA_OK(cudaGetLastError());
    return PEUQSE_OK;
}

extern "C" int peuqse_merge_acc(const double* d_all, int32_t n, int32_t k, double* d_out, void* stream) {
    if (!d_all || !d_out || n < 1 || k < 1 || k > PQ_MAXK) return pq_set_error(PEUQSE_ERR_INVALID, "bad merge_acc arguments");
    k_merge_acc<<<1, 128, 0, (cudaStream_t)stream>>>(d_all, n, k, d_out);
    PQ_CUDA_OK(cudaGetLastError());
    return PEUQSE_OK;
}

// raw Philox blocks (tests): d_out[count][4]
__global__ void k_philox_blocks(unsigned long long seed, uint32_t stream_id, long long first, long long count,
                                uint32_t block, uint32_t* __restrict__ out) {
    const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= count) return;
    const Philox4 r = philox_block(seed, stream_id, (unsigned long long)(first + i), block);
    out[4 * i + 0] = r.x; out[4 * i + 1] = r.y; out[4 * i + 2] = r.z; out[4 * i + 3] = r.w;
}

extern "C" int peuqse_philox_blocks(uint64_t seed, uint32_t rng_stream, int64_t first, int64_t count, uint32_t block,
                                    uint32_t* d_out, void* stream) {
    if (count <= 0) return PEUQSE_OK;
    k_philox_blocks<<<(unsigned)((count + 255) / 256), 256, 0, (cudaStream_t)stream>>>(seed, rng_stream, first, count, block, d_out);
    PQ_CUDA_OK(cudaGetLastError());
    return PEUQSE_OK;
}
