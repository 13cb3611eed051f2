// Multistart sweeps (kernels K1 + K6 of SURVEY section 2): grid / uniform point generation fused with
// simulate -> prior -> likelihood and the reductions the reference does afterwards on the host.
// Replaces doMultiStart / doListOfPermutationsSearch's Python loop (IP.py:844-956), the list built by
// CombinationGeneratorModule.combinationGenerator (CombinationGeneratorModule.py:40-81),
// convertPermutationsToSamples (IP.py:3463-3491) and the moment part of calculatePostBurnInStatistics
// (IP.py:1923-1931).
#include <cstdlib>
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

// shortlist helpers (sweep_common.cuh): monotone double <-> key, append
__device__ __forceinline__ unsigned long long short_double_to_key(double v) {
    const long long b = __double_as_longlong(v);
    return b < 0 ? ~(unsigned long long)b : ((unsigned long long)b | 0x8000000000000000ULL);
}
__device__ __forceinline__ double short_key_to_double(unsigned long long key) {
    return __longlong_as_double((long long)((key & 0x8000000000000000ULL) ? (key & 0x7fffffffffffffffULL) : ~key));
}
__device__ __forceinline__ void short_push(ShortHdr* sh, long long* c_idx, double* c_logp, long long g, double v) {
    const unsigned long long slot = atomicAdd(&sh->count, 1ULL);
    if (slot < (unsigned long long)PQ_SHORT_CAP) {
        c_idx[slot] = g;
        c_logp[slot] = v;
    }
}
__device__ __forceinline__ bool short_usable(const ShortHdr* sh, int margin_ok) {
    return sh && margin_ok && sh->valid == 1u && sh->count <= (unsigned long long)PQ_SHORT_CAP;
}

// ---------------------------------------------------------------------------------------------
// pass 1 of a sweep whose logP is NOT kept (config C5: 1e10 uniform points), linear model, one diagonal-covariance
// response of 8 NF observations: the simulation Theta[32 points, k] Phi^T[k, n] of a warp's 32 points runs on the fp64
// TENSOR pipe (mma.sync.m8n8k4.f64, 4 row fragments x NF column fragments x ceil(k/4) steps), which works beside the fp64
// pipe that keeps the point generation (Philox -> theta), the prior and the residual / square / sum of the likelihood:
// the scalar kernel below spends 2 k n of its ~1.6 k fp64 instructions per point on that product.  Per lane: its own point's
// theta and prior; the B fragments of Phi^T and {y w, w} live in shared memory, the warp's thetas pass through a shared-memory
// tile to become A fragments; a lane ends up with 2 NF simulated values of each of 4 points, squares and sums its residuals,
// the 4 lanes of a point combine, the point's owner picks the sum up.  Summation order differs from the scalar kernel's
// (the reference's): values agree to ~1e-13 relative, inside the 1e-12 budget -- which is why only the no-store sweep, whose
// per-point values never reach the caller, takes this path; sweeps that return logP keep the bit-exact scalar order.
// ---------------------------------------------------------------------------------------------
#define P1D_THREADS 128
template <int KT, int NF, int MINB, int MG = 2>
__global__ void __launch_bounds__(P1D_THREADS, MINB) k_sweep_pass1_lin_dmma(const __grid_constant__ DevProblem P,
                                                                         const __grid_constant__ DevSweep S, long long first,
                                                                         long long count, peuqse_best* __restrict__ best_out,
                                                                         Best* __restrict__ partial, unsigned int* __restrict__ ticket,
                                                                         ShortHdr* __restrict__ sh, long long* __restrict__ c_idx,
                                                                         double* __restrict__ c_logp) {
    // the product is augmented by one column: [theta, 1] [Phi^T w ; -y w] = (sim - y) w, the weighted residual itself
    constexpr int KS = (KT + 4) / 4, KP = KS * 4, NW = P1D_THREADS / 32;
    __shared__ double s_bf[KS][NF][32];                 // B fragments: lane l of (s, f) holds w_j Phi[j = 8 f + l/4][4 s + l%4], -y_j w_j at column k
    __shared__ double s_th[NW][32][KP];                 // the warp's 32 points, zero padded to 4 KS parameters
    __shared__ Best s_best[NW];
    __shared__ bool s_last;
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const int k = P.k;
    for (int idx = threadIdx.x; idx < KS * NF * 32; idx += blockDim.x) {
        const int l = idx & 31, f = (idx >> 5) % NF, sstep = idx / (32 * NF);
        const int j = 8 * f + (l >> 2), kk = 4 * sstep + (l & 3);
        s_bf[sstep][f][l] = (kk < k) ? P.mconst[(size_t)j * k + kk] * P.w[j] : (kk == k ? -(P.y[j] * P.w[j]) : 0.0);
    }
    __syncthreads();
    // diagonal prior in natural order (the usual case): z = theta / scale - mu_scaled as a multiplication by 1/scale
    const bool fast_prior = P.prior_mode == PEUQSE_PRIOR_DIAG && P.prior_identity && P.kp == k;
    double rs[KT];                                              // 1 / scale, once (a division per use otherwise)
#pragma unroll
    for (int d = 0; d < KT; ++d) rs[d] = (d < k && P.scaling_on) ? 1.0 / P.scale[d] : 1.0;
    Best best = best_none();
    const long long wstride = (long long)gridDim.x * NW * 32;
    for (long long base = ((long long)blockIdx.x * NW + warp) * 32; base < count; base += wstride) {
        const long long i = base + lane;
        const bool valid = i < count;
        const long long g = first + (valid ? i : base);
        double gm = pq_neg_inf();                               // best logP published so far (shortlist mode)
        if (sh) {
            const unsigned long long key = *(volatile unsigned long long*)&sh->gmax_key;
            if (key) gm = short_key_to_double(key);
        }
        double th[KT];
        sweep_point<KT>(S, (unsigned long long)g, th);
        double lp;
        if (fast_prior) {
            bool ok = true;
            double q0 = 0.0, q1 = 0.0;
#pragma unroll
            for (int d = 0; d < KT; ++d) {
                if (d < k) {
                    if (P.has_ub) ok = ok && (th[d] <= P.ub[d]);
                    if (P.has_lb) ok = ok && (th[d] >= P.lb[d]);
                    const double zz = th[d] * rs[d] - P.mus[d];
                    const double t = zz * P.pw[d];
                    if (d & 1) q1 = fma(t, t, q1); else q0 = fma(t, t, q0);
                }
            }
            lp = ok ? -0.5 * (P.prior_c + (q0 + q1)) : pq_neg_inf();
        } else {
            lp = eval_prior<KT>(P, th);
        }
#pragma unroll
        for (int d = 0; d < KP; ++d) s_th[warp][lane][d] = (d < KT) ? ((d < k) ? th[d < KT ? d : 0] : (d == k ? 1.0 : 0.0)) : (d == k ? 1.0 : 0.0);
        __syncwarp();
        double maha = 0.0;
        // MG row fragments (8 MG points) at a time: every B fragment read feeds MG DMMAs, 2 MG independent epilogue chains
#pragma unroll
        for (int m = 0; m < 4; m += MG) {
            double acc[MG][NF][2];
#pragma unroll
            for (int r = 0; r < MG; ++r)
#pragma unroll
                for (int f = 0; f < NF; ++f) acc[r][f][0] = acc[r][f][1] = 0.0;
#pragma unroll
            for (int sstep = 0; sstep < KS; ++sstep) {
                double a[MG];
#pragma unroll
                for (int r = 0; r < MG; ++r) a[r] = s_th[warp][8 * (m + r) + (lane >> 2)][4 * sstep + (lane & 3)];
#pragma unroll
                for (int f = 0; f < NF; ++f) {
                    const double bq = s_bf[sstep][f][lane];
#pragma unroll
                    for (int r = 0; r < MG; ++r) pq_dmma884(acc[r][f][0], acc[r][f][1], a[r], bq);
                }
            }
#pragma unroll
            for (int r = 0; r < MG; ++r) {
                double pa = 0.0, pb = 0.0;                          // even / odd observation of the lane's pair, per fragment
#pragma unroll
                for (int f = 0; f < NF; ++f) {
                    pa = fma(acc[r][f][0], acc[r][f][0], pa);       // ((sim - y) w)^2
                    pb = fma(acc[r][f][1], acc[r][f][1], pb);
                }
                double part = pa + pb;
                part += __shfl_xor_sync(0xffffffffu, part, 1);
                part += __shfl_xor_sync(0xffffffffu, part, 2);      // the 4 lanes of row lane/4 hold the sum of point 8 (m + r) + lane/4
                const double v = __shfl_sync(0xffffffffu, part, 4 * (lane & 7));
                if ((lane >> 3) == m + r) maha = v;                 // this lane's own point
            }
        }
        __syncwarp();
        double ll = -0.5 * (P.resp_c[0] + maha);
        if (ll != ll) ll = pq_neg_inf();                        // IP.py:2880-2882
        const double va = pq_is_neg_inf(lp) ? lp : (ll + lp);   // IP.py:1649-1653
        if (valid) {
            if (va > best.v) { best.v = va; best.i = g; }       // strict '>', increasing g: first maximum
            if (sh) {
                if (va > gm - PQ_SHORT_MARGIN) short_push(sh, c_idx, c_logp, g, va);
                if (va > gm) atomicMax(&sh->gmax_key, short_double_to_key(va));
            }
        }
    }
    best = warp_best(best);
    if (lane == 0) s_best[warp] = best;
    __syncthreads();
    if (threadIdx.x < 32) {
        Best b = (threadIdx.x < NW) ? s_best[threadIdx.x] : best_none();
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
            if (sh) sh->valid = 1u;
        }
    }
}

// ---------------------------------------------------------------------------------------------
// pass 1: logP of every point (+ optional store) and the first strict maximum
// ---------------------------------------------------------------------------------------------
#define P1_THREADS 128
template <int MODEL, int KT, int SMEM>
__global__ void __launch_bounds__(P1_THREADS, 3) k_sweep_pass1(const __grid_constant__ DevProblem P, const __grid_constant__ DevSweep S,
                                                     long long first, long long count, double* __restrict__ logp_out,
                                                     peuqse_best* __restrict__ best_out, Best* __restrict__ partial,
                                                     unsigned int* __restrict__ ticket, int use_smem,
                                                     ShortHdr* __restrict__ sh, long long* __restrict__ c_idx,
                                                     double* __restrict__ c_logp) {
    extern __shared__ __align__(16) double smem[];
    __shared__ Best s_best[8];
    __shared__ bool s_last;
    const ObsView O = stage_obs<SMEM>(P, smem, use_smem);
    Best best = best_none();
    const long long stride = (long long)gridDim.x * blockDim.x;
    // two points per trip (i and i + stride) share every load of the per-observation constants
    for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < count; i += 2 * stride) {
        const long long g = first + i;
        const bool two = i + stride < count;
        double gm = pq_neg_inf();                               // best logP published so far (shortlist mode)
        if (sh) {
            const unsigned long long key = *(volatile unsigned long long*)&sh->gmax_key;
            if (key) gm = short_key_to_double(key);
        }
        double tha[KT], thb[KT];
        sweep_point<KT>(S, (unsigned long long)g, tha);
        sweep_point<KT>(S, (unsigned long long)(two ? g + stride : g), thb);
        const double lpa = eval_prior<KT>(P, tha), lpb = eval_prior<KT>(P, thb);
        double va = lpa, vb = lpb;
        if (!(pq_is_neg_inf(lpa) && pq_is_neg_inf(lpb))) {
            double lla, llb;
            eval_loglik_thread2<MODEL, KT>(P, O, tha, thb, lla, llb);
            if (!pq_is_neg_inf(lpa)) va = lla + lpa;            // IP.py:1649-1653
            if (!pq_is_neg_inf(lpb)) vb = llb + lpb;
        }
        if (logp_out) __stcs(logp_out + i, va);
        if (va > best.v) { best.v = va; best.i = g; }          // strict '>', increasing g: first maximum
        if (two) {
            if (logp_out) __stcs(logp_out + i + stride, vb);
            if (vb > best.v) { best.v = vb; best.i = g + stride; }
        }
        if (sh) {
            const double lim = gm - PQ_SHORT_MARGIN;
            if (va > lim) short_push(sh, c_idx, c_logp, g, va);
            if (two && vb > lim) short_push(sh, c_idx, c_logp, g + stride, vb);
            if (va > gm) { atomicMax(&sh->gmax_key, short_double_to_key(va)); gm = va; }
            if (two && vb > gm) atomicMax(&sh->gmax_key, short_double_to_key(vb));
        }
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
            if (sh) sh->valid = 1u;
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

// one retained point into the running sums (convertPermutationsToSamples' replication as a weight, IP.py:3485-3486)
template <int KT>
__device__ __forceinline__ void pass2_accumulate(Pass2Acc<KT>& A, const DevSweep& S, const double (&th)[KT],
                                                 const double (&shift)[KT], double v, double thr) {
    const double reps = exp(v - thr);                           // IP.py:3485
    const double w = trunc(reps);                               // int(...), IP.py:3486
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
}

template <int MODEL, int KT>
__global__ void __launch_bounds__(256) k_sweep_pass2(const __grid_constant__ DevProblem P, const __grid_constant__ DevSweep S,
                                                     long long first, long long count, const double* __restrict__ logp_in,
                                                     const peuqse_best* __restrict__ best_in, double log_rel,
                                                     double* __restrict__ partial, double* __restrict__ acc_out,
                                                     unsigned int* __restrict__ ticket, long long* __restrict__ keep_idx,
                                                     double* __restrict__ keep_logp, long long keep_cap,
                                                     unsigned long long* __restrict__ keep_count, int use_smem,
                                                     const ShortHdr* __restrict__ sh) {
    extern __shared__ double smem[];
    __shared__ double s_red[8];
    __shared__ bool s_last;
    if (sh && !sh->full_pass2) return;                         // k_short_pass2 already produced the statistics
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
        pass2_accumulate<KT>(A, S, th, shift, v, thr);
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
// pass 2 from the shortlist (logP not stored): one CTA filters the candidates against the final threshold, orders
// the retained ones by global index (bitonic sort in shared memory, so the floating-point sums do not depend on
// the order in which pass 1's threads appended them) and accumulates them.  When the list is unusable or more than
// PQ_SHORT_RCAP points are retained it raises full_pass2 and the recomputing kernel that follows does the work.
// ---------------------------------------------------------------------------------------------
template <int KT>
__global__ void __launch_bounds__(256) k_short_pass2(const __grid_constant__ DevSweep S, ShortHdr* __restrict__ sh,
                                                     const long long* __restrict__ c_idx, const double* __restrict__ c_logp,
                                                     const peuqse_best* __restrict__ best_in, double log_rel, int margin_ok,
                                                     double* __restrict__ acc_out, long long* __restrict__ keep_idx,
                                                     double* __restrict__ keep_logp, long long keep_cap,
                                                     unsigned long long* __restrict__ keep_count) {
    __shared__ long long s_idx[PQ_SHORT_RCAP];
    __shared__ double s_lp[PQ_SHORT_RCAP];
    __shared__ double s_red[8];
    __shared__ unsigned int s_n;
    if (!short_usable(sh, margin_ok)) {
        if (threadIdx.x == 0) sh->full_pass2 = 1u;
        return;
    }
    if (best_in->index < 0) {                                  // no finite point: nothing is retained
        if (threadIdx.x == 0) sh->full_pass2 = 0u;
        return;
    }
    const double thr = best_in->logp + log_rel;
    if (threadIdx.x == 0) s_n = 0u;
    __syncthreads();
    const unsigned int nc = (unsigned int)sh->count;
    for (unsigned int c = threadIdx.x; c < nc; c += blockDim.x) {
        const double v = c_logp[c];
        if (v > thr) {
            const unsigned int slot = atomicAdd(&s_n, 1u);
            if (slot < PQ_SHORT_RCAP) { s_idx[slot] = c_idx[c]; s_lp[slot] = v; }
        }
    }
    __syncthreads();
    const unsigned int nr = s_n;
    if (nr > PQ_SHORT_RCAP) {
        if (threadIdx.x == 0) sh->full_pass2 = 1u;
        return;
    }
    unsigned int np2 = 1;
    while (np2 < nr) np2 <<= 1;
    for (unsigned int c = nr + threadIdx.x; c < np2; c += blockDim.x) { s_idx[c] = 0x7fffffffffffffffLL; s_lp[c] = 0.0; }
    __syncthreads();
    for (unsigned int kk = 2; kk <= np2; kk <<= 1) {
        for (unsigned int j = kk >> 1; j > 0; j >>= 1) {
            for (unsigned int t = threadIdx.x; t < np2; t += blockDim.x) {
                const unsigned int x = t ^ j;
                if (x > t) {
                    const bool up = (t & kk) == 0;
                    const long long a = s_idx[t], b = s_idx[x];
                    if ((a > b) == up) {
                        s_idx[t] = b; s_idx[x] = a;
                        const double la = s_lp[t]; s_lp[t] = s_lp[x]; s_lp[x] = la;
                    }
                }
            }
            __syncthreads();
        }
    }
    double shift[KT];
    sweep_point<KT>(S, (unsigned long long)best_in->index, shift);
    Pass2Acc<KT> A;
    A.w = A.wexp = A.nk = 0.0;
#pragma unroll
    for (int d = 0; d < KT; ++d) { A.s1[d] = 0.0; A.mn[d] = pq_pos_inf(); A.mx[d] = pq_neg_inf(); }
#pragma unroll
    for (int e = 0; e < KT * (KT + 1) / 2; ++e) A.s2[e] = 0.0;
    for (unsigned int c = threadIdx.x; c < nr; c += blockDim.x) {
        double th[KT];
        sweep_point<KT>(S, (unsigned long long)s_idx[c], th);
        pass2_accumulate<KT>(A, S, th, shift, s_lp[c], thr);
        if (keep_idx && (long long)c < keep_cap) { keep_idx[c] = s_idx[c]; keep_logp[c] = s_lp[c]; }
    }
    if (keep_idx && threadIdx.x == 0) *keep_count += nr;
    const int k = S.k;
    double t;
    t = block_sum(A.w, s_red);    if (threadIdx.x == 0) acc_out[PEUQSE_ACC_W] += t;
    t = block_sum(A.wexp, s_red); if (threadIdx.x == 0) acc_out[PEUQSE_ACC_WEXP] += t;
    t = block_sum(A.nk, s_red);   if (threadIdx.x == 0) acc_out[PEUQSE_ACC_NKEPT] += t;
    int e = 0;
#pragma unroll
    for (int d = 0; d < KT; ++d) {
        if (d < k) {
            t = block_sum(A.s1[d], s_red);
            if (threadIdx.x == 0) acc_out[PEUQSE_ACC_S1 + d] += t;
            t = block_min(A.mn[d], s_red);
            if (threadIdx.x == 0) acc_out[PEUQSE_ACC_MIN(k) + d] = fmin(acc_out[PEUQSE_ACC_MIN(k) + d], t);
            t = -block_min(-A.mx[d], s_red);
            if (threadIdx.x == 0) acc_out[PEUQSE_ACC_MAX(k) + d] = fmax(acc_out[PEUQSE_ACC_MAX(k) + d], t);
        }
#pragma unroll
        for (int d2 = d; d2 < KT; ++d2) {
            if (d < k && d2 < k) {
                t = block_sum(A.s2[e], s_red);
                if (threadIdx.x == 0) {
                    acc_out[PEUQSE_ACC_S2(k) + d * k + d2] += t;
                    if (d2 != d) acc_out[PEUQSE_ACC_S2(k) + d2 * k + d] += t;
                }
            }
            ++e;
        }
    }
    if (threadIdx.x == 0) sh->full_pass2 = 0u;
}

// np.histogram bin of x on the given edges (numpy/lib/_histograms_impl.py: index from the normalised position,
// corrected against the actual edges; the last bin is closed on the right); -1 when x is outside [lo, hi]
__device__ __forceinline__ int hist_bin(const double* __restrict__ ed, int nbins, double x) {
    const double lo = ed[0], hi = ed[nbins];
    if (!(x >= lo && x <= hi)) return -1;
    int b = (int)(((x - lo) / (hi - lo)) * (double)nbins);
    if (b >= nbins) b = nbins - 1;
    if (b < 0) b = 0;
    if (x < ed[b]) --b;
    else if (b != nbins - 1 && x >= ed[b + 1]) ++b;
    return b;
}

// pass 3 from the shortlist: integer weight sums, so the order of the candidates does not matter
template <int KT>
__global__ void __launch_bounds__(256) k_short_hist(const __grid_constant__ DevSweep S, const ShortHdr* __restrict__ sh,
                                                    const long long* __restrict__ c_idx, const double* __restrict__ c_logp,
                                                    const peuqse_best* __restrict__ best_in, double log_rel, int margin_ok,
                                                    const double* __restrict__ edges, int nbins,
                                                    unsigned long long* __restrict__ hist) {
    if (!short_usable(sh, margin_ok) || best_in->index < 0) return;
    const double thr = best_in->logp + log_rel;
    const unsigned int nc = (unsigned int)sh->count;
    for (unsigned int c = blockIdx.x * blockDim.x + threadIdx.x; c < nc; c += gridDim.x * blockDim.x) {
        const double v = c_logp[c];
        if (!(v > thr)) continue;
        double th[KT];
        sweep_point<KT>(S, (unsigned long long)c_idx[c], th);
        const unsigned long long w = (unsigned long long)trunc(exp(v - thr));
#pragma unroll
        for (int d = 0; d < KT; ++d) {
            if (d < S.k) {
                const int b = hist_bin(edges + d * (nbins + 1), nbins, th[d]);
                if (b >= 0) atomicAdd(hist + d * nbins + b, w);
            }
        }
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
                                                    unsigned long long* __restrict__ hist, int use_smem,
                                                    const ShortHdr* __restrict__ shl, int margin_ok) {
    extern __shared__ double smem[];
    if (short_usable(shl, margin_ok)) return;                  // k_short_hist binned the retained points already
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
                const int b = hist_bin(edges + d * (nbins + 1), nbins, th[d]);
                if (b >= 0) atomicAdd(sh + d * nbins + b, w);
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

// linear model with k == KT (even): the instantiation that reads whole rows as double2 (eval.cuh)
template <int MODEL, int KT>
struct FullRows {
    static constexpr int value = (MODEL == PEUQSE_MODEL_LINEAR && KT % 2 == 0) ? PQ_MODEL_LINEAR_FULL : MODEL;
};

template <int MODEL, int KT>
int launch_pass1(const peuqse_problem* p, const DevSweep& S, long long first, long long count, double* logp,
                 peuqse_best* best, void* work, cudaStream_t st) {
    int use_smem;
    const size_t smem = pq_pick_smem(p, &use_smem);
    constexpr int MF = FullRows<MODEL, KT>::value;
    const bool full = MF != MODEL && p->d.k == KT;
    void (*kern)(const DevProblem, const DevSweep, long long, long long, double*, peuqse_best*, Best*, unsigned int*, int,
                 ShortHdr*, long long*, double*) =
        full ? (use_smem ? k_sweep_pass1<MF, KT, 1> : k_sweep_pass1<MF, KT, -1>)
             : (use_smem ? k_sweep_pass1<MODEL, KT, 1> : k_sweep_pass1<MODEL, KT, -1>);
    PQ_CUDA_OK(pq_allow_smem((const void*)kern, smem));
    unsigned int* ticket = (unsigned int*)work;
    Best* partial = (Best*)((char*)work + PQ_SWEEP_TICKET_BYTES);
    if constexpr (MODEL == PEUQSE_MODEL_LINEAR) {
        // logP not kept + linear model + one diagonal-covariance response of 32 / 64 / 128 observations: tensor-pipe simulation
        static int off = -1;
        if (off < 0) { const char* e = getenv("PEUQSE_SWEEP_NO_DMMA"); off = (e && e[0] == '1') ? 1 : 0; }
        const int n = p->d.n_total;
        if (!off && !logp && p->d.n_resp == 1 && p->d.resp_mode[0] == PEUQSE_RESP_MVN_DIAG && !p->d.any_per_value &&
            (n == 32 || n == 64 || n == 128) && KT <= 12) {
            ShortHdr* shd = (ShortHdr*)((char*)work + PQ_SHORT_HDR_OFFSET);
            long long* ci = (long long*)((char*)work + PQ_SWEEP_BASE_BYTES);
            double* cl = (double*)(ci + PQ_SHORT_CAP);
            static int minb = -1;
            if (minb < 0) { const char* e = getenv("PEUQSE_SWEEP_DMMA_CTAS"); minb = e ? atoi(e) : 4; }
            const long long wantd = (count + 127) / 128;
            const int per_sm = minb == 4 ? 4 : (minb == 5 ? 5 : 3);    // 5: three CTAs per SM, all four row fragments of a warp at once
            const int gd = (int)std::max<long long>(1, std::min<long long>(wantd, std::min<long long>((long long)p->sm_count * (per_sm == 5 ? 3 : per_sm), PQ_SWEEP_MAX_BLOCKS)));
            if (n == 32) k_sweep_pass1_lin_dmma<KT, 4, 4><<<gd, P1D_THREADS, 0, st>>>(p->d, S, first, count, best, partial, ticket, shd, ci, cl);
            else if (n == 64 && per_sm == 5) k_sweep_pass1_lin_dmma<KT, 8, 3, 4><<<gd, P1D_THREADS, 0, st>>>(p->d, S, first, count, best, partial, ticket, shd, ci, cl);
            else if (n == 64 && per_sm == 4) k_sweep_pass1_lin_dmma<KT, 8, 4><<<gd, P1D_THREADS, 0, st>>>(p->d, S, first, count, best, partial, ticket, shd, ci, cl);
            else if (n == 64) k_sweep_pass1_lin_dmma<KT, 8, 3><<<gd, P1D_THREADS, 0, st>>>(p->d, S, first, count, best, partial, ticket, shd, ci, cl);
            else k_sweep_pass1_lin_dmma<KT, 16, 3><<<gd, P1D_THREADS, 0, st>>>(p->d, S, first, count, best, partial, ticket, shd, ci, cl);
            PQ_CUDA_OK(cudaGetLastError());
            return PEUQSE_OK;
        }
    }
    // 128-thread CTAs, three per SM (168 registers for the two points a thread carries); 12 CTAs queued per SM
    const long long want = (count + 2 * P1_THREADS - 1) / (2 * P1_THREADS);
    const int grid = (int)std::max<long long>(1, std::min<long long>(want, std::min<long long>((long long)p->sm_count * 12, PQ_SWEEP_MAX_BLOCKS)));
    // logP not kept: shortlist mode (sweep_common.cuh)
    ShortHdr* sh = logp ? nullptr : (ShortHdr*)((char*)work + PQ_SHORT_HDR_OFFSET);
    long long* c_idx = (long long*)((char*)work + PQ_SWEEP_BASE_BYTES);
    double* c_logp = (double*)(c_idx + PQ_SHORT_CAP);
    kern<<<grid, P1_THREADS, smem, st>>>(p->d, S, first, count, logp, best, partial, ticket, use_smem, sh, c_idx, c_logp);
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
    ShortHdr* sh = nullptr;
    if (!logp) {                                                // statistics from pass 1's shortlist when it is usable
        sh = (ShortHdr*)((char*)work + PQ_SHORT_HDR_OFFSET);
        const long long* c_idx = (const long long*)((char*)work + PQ_SWEEP_BASE_BYTES);
        const double* c_logp = (const double*)(c_idx + PQ_SHORT_CAP);
        k_short_pass2<KT><<<1, 256, 0, st>>>(S, sh, c_idx, c_logp, best, log_rel, (-log_rel <= PQ_SHORT_MARGIN) ? 1 : 0, acc, kidx, klogp,
                                             kcap, (unsigned long long*)kcount);
        PQ_CUDA_OK(cudaGetLastError());
    }
    kern<<<sweep_grid(p, count), 256, smem, st>>>(p->d, S, first, count, logp, best, log_rel, partial, acc, ticket, kidx, klogp, kcap,
                                                  (unsigned long long*)kcount, use_smem, sh);
    PQ_CUDA_OK(cudaGetLastError());
    return PEUQSE_OK;
}

template <int MODEL, int KT>
int launch_hist(const peuqse_problem* p, const DevSweep& S, long long first, long long count, const double* logp,
                const peuqse_best* best, double log_rel, const double* edges, int nbins, unsigned long long* hist, void* work,
                cudaStream_t st) {
    int use_smem;
    size_t smem = pq_pick_smem(p, &use_smem);
    smem += (size_t)S.k * nbins * sizeof(unsigned long long);
    auto kern = k_sweep_hist<MODEL, KT>;
    PQ_CUDA_OK(pq_allow_smem((const void*)kern, smem));
    const ShortHdr* sh = nullptr;
    const int margin_ok = (-log_rel <= PQ_SHORT_MARGIN) ? 1 : 0;
    if (!logp && work) {
        sh = (const ShortHdr*)((char*)work + PQ_SHORT_HDR_OFFSET);
        const long long* c_idx = (const long long*)((char*)work + PQ_SWEEP_BASE_BYTES);
        const double* c_logp = (const double*)(c_idx + PQ_SHORT_CAP);
        k_short_hist<KT><<<64, 256, 0, st>>>(S, sh, c_idx, c_logp, best, log_rel, margin_ok, edges, nbins, hist);
        PQ_CUDA_OK(cudaGetLastError());
    }
    kern<<<sweep_grid(p, count), 256, smem, st>>>(p->d, S, first, count, logp, best, log_rel, edges, nbins, hist, use_smem, sh, margin_ok);
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
    const long long shortlist = PQ_SHORT_CAP * 16LL;          // (index, logP) entries of the no-store mode
    return PQ_SWEEP_BASE_BYTES + std::max(aux, shortlist);
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
    PQ_CUDA_OK(cudaMemsetAsync(d_work, 0, PQ_SHORT_HDR_OFFSET, st));   // the ticket; pass 1's shortlist header stays
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
#define P3_ARGS p, S, first, count, d_logp, d_best, log_rel_threshold, d_edges, nbins, (unsigned long long*)d_hist, d_work, st
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
