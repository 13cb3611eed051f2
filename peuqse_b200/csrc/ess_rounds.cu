// Round-synchronous ensemble slice sampling (kernel K4 of SURVEY section 2, batched form): the driver of
// doEnsembleSliceSampling (IP.py:2362-2547 -> zeus.EnsembleSampler.run_mcmc, call sites IP.py:2495-2518) for likelihoods that
// must be evaluated in BATCHES -- the full-covariance DMMA contraction (csrc/fullcov.cu, config C4) and torch-callable
// simulators.  Algorithm: Karamanis & Beutler, arXiv:2002.06212, differential move (zeus is third-party and not vendored:
// parity against zeus unpinned; pinned against oracle.ess_run_philox and csrc/ess.cu on the same Philox draws).
//
// A half-sweep is a sequence of ROUNDS.  Every round the device packs the points that still need a log-posterior into
// one dense block of candidate rows -- stepping out: the left and the right interval ends of every walker whose end is
// still moving (both ends in the same round: no draws are consumed while stepping out, so the order of evaluation is
// free); shrinking: one fresh point per walker that has not accepted yet -- the caller evaluates the block with ONE
// batched launch, and the next round's kernels apply the results (masked updates) and pack the next block.  Two small
// kernels per round (apply + count, scatter), whatever the ensemble size; the only host traffic is the row count of the
// block (mapped pinned memory, announced by a ticket word the host spins on: no stream synchronisation per round).
// The order of the rows in a block is deterministic (left ends by slot, right ends by slot / walkers by slot).
#include <atomic>
#include <chrono>
#include <cstring>
#include <new>
#include "common.cuh"
#include "launch.cuh"
#include "philox.cuh"
#include "ess_common.cuh"

#define ER_THREADS 256

struct EssrDev {
    EssState* st;
    int* perm;
    unsigned long long* keys;
    int* wid;                       // [H] walker of slot ia
    double *x0, *dir;               // [H][k]
    double *z0, *Lb, *Rb, *J, *K;   // [H] slice level, interval ends, remaining stepping-out budgets
    double *wd, *znew, *wdacc;      // [H] shrink: current point on the line, accepted log-posterior / point
    unsigned int* rsh;              // [H] shrink rounds done
    unsigned char *actL, *actR, *actS, *moved;   // [H]
    int *rowL, *rowR, *rowS;        // [H] row of the current candidate block that holds the task's point
    int *cntL, *cntR, *cntS;        // [G] per-CTA numbers of active tasks
    int* ctl;                       // [0] phase of the block just written (0 stepping out, 1 shrinking) [1] rows [2] stepping-out round
    long long* h_status;            // mapped pinned host memory: [0] rows of the block, [1] phase, [2] error code, [3] ticket
};

struct peuqse_essr {
    EssrDev D;
    DevEss E;
    long long H;
    int G;
    void* d_work;
    long long* h_status;
    long long it;
    long long ticket;               // number of candidate blocks announced so far (the host waits for h_status[3] == ticket)
    int half;
};

namespace {

// numbers of raised flags in this CTA -> cnt*[blockIdx.x]
__device__ __forceinline__ void essr_counts(const EssrDev& D, bool fL, bool fR, bool fS) {
    const int nL = __syncthreads_count(fL), nR = __syncthreads_count(fR), nS = __syncthreads_count(fS);
    if (threadIdx.x == 0) { D.cntL[blockIdx.x] = nL; D.cntR[blockIdx.x] = nR; D.cntS[blockIdx.x] = nS; }
}

// start of a half-sweep: direction, slice level, initial interval and stepping-out budgets of every active walker
// (the preamble of k_ess_half, csrc/ess.cu, draw for draw)
__global__ void __launch_bounds__(ER_THREADS) k_essr_setup(const EssrDev D, const DevEss E, long long it, int half,
                                                          const double* __restrict__ X, const double* __restrict__ Z) {
    const long long H = E.W / 2;
    const long long ia = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    bool fL = false, fR = false, fS = false;
    if (ia < H) {
        const int k = E.k;
        const long long w = D.perm[half * H + ia];
        const int* comp = D.perm + (1 - half) * H;
        const unsigned long long idx = ess_idx(E, it, half, w);
        const double mu = D.st->mu;
        const Philox4 r0 = philox_block(E.seed, PQ_STREAM_ESS, idx, 0u);
        long long i1 = (long long)(philox_u53(r0.x, r0.y) * (double)H);
        long long i2 = (long long)(philox_u53(r0.z, r0.w) * (double)(H - 1));
        if (i1 >= H) i1 = H - 1;
        if (i2 >= H - 1) i2 = H - 2;
        if (i2 >= i1) ++i2;
        const long long c1 = comp[i1], c2 = comp[i2];
        for (int d = 0; d < k; ++d) {
            D.x0[ia * k + d] = X[w * k + d];
            D.dir[ia * k + d] = 2.0 * mu * (X[c1 * k + d] - X[c2 * k + d]);
        }
        const Philox4 r1 = philox_block(E.seed, PQ_STREAM_ESS, idx, 1u);
        const double zw = Z[w];
        D.z0[ia] = zw - (-log(1.0 - philox_u53(r1.x, r1.y)));              // slice level: logP - Exp(1)
        const double L = -philox_u53(r1.z, r1.w);
        D.Lb[ia] = L;
        D.Rb[ia] = L + 1.0;
        const Philox4 r2 = philox_block(E.seed, PQ_STREAM_ESS, idx, 2u);
        const double J = floor(ESS_MAXSTEPS * philox_u53(r2.x, r2.y));
        const double K = (ESS_MAXSTEPS - 1.0) - J;
        D.J[ia] = J;
        D.K[ia] = K;
        D.wid[ia] = (int)w;
        D.znew[ia] = zw;
        D.rsh[ia] = 0u;
        D.moved[ia] = 0;
        fL = J >= 1.0; fR = K >= 1.0; fS = true;
        D.actL[ia] = fL; D.actR[ia] = fR; D.actS[ia] = 1;
    }
    essr_counts(D, fL, fR, fS);
    if (ia == 0) D.ctl[2] = 0;
}

// exclusive rank of this thread's flag within the CTA (thread order) + the CTA's base
__device__ __forceinline__ int essr_rank(bool f, int base, int* s_warp) {
    const unsigned b = __ballot_sync(0xffffffffu, f);
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    __syncthreads();
    if (lane == 0) s_warp[warp] = __popc(b);
    __syncthreads();
    int off = base;
    for (int v = 0; v < warp; ++v) off += s_warp[v];
    return off + __popc(b & ((1u << lane) - 1u));
}

// pack the next block of candidate rows: stepping out while any interval end is still moving, shrinking afterwards
__global__ void __launch_bounds__(ER_THREADS) k_essr_scatter(const EssrDev D, const DevEss E, long long it, int half,
                                                            double* __restrict__ cand, long long ticket) {
    __shared__ int s_red[6][ER_THREADS / 32];
    __shared__ int s_tot[6];
    __shared__ int s_warp[ER_THREADS / 32];
    const long long H = E.W / 2;
    const int G = gridDim.x, me = blockIdx.x;
    // totals and the sums over the preceding CTAs of the three per-CTA counts
    int v[6] = {0, 0, 0, 0, 0, 0};
    for (int b = threadIdx.x; b < G; b += blockDim.x) {
        const int l = D.cntL[b], r = D.cntR[b], s = D.cntS[b];
        v[0] += l; v[1] += r; v[2] += s;
        if (b < me) { v[3] += l; v[4] += r; v[5] += s; }
    }
#pragma unroll
    for (int q = 0; q < 6; ++q) {
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) v[q] += __shfl_xor_sync(0xffffffffu, v[q], o);
        if ((threadIdx.x & 31) == 0) s_red[q][threadIdx.x >> 5] = v[q];
    }
    __syncthreads();
    if (threadIdx.x < 6) {
        int t = 0;
        for (int w = 0; w < ER_THREADS / 32; ++w) t += s_red[threadIdx.x][w];
        s_tot[threadIdx.x] = t;
    }
    __syncthreads();
    const int totL = s_tot[0], totR = s_tot[1], totS = s_tot[2], preL = s_tot[3], preR = s_tot[4], preS = s_tot[5];
    const bool stepping = (totL + totR) > 0;
    const long long ia = (long long)me * blockDim.x + threadIdx.x;
    const bool in = ia < H;
    const int k = E.k;
    if (stepping) {
        const bool fL = in && D.actL[ia], fR = in && D.actR[ia];
        const int rl = essr_rank(fL, preL, s_warp);
        const int rr = essr_rank(fR, totL + preR, s_warp);
        if (fL) {
            const double b = D.Lb[ia];
            for (int d = 0; d < k; ++d) cand[(long long)rl * k + d] = D.x0[ia * k + d] + b * D.dir[ia * k + d];
            D.rowL[ia] = rl;
        }
        if (fR) {
            const double b = D.Rb[ia];
            for (int d = 0; d < k; ++d) cand[(long long)rr * k + d] = D.x0[ia * k + d] + b * D.dir[ia * k + d];
            D.rowR[ia] = rr;
        }
    } else {
        const bool fS = in && D.actS[ia];
        const int rs = essr_rank(fS, preS, s_warp);
        if (fS) {
            const unsigned int r = D.rsh[ia];
            const Philox4 rr = philox_block(E.seed, PQ_STREAM_ESS, ess_idx(E, it, half, D.wid[ia]), 3u + (r >> 1));
            const double u = (r & 1u) ? philox_u53(rr.z, rr.w) : philox_u53(rr.x, rr.y);
            const double L = D.Lb[ia], R = D.Rb[ia];
            const double wd = L + u * (R - L);
            D.wd[ia] = wd;
            for (int d = 0; d < k; ++d) cand[(long long)rs * k + d] = D.x0[ia * k + d] + wd * D.dir[ia * k + d];
            D.rowS[ia] = rs;
        }
    }
    if (me == 0 && threadIdx.x == 0) {
        const int rows = stepping ? totL + totR : totS;
        D.ctl[0] = stepping ? 0 : 1;
        D.ctl[1] = rows;
        if (stepping) D.ctl[2] += 1;
        D.h_status[0] = rows;
        D.h_status[1] = stepping ? 0 : 1;
        D.h_status[2] = D.st->error;
        __threadfence_system();
    }
    // the ticket goes out when EVERY CTA has written its candidate rows: last CTA to arrive announces the block to the host
    __syncthreads();
    if (threadIdx.x == 0) {
        __threadfence();
        const unsigned int done = atomicAdd(reinterpret_cast<unsigned int*>(D.ctl + 3), 1u);
        if (done == gridDim.x - 1) {
            D.ctl[3] = 0;
            __threadfence_system();
            *(volatile long long*)(D.h_status + 3) = ticket;
            __threadfence_system();
        }
    }
}

// apply the log-posteriors of the block the previous scatter wrote (masked updates), count what is still active
__global__ void __launch_bounds__(ER_THREADS) k_essr_apply(const EssrDev D, const DevEss E, const double* __restrict__ logp_rows) {
    __shared__ unsigned long long s_cnt[3];
    const long long H = E.W / 2;
    const long long ia = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    const int phase = D.ctl[0];
    if (threadIdx.x < 3) s_cnt[threadIdx.x] = 0ull;
    __syncthreads();
    bool fL = false, fR = false, fS = false;
    unsigned int nexp = 0, ncon = 0, neval = 0;
    int err = 0;
    if (ia < H) {
        const double z0 = D.z0[ia];
        fL = D.actL[ia]; fR = D.actR[ia]; fS = D.actS[ia];
        if (phase == 0) {
            const int round = D.ctl[2] - 1;                                  // == `rounds` of the loops in k_ess_half
            if (fL) {
                const double z = logp_rows[D.rowL[ia]];
                ++neval;
                if (z0 < z) {
                    const double j = D.J[ia] - 1.0;
                    D.Lb[ia] -= 1.0; D.J[ia] = j; ++nexp;
                    fL = j >= 1.0;
                    if (round > E.max_rounds) err = 1;
                } else fL = false;
                D.actL[ia] = fL;
            }
            if (fR) {
                const double z = logp_rows[D.rowR[ia]];
                ++neval;
                if (z0 < z) {
                    const double kk = D.K[ia] - 1.0;
                    D.Rb[ia] += 1.0; D.K[ia] = kk; ++nexp;
                    fR = kk >= 1.0;
                    if (round > E.max_rounds) err = 1;
                } else fR = false;
                D.actR[ia] = fR;
            }
        } else if (fS) {
            const double z = logp_rows[D.rowS[ia]];
            const double wd = D.wd[ia];
            ++neval;
            if (z0 < z) {
                D.znew[ia] = z; D.wdacc[ia] = wd; D.moved[ia] = 1;
                fS = false;
            } else {
                if (wd < 0.0) D.Lb[ia] = wd;
                else if (wd > 0.0) D.Rb[ia] = wd;
                ++ncon;
                const unsigned int r = D.rsh[ia];
                if ((int)r > E.max_rounds) err = 2;
                D.rsh[ia] = r + 1u;
            }
            D.actS[ia] = fS;
        }
        if (err) { fL = fR = fS = false; D.actL[ia] = D.actR[ia] = D.actS[ia] = 0; }
    }
    essr_counts(D, fL, fR, fS);
    // counters: warp sums, one shared-memory atomic per warp, one global atomic per CTA
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
        nexp += __shfl_xor_sync(0xffffffffu, nexp, o);
        ncon += __shfl_xor_sync(0xffffffffu, ncon, o);
        neval += __shfl_xor_sync(0xffffffffu, neval, o);
    }
    if ((threadIdx.x & 31) == 0) {
        if (nexp) atomicAdd(&s_cnt[0], (unsigned long long)nexp);
        if (ncon) atomicAdd(&s_cnt[1], (unsigned long long)ncon);
        if (neval) atomicAdd(&s_cnt[2], (unsigned long long)neval);
    }
    if (err) atomicMax(&D.st->error, err);
    __syncthreads();
    if (threadIdx.x == 0) {
        if (s_cnt[0]) atomicAdd(&D.st->nexp, s_cnt[0]);
        if (s_cnt[1]) atomicAdd(&D.st->ncon, s_cnt[1]);
        if (s_cnt[2]) atomicAdd(&D.st->neval, s_cnt[2]);
    }
}

// end of a half-sweep: the walkers that accepted a point move there
__global__ void __launch_bounds__(ER_THREADS) k_essr_commit(const EssrDev D, const DevEss E, double* __restrict__ X,
                                                           double* __restrict__ Z) {
    const long long H = E.W / 2;
    const long long ia = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (ia >= H || !D.moved[ia]) return;
    const int k = E.k;
    const long long w = D.wid[ia];
    const double wd = D.wdacc[ia];
    for (int d = 0; d < k; ++d) X[w * k + d] = D.x0[ia * k + d] + wd * D.dir[ia * k + d];
    Z[w] = D.znew[ia];
}

size_t align256(size_t v) { return (v + 255) & ~(size_t)255; }

}  // namespace

extern "C" int peuqse_essr_create(const peuqse_ess_spec* s, peuqse_essr** out) {
    if (!s || !out) return pq_set_error(PEUQSE_ERR_INVALID, "null argument");
    if (s->n_walkers < 4 || (s->n_walkers & 1)) return pq_set_error(PEUQSE_ERR_INVALID, "ess needs an even number (>= 4) of walkers");
    if (s->n_walkers > (1LL << 30)) return pq_set_error(PEUQSE_ERR_INVALID, "too many walkers");
    if (s->k < 1 || s->k > PQ_MAXK) return pq_set_error(PEUQSE_ERR_INVALID, "k out of range");
    peuqse_essr* h = new (std::nothrow) peuqse_essr();
    if (!h) return pq_set_error(PEUQSE_ERR_NOMEM, "out of host memory");
    memset(h, 0, sizeof *h);
    h->E.W = s->n_walkers; h->E.first_iter = s->first_iter; h->E.seed = s->seed; h->E.k = s->k;
    h->E.max_rounds = s->max_rounds > 0 ? s->max_rounds : 10000;
    const long long W = s->n_walkers, H = W / 2;
    const int k = s->k;
    h->H = H;
    h->G = (int)((H + ER_THREADS - 1) / ER_THREADS);
    size_t off = 0;
    auto take = [&](size_t bytes) { const size_t o = off; off = align256(off + bytes); return o; };
    const size_t o_st = take(sizeof(EssState)), o_perm = take(W * sizeof(int)), o_keys = take(W * sizeof(unsigned long long)),
                 o_wid = take(H * sizeof(int)), o_x0 = take(H * k * sizeof(double)), o_dir = take(H * k * sizeof(double));
    size_t o_d[8];
    for (int i = 0; i < 8; ++i) o_d[i] = take(H * sizeof(double));
    const size_t o_rsh = take(H * sizeof(unsigned int));
    size_t o_f[4], o_r[3], o_c[3];
    for (int i = 0; i < 4; ++i) o_f[i] = take(H);
    for (int i = 0; i < 3; ++i) o_r[i] = take(H * sizeof(int));
    for (int i = 0; i < 3; ++i) o_c[i] = take(h->G * sizeof(int));
    const size_t o_ctl = take(4 * sizeof(int));
    cudaError_t e = cudaMalloc(&h->d_work, off);
    if (e != cudaSuccess) { delete h; return pq_set_cuda_error(e, "cudaMalloc(ess rounds state)", __LINE__); }
    e = cudaHostAlloc((void**)&h->h_status, 4 * sizeof(long long), cudaHostAllocMapped);
    if (e != cudaSuccess) { cudaFree(h->d_work); delete h; return pq_set_cuda_error(e, "cudaHostAlloc(ess rounds status)", __LINE__); }
    memset(h->h_status, 0, 4 * sizeof(long long));
    char* b = (char*)h->d_work;
    EssrDev& D = h->D;
    D.st = (EssState*)(b + o_st); D.perm = (int*)(b + o_perm); D.keys = (unsigned long long*)(b + o_keys);
    D.wid = (int*)(b + o_wid); D.x0 = (double*)(b + o_x0); D.dir = (double*)(b + o_dir);
    D.z0 = (double*)(b + o_d[0]); D.Lb = (double*)(b + o_d[1]); D.Rb = (double*)(b + o_d[2]); D.J = (double*)(b + o_d[3]);
    D.K = (double*)(b + o_d[4]); D.wd = (double*)(b + o_d[5]); D.znew = (double*)(b + o_d[6]); D.wdacc = (double*)(b + o_d[7]);
    D.rsh = (unsigned int*)(b + o_rsh);
    D.actL = (unsigned char*)(b + o_f[0]); D.actR = (unsigned char*)(b + o_f[1]); D.actS = (unsigned char*)(b + o_f[2]);
    D.moved = (unsigned char*)(b + o_f[3]);
    D.rowL = (int*)(b + o_r[0]); D.rowR = (int*)(b + o_r[1]); D.rowS = (int*)(b + o_r[2]);
    D.cntL = (int*)(b + o_c[0]); D.cntR = (int*)(b + o_c[1]); D.cntS = (int*)(b + o_c[2]);
    D.ctl = (int*)(b + o_ctl);
    e = cudaHostGetDevicePointer((void**)&D.h_status, h->h_status, 0);
    if (e != cudaSuccess) { cudaFreeHost(h->h_status); cudaFree(h->d_work); delete h; return pq_set_cuda_error(e, "cudaHostGetDevicePointer", __LINE__); }
    e = cudaMemset(h->d_work, 0, off);
    if (e != cudaSuccess) { cudaFreeHost(h->h_status); cudaFree(h->d_work); delete h; return pq_set_cuda_error(e, "cudaMemset", __LINE__); }
    k_ess_init<<<1, 1>>>(D.st, s->mu0 > 0 ? s->mu0 : 1.0, s->tune);
    e = cudaDeviceSynchronize();
    if (e != cudaSuccess) { cudaFreeHost(h->h_status); cudaFree(h->d_work); delete h; return pq_set_cuda_error(e, "k_ess_init", __LINE__); }
    *out = h;
    return PEUQSE_OK;
}

extern "C" int peuqse_essr_destroy(peuqse_essr* h) {
    if (!h) return PEUQSE_OK;
    cudaFreeHost(h->h_status);
    cudaFree(h->d_work);
    delete h;
    return PEUQSE_OK;
}

// The host learns the size of the next block from mapped pinned memory: it SPINS on the ticket word the scatter kernel writes
// last (a stream synchronisation per round goes through the driver's blocking wait, whose wake-up latency -- tens of
// microseconds, far more on a busy host -- would be paid ~40 times per ensemble iteration); the kernels that follow are
// stream-ordered behind the scatter kernel anyway.  If the ticket does not arrive within 20 s the stream is synchronised
// to surface the error.
static int essr_read_status(peuqse_essr* h, int64_t* n_rows, int32_t* phase, cudaStream_t st) {
    const volatile long long* hs = h->h_status;
    {
        const auto t0 = std::chrono::steady_clock::now();
        unsigned long long spins = 0;
        while (hs[3] != h->ticket) {
            if ((++spins & 0xfffffULL) == 0 && std::chrono::steady_clock::now() - t0 > std::chrono::seconds(20)) {
                PQ_CUDA_OK(cudaStreamSynchronize(st));
                if (hs[3] != h->ticket) return pq_set_error(PEUQSE_ERR_CUDA, "ensemble slice sampling: the device did not announce the next round");
            }
        }
        std::atomic_thread_fence(std::memory_order_acquire);
    }
    if (n_rows) *n_rows = hs[0];
    if (phase) *phase = (int32_t)hs[1];
    if (hs[2] == 1) return pq_set_error(PEUQSE_ERR_INVALID, "Number of expansions exceeded maximum limit! Increase mcmc_maxiter.");
    if (hs[2] == 2) return pq_set_error(PEUQSE_ERR_INVALID, "Number of contractions exceeded maximum limit! Increase mcmc_maxiter.");
    return PEUQSE_OK;
}

extern "C" int peuqse_essr_begin_half(peuqse_essr* h, int64_t it, int32_t half, const double* d_x, const double* d_logp_x,
                                      double* d_cand, int64_t* n_rows, void* stream) {
    if (!h || !d_x || !d_logp_x || !d_cand || !n_rows || half < 0 || half > 1) return pq_set_error(PEUQSE_ERR_INVALID, "bad argument");
    cudaStream_t st = (cudaStream_t)stream;
    h->it = it;
    h->half = half;
    if (half == 0) {
        const unsigned wb = (unsigned)((h->E.W + 127) / 128);
        k_ess_shuffle<<<wb, 128, 0, st>>>(h->E, it, h->D.perm, h->D.keys);
        k_ess_rank<<<wb, 128, 0, st>>>(h->E, h->D.keys, h->D.perm);
    }
    k_essr_setup<<<h->G, ER_THREADS, 0, st>>>(h->D, h->E, it, half, d_x, d_logp_x);
    k_essr_scatter<<<h->G, ER_THREADS, 0, st>>>(h->D, h->E, it, half, d_cand, ++h->ticket);
    PQ_CUDA_OK(cudaGetLastError());
    return essr_read_status(h, n_rows, nullptr, st);
}

extern "C" int peuqse_essr_round(peuqse_essr* h, const double* d_logp_rows, double* d_cand, int64_t* n_rows, void* stream) {
    if (!h || !d_logp_rows || !d_cand || !n_rows) return pq_set_error(PEUQSE_ERR_INVALID, "bad argument");
    cudaStream_t st = (cudaStream_t)stream;
    k_essr_apply<<<h->G, ER_THREADS, 0, st>>>(h->D, h->E, d_logp_rows);
    k_essr_scatter<<<h->G, ER_THREADS, 0, st>>>(h->D, h->E, h->it, h->half, d_cand, ++h->ticket);
    PQ_CUDA_OK(cudaGetLastError());
    return essr_read_status(h, n_rows, nullptr, st);
}

extern "C" int peuqse_essr_end_half(peuqse_essr* h, double* d_x, double* d_logp_x, void* stream) {
    if (!h || !d_x || !d_logp_x) return pq_set_error(PEUQSE_ERR_INVALID, "bad argument");
    k_essr_commit<<<h->G, ER_THREADS, 0, (cudaStream_t)stream>>>(h->D, h->E, d_x, d_logp_x);
    PQ_CUDA_OK(cudaGetLastError());
    return PEUQSE_OK;
}

extern "C" int peuqse_essr_end_iteration(peuqse_essr* h, void* stream) {
    if (!h) return pq_set_error(PEUQSE_ERR_INVALID, "bad argument");
    k_ess_tune<<<1, 1, 0, (cudaStream_t)stream>>>(h->D.st);
    PQ_CUDA_OK(cudaGetLastError());
    return PEUQSE_OK;
}

extern "C" int peuqse_essr_info(peuqse_essr* h, double* d_info, void* stream) {
    if (!h || !d_info) return pq_set_error(PEUQSE_ERR_INVALID, "bad argument");
    k_ess_info<<<1, 1, 0, (cudaStream_t)stream>>>(h->D.st, d_info);
    PQ_CUDA_OK(cudaGetLastError());
    return PEUQSE_OK;
}

// the whole run for a built-in model (any response covariance): every round's block goes through peuqse_logp_batch
int pq_ess_run_rounds(const peuqse_problem* p, const peuqse_ess_spec* s, double* d_x, double* d_logp_x, double* d_chain,
                      double* d_chain_logp, double* d_info, cudaStream_t st) {
    peuqse_essr* h = nullptr;
    int rc = peuqse_essr_create(s, &h);
    if (rc) return rc;
    const long long W = s->n_walkers;
    const int k = s->k;
    double* cand = nullptr;
    cudaError_t e = cudaMallocAsync((void**)&cand, (size_t)W * (k + 1) * sizeof(double), st);
    if (e != cudaSuccess) { peuqse_essr_destroy(h); return pq_set_cuda_error(e, "cudaMallocAsync(ess candidates)", __LINE__); }
    double* rows = cand + (size_t)W * k;
    for (long long it = 0; it < s->n_iters && rc == PEUQSE_OK; ++it) {
        for (int half = 0; half < 2 && rc == PEUQSE_OK; ++half) {
            int64_t n = 0;
            rc = peuqse_essr_begin_half(h, it, half, d_x, d_logp_x, cand, &n, st);
            while (rc == PEUQSE_OK && n > 0) {
                rc = peuqse_logp_batch(p, cand, n, rows, nullptr, nullptr, st);
                if (rc == PEUQSE_OK) rc = peuqse_essr_round(h, rows, cand, &n, st);
            }
            if (rc == PEUQSE_OK) rc = peuqse_essr_end_half(h, d_x, d_logp_x, st);
        }
        if (rc != PEUQSE_OK) break;
        rc = peuqse_essr_end_iteration(h, st);
        if (d_chain && cudaMemcpyAsync(d_chain + it * W * k, d_x, (size_t)W * k * sizeof(double), cudaMemcpyDeviceToDevice, st) != cudaSuccess)
            rc = pq_set_error(PEUQSE_ERR_CUDA, "copy of the ensemble into the chain failed");
        if (d_chain_logp && cudaMemcpyAsync(d_chain_logp + it * W, d_logp_x, (size_t)W * sizeof(double), cudaMemcpyDeviceToDevice, st) != cudaSuccess)
            rc = pq_set_error(PEUQSE_ERR_CUDA, "copy of the log-posteriors into the chain failed");
    }
    if (rc == PEUQSE_OK && d_info) rc = peuqse_essr_info(h, d_info, st);
    cudaFreeAsync(cand, st);
    cudaStreamSynchronize(st);          // the handle's state must outlive the kernels that use it
    peuqse_essr_destroy(h);
    return rc;
}
