// Definitions shared by the two ensemble-slice drivers: csrc/ess.cu (one thread / warp owns a walker's whole slice update)
// and csrc/ess_rounds.cu (round-synchronous: every stepping-out / shrink round is one batched log-posterior launch).
// Both consume the SAME Philox draws -- key = seed, stream PQ_STREAM_ESS, counter index ess_idx(iteration, half, walker):
//   block 0xffffffff (half 0): shuffle key of the walker     block 0: the two complementary walkers of the direction
//   block 1: Exp(1) of the slice level, left end L           block 2: J (stepping-out budget to the left)
//   block 3 + r/2: the uniform of shrink round r (words x,y for even r, z,w for odd r)
// so that a walker's trajectory does not depend on the driver (tests/test_gpu_ess.py; oracle/peuqse_oracle.py: ess_run_philox).
#pragma once
#include "common.cuh"
#include "philox.cuh"

#define ESS_MAXSTEPS 10000.0          // zeus default `maxsteps`: bound on the stepping-out procedure
#define ESS_TOLERANCE 0.05            // zeus defaults for the tuning of mu
#define ESS_PATIENCE 5

struct EssState {                     // lives at the start of d_work
    double mu;
    unsigned long long nexp, ncon;    // of the current iteration
    unsigned long long neval, tot_exp, tot_con;
    int tune, ncount, error, pad;
};

struct DevEss {
    long long W, first_iter;
    unsigned long long seed;
    int k, max_rounds;
};

__device__ __forceinline__ unsigned long long ess_idx(const DevEss& E, long long it, int half, long long w) {
    return ((unsigned long long)((E.first_iter + it) * 2 + half) << 32) | (unsigned long long)w;
}


// random permutation of the walkers: rank of a Philox key (ties broken by index)
static __global__ void k_ess_shuffle(const DevEss E, long long it, int* __restrict__ perm, unsigned long long* __restrict__ keys) {
    const long long w = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (w >= E.W) return;
    const Philox4 r = philox_block(E.seed, PQ_STREAM_ESS, ess_idx(E, it, 0, w), 0xffffffffu);
    keys[w] = ((unsigned long long)r.x << 32) | r.y;
}
static __global__ void k_ess_rank(const DevEss E, const unsigned long long* __restrict__ keys, int* __restrict__ perm) {
    const long long w = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (w >= E.W) return;
    const unsigned long long kw = keys[w];
    int rank = 0;
    for (long long v = 0; v < E.W; ++v) {
        const unsigned long long kv = keys[v];
        rank += (kv < kw) || (kv == kw && v < w);
    }
    perm[rank] = (int)w;
}

// after both halves: adapt mu (zeus' tuning rule), fold the counters, record the ensemble
static __global__ void k_ess_tune(EssState* __restrict__ st) {
    if (threadIdx.x || blockIdx.x) return;
    const unsigned long long ne = st->nexp, nc = st->ncon;
    st->tot_exp += ne;
    st->tot_con += nc;
    if (st->tune) {
        const double nexp = ne < 1 ? 1.0 : (double)ne;
        const double frac = nexp / (nexp + (double)nc);
        st->mu *= 2.0 * frac;
        if (fabs(frac - 0.5) < ESS_TOLERANCE) st->ncount += 1;
        if (st->ncount > ESS_PATIENCE) st->tune = 0;
    }
    st->nexp = 0;
    st->ncon = 0;
}

static __global__ void k_ess_init(EssState* st, double mu0, int tune) {
    st->mu = mu0; st->nexp = 0; st->ncon = 0; st->neval = 0; st->tot_exp = 0; st->tot_con = 0;
    st->tune = tune; st->ncount = 0; st->error = 0; st->pad = 0;
}

static __global__ void k_ess_info(const EssState* st, double* info) {
    info[0] = (double)st->neval;
    info[1] = st->mu;
    info[2] = (double)st->tot_exp;
    info[3] = (double)st->tot_con;
    info[4] = (double)st->error;
}

