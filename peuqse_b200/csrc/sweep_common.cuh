// Sweep point generation shared by sweep.cu (generic kernels) and sweep_grid.cu (specialised grid kernels).
#pragma once
#include "common.cuh"
#include "philox.cuh"

#define PQ_SWEEP_MAX_BLOCKS 2048
#define PQ_SWEEP_TICKET_BYTES 256
#define PQ_ACC_MAX PEUQSE_ACC_DOUBLES(PQ_MAXK)

// Shortlist of a sweep whose logP is not kept in HBM (sweep.cu): while pass 1 runs, every point whose logP is within
// PQ_SHORT_MARGIN of the best value seen so far (a lower bound of the final maximum, shared through one 64-bit
// atomicMax key) is appended as (index, logP).  The final threshold max + log(10^-coef) is >= best-so-far - margin
// whenever -log_rel <= margin, so the list is a superset of the retained points and passes 2-3 read it instead of
// evaluating every point again.  Header at d_work + PQ_SHORT_HDR_OFFSET (inside the 256-byte ticket area, never
// overlapping the grid kernels' aux data); entries at d_work + PQ_SWEEP_BASE_BYTES (that region is only used by
// the grid kernels when logP IS stored).
#define PQ_SHORT_HDR_OFFSET 64
#define PQ_SHORT_VALID_WORD ((PQ_SHORT_HDR_OFFSET + 16) / 4)
#define PQ_SHORT_CAP (1LL << 20)
#define PQ_SHORT_RCAP 2048
#define PQ_SHORT_MARGIN 50.0
struct ShortHdr {
    unsigned long long gmax_key;   // order-preserving key of the best logP published so far (0 = none)
    unsigned long long count;      // entries appended; > PQ_SHORT_CAP means overflow (list unusable)
    unsigned int valid;            // pass 1 finished in shortlist mode on this work buffer
    unsigned int full_pass2;       // set by k_short_pass2: the statistics must be recomputed from all points
};

struct FastDiv {            // unsigned 64-bit division by an invariant (Granlund & Montgomery 1994, fig. 4.1)
    unsigned long long m;
    int sh1, sh2;
};

struct DevSweep {
    int kind, k;
    double center[PQ_MAXK], interval[PQ_MAXK], stdp[PQ_MAXK];
    unsigned long long r2[PQ_MAXK];       // 2*radius
    unsigned long long mprod[PQ_MAXK + 1];  // prod_{d<p}(2 r_d + 1)
    FastDiv div[PQ_MAXK];
    double spread;
    unsigned long long seed;
};

static inline FastDiv make_fastdiv(unsigned long long d) {
    FastDiv f;
    if (d == 0) d = 1;
    int l = 0;
    while (l < 64 && (((unsigned __int128)1) << l) < d) ++l;
    const unsigned __int128 two64 = ((unsigned __int128)1) << 64;
    const unsigned __int128 num = two64 * ((((unsigned __int128)1) << l) - d);
    f.m = (unsigned long long)(num / d + 1);
    f.sh1 = l < 1 ? l : 1;
    f.sh2 = l > 1 ? l - 1 : 0;
    return f;
}

__device__ __forceinline__ unsigned long long fastdiv(unsigned long long n, const FastDiv& f) {
    const unsigned long long t = __umul64hi(f.m, n);
    return (t + ((n - t) >> f.sh1)) >> f.sh2;
}

static inline int build_sweep(const peuqse_sweep_spec* s, DevSweep* d, long long* total) {
    if (!s || s->k < 1 || s->k > PQ_MAXK) return pq_set_error(PEUQSE_ERR_INVALID, "bad sweep spec");
    d->kind = s->kind;
    d->k = s->k;
    d->spread = s->spread;
    d->seed = s->seed;
    unsigned __int128 m = 1;
    d->mprod[0] = 1;
    for (int p = 0; p < PQ_MAXK; ++p) {
        const bool on = p < s->k;
        d->center[p] = on ? s->center[p] : 0.0;
        d->interval[p] = on ? s->interval[p] : 0.0;
        d->stdp[p] = on ? s->std_prior[p] : 0.0;
        const long long r = on ? s->radii[p] : 0;
        if (r < 0) return pq_set_error(PEUQSE_ERR_INVALID, "negative grid radius");
        d->r2[p] = 2ULL * (unsigned long long)r;
        d->div[p] = make_fastdiv(d->r2[p]);
        m *= (unsigned __int128)(d->r2[p] + 1);
        if (m >> 62) return pq_set_error(PEUQSE_ERR_INVALID, "grid has more than 2^62 points");
        d->mprod[p + 1] = (unsigned long long)m;
    }
    if (total) *total = (s->kind == 0) ? (long long)d->mprod[s->k] : (long long)0x3fffffffffffffffLL;
    return PEUQSE_OK;
}

// element g of combinationGenerator's list: from the last axis down; point value c -/+ (h*i)
// (CombinationGeneratorModule.py:59, 71: one multiply, one add)
// kd = number of leading axes to decode (S.k for a full point; S.k-1 decodes a row index of the last axis' bulk)
template <int KT>
__device__ __forceinline__ void grid_point(const DevSweep& S, unsigned long long g, double (&th)[KT], int kd) {
#pragma unroll
    for (int p = KT - 1; p >= 0; --p) {
        th[p] = 0.0;
        if (p < kd) {
            const unsigned long long M = S.mprod[p];
            if (g < M) {
                th[p] = S.center[p];
            } else {
                const unsigned long long gp = g - M;
                const unsigned long long b = fastdiv(gp, S.div[p]);
                const unsigned long long j = gp - b * S.r2[p];
                const double step = S.interval[p] * (double)((j >> 1) + 1ULL);
                th[p] = (j & 1ULL) ? (S.center[p] + step) : (S.center[p] - step);
                g = b;
            }
        }
    }
}

// generateInitialPoints('uniform') (IP.py:706, 749): ((spread*(-2+4u))*std) + center, u = Philox(seed, g)
template <int KT>
__device__ __forceinline__ void uniform_point(const DevSweep& S, unsigned long long g, double (&th)[KT]) {
#pragma unroll
    for (int b = 0; b < (KT + 1) / 2; ++b) {
        if (2 * b < S.k) {
            const Philox4 r = philox_block(S.seed, PQ_STREAM_UNIFORM_SWEEP, g, (uint32_t)b);
            const double ua = philox_u53(r.x, r.y), ub = philox_u53(r.z, r.w);
            th[2 * b] = (S.spread * (-2.0 + 4.0 * ua)) * S.stdp[2 * b] + S.center[2 * b];
            if (2 * b + 1 < KT) {
                th[2 * b + 1] = 0.0;
                if (2 * b + 1 < S.k) th[2 * b + 1] = (S.spread * (-2.0 + 4.0 * ub)) * S.stdp[2 * b + 1] + S.center[2 * b + 1];
            }
        } else {
            th[2 * b] = 0.0;
            if (2 * b + 1 < KT) th[2 * b + 1] = 0.0;
        }
    }
}

template <int KT>
__device__ __forceinline__ void sweep_point(const DevSweep& S, unsigned long long g, double (&th)[KT]) {
    if (S.kind == 0) grid_point<KT>(S, g, th, S.k);
    else uniform_point<KT>(S, g, th);
}

