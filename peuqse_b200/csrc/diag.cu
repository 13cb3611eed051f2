// Geweke convergence diagnostic for thousands of chains: geweke_diagnostic and the
// window / chain / parameter loops around it in calculateAndPlotConvergenceDiagnostics (IP.py:3960-3975, 4012-4062).
// The reference evaluates, for each of 20 growing windows of every chain and parameter, 20 z-scores that compare the mean
// of an early slice with the mean of the last half, each variance being the order-2 Yule-Walker spectral density at zero
// (statsmodels.yule_walker, method 'adjusted', demeaned): ~95 passes over every chain in numpy.
// Every quantity of a slice [a, b) -- sum x, sum x^2, lag-1 and lag-2 product sums -- is a difference of PREFIX sums, and
// the 1600 slice ends of a chain fall on ~1300 distinct rows.  So:
//   k_geweke_prefix   ONE streaming pass over samples[R][C*k] (coalesced across columns, rows cut into segments for
//                     parallelism, 4 rows in flight per thread): running sums about the column's first value, recorded at
//                     the listed rows only;
//   k_geweke_carry    turns the segments' totals into what has to be added to a record of each segment (tiny);
//   k_geweke_z        one thread per (z-score, column): slice statistics from six records each, the 2x2 Yule-Walker solve,
//                     the z-score.
// HBM traffic: 8 B per sample read once + 32 B x positions x columns of records.
#include <algorithm>
#include <vector>
#include "common.cuh"

namespace {

#define GW_THREADS 128
#define GW_U 8

struct GwGeom {
    long long R, CK, seg_len;
    int n_seg, nb;
};

// prefix sums over rows u < p of x_u, x_u^2, x_{u-1} x_u, x_{u-2} x_u (x taken about the column's first value)
__global__ void __launch_bounds__(GW_THREADS) k_geweke_prefix(const double* __restrict__ x, const GwGeom G,
                                                              const long long* __restrict__ pos, const int* __restrict__ seg_first,
                                                              double* __restrict__ rec, double* __restrict__ tot) {
    const long long col = (long long)blockIdx.x * GW_THREADS + threadIdx.x;
    const int s = blockIdx.y;
    if (col >= G.CK) return;
    const long long t0 = (long long)s * G.seg_len, t1 = min(G.R, t0 + G.seg_len);
    const double pivot = x[col];
    double f0 = 0.0, f2 = 0.0, h1 = 0.0, h2 = 0.0;
    double p1 = (t0 >= 1) ? x[(t0 - 1) * G.CK + col] - pivot : 0.0;       // x_{t-1}
    double p2 = (t0 >= 2) ? x[(t0 - 2) * G.CK + col] - pivot : 0.0;       // x_{t-2}
    int j = seg_first[s];
    const int jend = seg_first[s + 1];
    const size_t qs = (size_t)G.nb * G.CK;
    auto record = [&]() {                                                 // the sums over the rows before the current one
        const size_t o = (size_t)j * G.CK + col;
        rec[o] = f0; rec[qs + o] = f2; rec[2 * qs + o] = h1; rec[3 * qs + o] = h2;
        ++j;
    };
    // rows are walked in runs that end at the next recorded row, so the hot loop carries no position test: per element one
    // load, one subtraction and four accumulations (x_{-1} = x_{-2} = 0 makes the first two lag products vanish)
    const double* px = x + t0 * G.CK + col;
    long long t = t0;
    while (t < t1) {
        while (j < jend && pos[j] == t) record();
        const long long stop = (j < jend && pos[j] < t1) ? pos[j] : t1;
        for (; t + GW_U <= stop; t += GW_U) {
            double v[GW_U];
#pragma unroll
            for (int u = 0; u < GW_U; ++u) v[u] = px[(size_t)u * G.CK];
            px += (size_t)GW_U * G.CK;
#pragma unroll
            for (int u = 0; u < GW_U; ++u) {
                const double xv = v[u] - pivot;
                f0 += xv;
                f2 = fma(xv, xv, f2);
                h1 = fma(p1, xv, h1);
                h2 = fma(p2, xv, h2);
                p2 = p1; p1 = xv;
            }
        }
        for (; t < stop; ++t, px += G.CK) {
            const double xv = *px - pivot;
            f0 += xv;
            f2 = fma(xv, xv, f2);
            h1 = fma(p1, xv, h1);
            h2 = fma(p2, xv, h2);
            p2 = p1; p1 = xv;
        }
    }
    while (j < jend) record();                                            // position R: sums over every row
    const size_t ts = (size_t)G.n_seg * G.CK, o = (size_t)s * G.CK + col;
    tot[o] = f0; tot[ts + o] = f2; tot[2 * ts + o] = h1; tot[3 * ts + o] = h2;
}

// totals of the segments -> exclusive prefix over segments, in place: what has to be added to a record of segment s
__global__ void __launch_bounds__(GW_THREADS) k_geweke_carry(const GwGeom G, double* __restrict__ tot) {
    const long long col = (long long)blockIdx.x * GW_THREADS + threadIdx.x;
    const int q = blockIdx.y;
    if (col >= G.CK) return;
    double carry = 0.0;
    double* tq = tot + (size_t)q * G.n_seg * G.CK;
    for (int s = 0; s < G.n_seg; ++s) {
        const double v = tq[(size_t)s * G.CK + col];
        tq[(size_t)s * G.CK + col] = carry;
        carry += v;
    }
}

struct SliceStat {
    double mean, sv, n;
};

// sl: indices into pos[] of the rows a, a+1, a+2, b-2, b-1, b of the slice [a, b)
__device__ __forceinline__ SliceStat slice_stat(const double* __restrict__ rec, const double* __restrict__ carry, const GwGeom& G,
                                                const long long* __restrict__ pos, const int* __restrict__ sl, long long col) {
    const size_t qs = (size_t)G.nb * G.CK, ts = (size_t)G.n_seg * G.CK;
    // prefix value = the record (sums since its segment's first row) + the totals of the segments before it
    auto RQ = [&](int q, int i) {
        const long long sg = min(pos[sl[i]] / G.seg_len, (long long)G.n_seg - 1);
        return rec[q * qs + (size_t)sl[i] * G.CK + col] + carry[q * ts + (size_t)sg * G.CK + col];
    };
    auto R0 = [&](int i) { return RQ(0, i); };
    SliceStat S;
    const double n = (double)(pos[sl[5]] - pos[sl[0]]);
    S.n = n;
    const double nan = __longlong_as_double(0x7ff8000000000000LL);
    if (n < 3.0) { S.mean = nan; S.sv = nan; return S; }
    const double fa = R0(0), fa1 = R0(1), fa2 = R0(2), fb2 = R0(3), fb1 = R0(4), fb = R0(5);
    const double s1 = fb - fa, m = s1 / n;
    const double q0 = RQ(1, 5) - RQ(1, 0);
    const double l1 = RQ(2, 5) - RQ(2, 1);                                // sum_{v=a+1}^{b-1} x_{v-1} x_v
    const double l2 = RQ(3, 5) - RQ(3, 2);                                // sum_{v=a+2}^{b-1} x_{v-2} x_v
    // demeaned autocovariances, statsmodels' 'adjusted' denominators n - k
    const double r0 = (q0 - n * m * m) / n;
    const double r1 = (l1 - m * ((fb1 - fa) + (fb - fa1)) + (n - 1.0) * m * m) / (n - 1.0);
    const double r2 = (l2 - m * ((fb2 - fa) + (fb - fa2)) + (n - 2.0) * m * m) / (n - 2.0);
    const double det = r0 * r0 - r1 * r1;
    const double rho1 = (r0 * r1 - r1 * r2) / det, rho2 = (r0 * r2 - r1 * r1) / det;
    const double sigmasq = r0 - (r1 * rho1 + r2 * rho2);
    const double d = 1.0 - (rho1 + rho2);
    S.mean = m;
    S.sv = (sigmasq > 0.0) ? sigmasq / (d * d) : nan;                     // sigma is NaN unless sigma^2 > 0
    return S;
}

__global__ void __launch_bounds__(GW_THREADS) k_geweke_z(const double* __restrict__ rec, const double* __restrict__ carry, const GwGeom G,
                                                         const long long* __restrict__ pos, const int* __restrict__ slices,
                                                         int nz, double* __restrict__ z) {
    const long long col = (long long)blockIdx.x * GW_THREADS + threadIdx.x;
    const int zi = blockIdx.y;
    if (col >= G.CK) return;
    const SliceStat A = slice_stat(rec, carry, G, pos, slices + (size_t)zi * 12, col);
    const SliceStat B = slice_stat(rec, carry, G, pos, slices + (size_t)zi * 12 + 6, col);
    z[(size_t)zi * G.CK + col] = (A.mean - B.mean) / sqrt(A.sv / A.n + B.sv / B.n);     // IP.py:4056-4057
}

int gw_sm_count() {
    static int sm = 0;
    if (!sm) {
        int dev = 0;
        cudaGetDevice(&dev);
        if (cudaDeviceGetAttribute(&sm, cudaDevAttrMultiProcessorCount, dev) != cudaSuccess || sm < 1) sm = 148;
    }
    return sm;
}

// rows are cut into as many segments as give ONE resident wave of CTAs (16 CTAs of 128 threads per SM at 32 registers): with
// 2496 CTAs on 2368 slots the kernel ran a second, nearly empty wave
GwGeom gw_geom(long long R, long long CK, int nb) {
    GwGeom G;
    G.R = R; G.CK = CK; G.nb = nb;
    const long long gx = (CK + GW_THREADS - 1) / GW_THREADS;
    long long want = ((long long)gw_sm_count() * 16) / gx;
    want = std::max(1LL, std::min(want, 256LL));
    long long len = std::max(64LL, (R + want - 1) / want);
    len = (len + GW_U - 1) / GW_U * GW_U;
    G.seg_len = len;
    G.n_seg = (int)std::max(1LL, (R + len - 1) / len);
    return G;
}


// Integrated autocorrelation time (zeus.autocorr.AutoCorrTime, method 'mk', called at IP.py:3916-3930): the walkers' chains
// of one parameter are laid end to end into ONE series whose autocorrelation function is taken by FFT.  This kernel is the
// layout change in front of the transform: samples[t][c][d] (row = step) -> series_d[c * n + t] - mean_d for every
// parameter in one read of the first n rows, through a 32 x 64 shared-memory tile so that both the reads (along c*k+d)
// and the writes (along t) are coalesced.  HBM traffic: 8 B read + 8 B written per sample.
#define AC_TT 32
#define AC_TC 64
__global__ void __launch_bounds__(256) k_act_concat(const double* __restrict__ x, long long n, long long C, int k,
                                                    const double* __restrict__ mean, double* __restrict__ out,
                                                    long long out_stride, unsigned col_tiles) {
    __shared__ double tile[AC_TT][AC_TC + 1];
    const long long CK = C * k;
    const long long col0 = (long long)(blockIdx.x % col_tiles) * AC_TC, t0 = (long long)(blockIdx.x / col_tiles) * AC_TT;
    {
        const int tx = threadIdx.x % AC_TC, ty = threadIdx.x / AC_TC;
        const long long col = col0 + tx;
#pragma unroll
        for (int r = ty; r < AC_TT; r += 256 / AC_TC) {
            const long long t = t0 + r;
            tile[r][tx] = (col < CK && t < n) ? x[t * CK + col] : 0.0;
        }
    }
    __syncthreads();
    const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
    const long long t = t0 + lane;
    // chain and parameter of the tile's first column by one 64-bit division, of the others by 32-bit steps from it (ncu: the
    // per-column 64-bit division made the kernel issue-bound, 73 % of the issue slots at 63 % of the DRAM throughput)
    const long long c0 = col0 / k;
    const int d0 = (int)(col0 - c0 * k);
    for (int j = w; j < AC_TC; j += 8) {
        if (col0 + j >= CK || t >= n) continue;
        const int q = (d0 + j) / k, d = (d0 + j) - q * k;
        out[(long long)d * out_stride + (c0 + q) * n + t] = tile[lane][j] - mean[d];
    }
}

}  // namespace

extern "C" int64_t peuqse_geweke_work_bytes(int64_t n_rows, int64_t n_cols, int32_t n_positions) {
    const GwGeom G = gw_geom(n_rows, n_cols, n_positions);
    return (int64_t)(4 * ((size_t)n_positions + G.n_seg) * n_cols * sizeof(double) + ((size_t)G.n_seg + 1) * sizeof(int) + 256);
}

extern "C" int peuqse_geweke(const double* d_samples, int64_t n_rows, int64_t n_cols, const int64_t* h_positions,
                             const int64_t* d_positions, int32_t n_positions, const int32_t* d_slices, int32_t n_z, double* d_z,
                             void* d_work, void* stream) {
    if (!d_samples || !h_positions || !d_positions || !d_slices || !d_z || !d_work) return pq_set_error(PEUQSE_ERR_INVALID, "null argument");
    if (n_rows < 1 || n_cols < 1 || n_positions < 1 || n_z < 1) return pq_set_error(PEUQSE_ERR_INVALID, "bad sizes");
    for (int j = 0; j < n_positions; ++j)
        if (h_positions[j] < 0 || h_positions[j] > n_rows || (j && h_positions[j] <= h_positions[j - 1]))
            return pq_set_error(PEUQSE_ERR_INVALID, "positions must be strictly increasing rows in [0, n_rows]");
    const GwGeom G = gw_geom(n_rows, n_cols, n_positions);
    cudaStream_t st = (cudaStream_t)stream;
    double* rec = (double*)d_work;
    double* tot = rec + 4 * (size_t)n_positions * n_cols;
    int* seg_first = (int*)(tot + 4 * (size_t)G.n_seg * n_cols);
    // first position of every segment (position n_rows belongs to the last one)
    std::vector<int> sf(G.n_seg + 1);
    int j = 0;
    for (int s = 0; s < G.n_seg; ++s) {
        const long long t0 = (long long)s * G.seg_len;
        while (j < n_positions && h_positions[j] < t0) ++j;
        sf[s] = j;
    }
    sf[G.n_seg] = n_positions;
    PQ_CUDA_OK(cudaMemcpyAsync(seg_first, sf.data(), sf.size() * sizeof(int), cudaMemcpyHostToDevice, st));
    PQ_CUDA_OK(cudaStreamSynchronize(st));                                 // sf is a stack-lifetime host buffer
    const unsigned gx = (unsigned)((n_cols + GW_THREADS - 1) / GW_THREADS);
    k_geweke_prefix<<<dim3(gx, G.n_seg), GW_THREADS, 0, st>>>(d_samples, G, (const long long*)d_positions, seg_first, rec, tot);
    k_geweke_carry<<<dim3(gx, 4), GW_THREADS, 0, st>>>(G, tot);
    k_geweke_z<<<dim3(gx, n_z), GW_THREADS, 0, st>>>(rec, tot, G, (const long long*)d_positions, d_slices, n_z, d_z);
    PQ_CUDA_OK(cudaGetLastError());
    return PEUQSE_OK;
}

extern "C" int peuqse_act_concat(const double* d_samples, int64_t n_rows, int64_t n_chains, int32_t k, const double* d_mean,
                                 double* d_series, int64_t series_stride, void* stream) {
    if (!d_samples || !d_mean || !d_series) return pq_set_error(PEUQSE_ERR_INVALID, "null argument");
    if (n_rows < 1 || n_chains < 1 || k < 1 || series_stride < n_rows * n_chains) return pq_set_error(PEUQSE_ERR_INVALID, "bad sizes");
    cudaStream_t st = (cudaStream_t)stream;
    const long long len = n_rows * n_chains, CK = n_chains * k;
    const long long gx = (CK + AC_TC - 1) / AC_TC, gy = (n_rows + AC_TT - 1) / AC_TT;
    if (gx * gy > 2147483647LL) return pq_set_error(PEUQSE_ERR_INVALID, "more than 2^31 tiles");
    if (series_stride > len)                                               // zero padding of the transform
        for (int d = 0; d < k; ++d)
            PQ_CUDA_OK(cudaMemsetAsync(d_series + (size_t)d * series_stride + len, 0, (size_t)(series_stride - len) * sizeof(double), st));
    k_act_concat<<<(unsigned)(gx * gy), 256, 0, st>>>(d_samples, n_rows, n_chains, k, d_mean, d_series, series_stride, (unsigned)gx);
    PQ_CUDA_OK(cudaGetLastError());
    return PEUQSE_OK;
}
