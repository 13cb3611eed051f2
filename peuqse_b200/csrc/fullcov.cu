// Full-covariance likelihood (kernel K5 of SURVEY section 2): for a block of B simulated response rows,
//   maha_b = || (y - sim_b) @ U ||^2,   ll_b += -0.5*(rank*log(2pi) + log_pdet + maha_b)
// with U = u*sqrt(1/s) from scipy's _PSD (scipy/stats/_multivariate.py:208-236, 590-591) hoisted to set-up -- or, when the
// caller hands over an upper-triangular whitening (the transposed inverse Cholesky factor), only its non-zero half.
// Replaces one multivariate_normal.logpdf (an n x n eigendecomposition) per point (IP.py:2960).
//
// T[B,n] = Rsd[B,n] @ U[n,n] runs on the fp64 tensor pipe (mma.sync.m8n8k4.f64 = DMMA; there is no tcgen05 kind for fp64)
// and is never written for large batches: the row-wise sum of squares is fused into the contraction.  Three shapes:
//   k_qf_tile<false>  enough tiles to fill the device: one CTA per 64 x 128 tile of T, 4-stage cp.async ring,
//                     8 warps x (32 x 32) accumulators, fused sum of squares;
//   k_qf_tile<true>   16 < B, fewer tiles than SMs x 2: the same tile per (tile, K chunk) CTA, chunk sums to memory; row
//                     fragments beyond the batch issue no DMMA, so a padded tile costs what its rows cost;
//   k_qf_skinny<MF>   B <= 16 (many rounds of an ensemble-slice half-sweep): GEMV-shaped -- every warp streams a
//                     16-column panel of one K chunk of U exactly once with 16-byte loads straight into DMMA B fragments,
//                     the (pre-arranged) residual fragments come from L1/L2; bound by the one read of U.
// All three define a row's value the same way, so that it is bit-identical whatever batch the row is evaluated in (the
// sharded ensemble-slice rounds and the pinning tests rely on it):
//   T[b][j]   = ((c0 + c1) + c2) + c3, c_i = DMMA accumulation over the i-th quarter of the column tile's K range
//               (k ascending, four k per instruction);
//   leaf      = fma-accumulated squares of the 8 values T[b][128 t + 32 w + 8 j + 2 q + e], j < 4, e < 2  (j, then e, ascending)
//   quad      = (leaf_0 + leaf_1) + (leaf_2 + leaf_3) over q;  tile_t = ((quad_0 + quad_1) + quad_2) + quad_3 over w;
//   maha_b    = tile_0 + tile_1 + ... (t ascending).
#include <cstdint>
#include <cstdlib>
#include "common.cuh"
#include "eval.cuh"
#include "launch.cuh"

int pq_logp_external(const peuqse_problem* p, const double* th, const double* sim, long long B, double* logp, double* lpr,
                     double* llk, cudaStream_t st);

#define QT_BM 64
#define QT_BN 128
#define QT_BK 16
#define QT_LDA (QT_BK + 4)      // padded leading dimensions (doubles; B operand: columns + 8): fragment loads touch every bank exactly twice,
                                // rows stay 16-byte aligned for cp.async
#define QT_STAGES 4
#define QT_CHUNKS 4
// a CTA of 2 x WN warps owns 64 rows x 32 WN columns of T (WN = 4: a whole canonical 128-column tile; WN = 2: half of one)
template <int WN> struct QtCfg {
    static constexpr int BN = 32 * WN, THREADS = 64 * WN, LDB = BN + 8;
    static constexpr int STAGE_DOUBLES = QT_BM * QT_LDA + QT_BK * LDB;
    static constexpr int SMEM_BYTES = QT_STAGES * STAGE_DOUBLES * 8;
};

struct QfResp {
    const double* U;            // padded whitening: rows n_pad (multiple of 16), leading dimension ldu (multiple of 128), zero filled
    int n, n_pad, ldu, off;     // off: first entry of the response in the problem's observation order
    int poff;                   // first column of the response in the padded residual block
    int tri, ntile;             // upper-triangular U; column tiles of 128
    double c;                   // rank*log(2pi) + log_pdet
};
struct QfBatch {
    int nresp, NP;              // full-covariance responses; columns of the padded residual block (sum of n_pad)
    QfResp r[PQ_MAXR];
};

__device__ __forceinline__ void dmma884(double& c0, double& c1, double a, double b) {
    asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};"
                 : "+d"(c0), "+d"(c1)
                 : "d"(a), "d"(b));
}
__device__ __forceinline__ void cp_async16(void* smem_dst, const void* gmem_src) {
    const unsigned s = (unsigned)__cvta_generic_to_shared(smem_dst);
    asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(s), "l"(gmem_src) : "memory");
}
__device__ __forceinline__ void cp_async_commit() { asm volatile("cp.async.commit_group;" ::: "memory"); }
template <int N>
__device__ __forceinline__ void cp_async_wait() { asm volatile("cp.async.wait_group %0;" ::"n"(N) : "memory"); }

// number of 16-row K tiles that column tile t of response R needs (all of them, or the non-zero part of a triangular U)
__device__ __forceinline__ int qf_ktiles(const QfResp& R, int t) {
    const int all = R.n_pad / QT_BK;
    if (!R.tri) return all;
    const int need = (t + 1) * (QT_BN / QT_BK);
    return need < all ? need : all;
}

// ---------------------------------------------------------------------------------------------------
// residuals y - sim (x - mean, IP.py:2960) of every full-covariance response, zero padded:
//   tile form  Rm[B_pad][NP]   (B_pad multiple of 64)                           -> A operand of k_qf_tile via cp.async
//   fragment form Af[resp][k/4][MF][32]: lane l of fragment m holds R[8 m + l/4][4 (k/4) + l%4]  -> A operand of k_qf_skinny
// ---------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256) k_qf_residual(const QfBatch Q, const double* __restrict__ sim, int ld_sim,
                                                    const int* __restrict__ src, const double* __restrict__ y, long long B,
                                                    long long B_pad, double* __restrict__ Rm) {
    const long long b = blockIdx.x;
    for (int c = blockIdx.y * blockDim.x + threadIdx.x; c < Q.NP; c += gridDim.y * blockDim.x) {
        int r = 0;
        while (r + 1 < Q.nresp && c >= Q.r[r + 1].poff) ++r;
        const int i = c - Q.r[r].poff;
        double v = 0.0;
        if (b < B && i < Q.r[r].n) v = y[Q.r[r].off + i] - sim[b * ld_sim + src[Q.r[r].off + i]];
        Rm[b * Q.NP + c] = v;
    }
}

template <int MF>
__global__ void __launch_bounds__(256) k_qf_residual_frag(const QfBatch Q, const double* __restrict__ sim, int ld_sim,
                                                         const int* __restrict__ src, const double* __restrict__ y, long long B,
                                                         double* __restrict__ Af) {
    // one thread per (k4 step, fragment, lane) of the concatenated padded responses
    const long long total = (long long)(Q.NP / 4) * MF * 32;
    for (long long e = (long long)blockIdx.x * blockDim.x + threadIdx.x; e < total; e += (long long)gridDim.x * blockDim.x) {
        const int lane = (int)(e & 31);
        const int m = (int)((e >> 5) % MF);
        const int c = (int)((e >> 5) / MF) * 4 + (lane & 3);       // column of the padded block
        const long long b = m * 8 + (lane >> 2);
        int r = 0;
        while (r + 1 < Q.nresp && c >= Q.r[r + 1].poff) ++r;
        const int i = c - Q.r[r].poff;
        double v = 0.0;
        if (b < B && i < Q.r[r].n) v = y[Q.r[r].off + i] - sim[b * ld_sim + src[Q.r[r].off + i]];
        Af[e] = v;
    }
}

// The same two forms for the built-in LINEAR model when every response has a full covariance (config C4): the simulated
// value sum_d Phi[j][d] theta_d (d ascending, multiply then add: the operation order of Sim<PEUQSE_MODEL_LINEAR>, eval.cuh)
// is formed here instead of being written to and read back from a [B][n] block by a separate simulation kernel.  A thread
// keeps its observation's row of Phi in registers (read from the transposed copy: coalesced) and walks over 8 rows of the
// batch (thetas in shared memory).
#define QR_ROWS 8
__global__ void __launch_bounds__(256) k_qf_residual_lin(const QfBatch Q, const double* __restrict__ theta, int k,
                                                        const double* __restrict__ PhiT, int n_total,
                                                        const double* __restrict__ y, long long B, long long B_pad,
                                                        double* __restrict__ Rm) {
    __shared__ double s_th[QR_ROWS][PQ_MAXK];
    const long long b0 = (long long)blockIdx.x * QR_ROWS;
    for (int e = threadIdx.x; e < QR_ROWS * k; e += blockDim.x) {
        const int r = e / k, d = e - r * k;
        s_th[r][d] = (b0 + r < B) ? theta[(b0 + r) * k + d] : 0.0;
    }
    __syncthreads();
    for (int c = blockIdx.y * blockDim.x + threadIdx.x; c < Q.NP; c += gridDim.y * blockDim.x) {
        int r = 0;
        while (r + 1 < Q.nresp && c >= Q.r[r + 1].poff) ++r;
        const int i = c - Q.r[r].poff;
        const bool in = i < Q.r[r].n;
        double row[PQ_MAXK];
        const int j = Q.r[r].off + (in ? i : 0);
#pragma unroll
        for (int d = 0; d < PQ_MAXK; ++d) row[d] = (d < k) ? PhiT[(size_t)d * n_total + j] : 0.0;   // neighbours read neighbours
        const double yj = y[j];
#pragma unroll
        for (int q = 0; q < QR_ROWS; ++q) {
            if (b0 + q >= B_pad) break;
            double acc = row[0] * s_th[q][0];
#pragma unroll
            for (int d = 1; d < PQ_MAXK; ++d)
                if (d < k) acc = acc + row[d] * s_th[q][d];
            Rm[(b0 + q) * Q.NP + c] = (in && b0 + q < B) ? (yj - acc) : 0.0;
        }
    }
}

template <int MF>
__global__ void __launch_bounds__(256) k_qf_residual_frag_lin(const QfBatch Q, const double* __restrict__ theta, int k,
                                                             const double* __restrict__ Phi, const double* __restrict__ y,
                                                             long long B, double* __restrict__ Af) {
    const long long total = (long long)(Q.NP / 4) * MF * 32;
    for (long long e = (long long)blockIdx.x * blockDim.x + threadIdx.x; e < total; e += (long long)gridDim.x * blockDim.x) {
        const int lane = (int)(e & 31);
        const int m = (int)((e >> 5) % MF);
        const int c = (int)((e >> 5) / MF) * 4 + (lane & 3);
        const long long b = m * 8 + (lane >> 2);
        int r = 0;
        while (r + 1 < Q.nresp && c >= Q.r[r + 1].poff) ++r;
        const int i = c - Q.r[r].poff;
        double v = 0.0;
        if (b < B && i < Q.r[r].n) {
            const int j = Q.r[r].off + i;
            const double* row = Phi + (size_t)j * k;
            const double* th = theta + b * k;
            double acc = row[0] * th[0];
            for (int d = 1; d < k; ++d) acc = acc + row[d] * th[d];
            v = y[j] - acc;
        }
        Af[e] = v;
    }
}

// ---------------------------------------------------------------------------------------------------
// tile kernel
// ---------------------------------------------------------------------------------------------------
// SPLIT = false: blockIdx.x = tile (column tile major, heaviest first), blockIdx.z = response; the CTA walks the four K
// chunks one after the other and writes partial[resp][column tile][row].  SPLIT = true: blockIdx.z = 4 * response + chunk;
// the chunk's accumulators go to Tc[chunk][row][column] and k_qf_finish_tc combines them -- the same additions in the same
// order.
template <bool SPLIT, int WN>
__global__ void __launch_bounds__(64 * WN, (SPLIT || WN == 2) ? 2 : 1) k_qf_tile(const QfBatch Q, const double* __restrict__ Rm, long long B,
                                                                      long long B_pad, int nrt, double* __restrict__ partial,
                                                                      double* __restrict__ Tc) {
    extern __shared__ __align__(16) double sm[];
    const int resp = SPLIT ? (int)blockIdx.z / QT_CHUNKS : (int)blockIdx.z;
    const QfResp& R = Q.r[resp];
    using Cfg = QtCfg<WN>;
    constexpr int BN = Cfg::BN, THREADS = Cfg::THREADS, LDB = Cfg::LDB, STAGE_DOUBLES = Cfg::STAGE_DOUBLES;
    const int cb = R.ntile * (QT_BN / BN) - 1 - (int)(blockIdx.x / nrt);   // column block; heavy (long K range) ones first
    if (cb < 0) return;
    const int ct = cb * BN / QT_BN, w0 = (cb * BN % QT_BN) / 32;    // canonical column tile and first warp column in it
    const int rt = (int)(blockIdx.x % nrt);
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int wm = warp / WN, wn = warp % WN;                      // 2 x WN warps, 32 x 32 of the tile each
    const long long bm = (long long)rt * QT_BM;
    const int bn = cb * BN;
    const int nkt = qf_ktiles(R, ct);
    int kt_first = 0, kt_last = nkt;
    if (SPLIT) {
        const int ch = (int)blockIdx.z % QT_CHUNKS;
        kt_first = (int)((long long)nkt * ch / QT_CHUNKS);
        kt_last = (int)((long long)nkt * (ch + 1) / QT_CHUNKS);
    }
    // row fragments of this warp that hold rows of the batch (the last row tile is padded): the others issue no DMMA
    int mfv = (int)((B - bm - wm * 32 + 7) / 8);
    mfv = mfv < 0 ? 0 : (mfv > 4 ? 4 : mfv);
    // cp.async sources of this thread as running pointers (one 16-byte chunk of A per 32 rows, of B per 256 / BN * 2 rows)
    constexpr int A_CP = 512 / THREADS, B_CP = 8 * BN / THREADS, A_ROWS = THREADS / 8, B_ROWS = THREADS / (BN / 2);
    const double* a_src = Rm + (bm + (tid >> 3)) * Q.NP + R.poff + (tid & 7) * 2 + (size_t)kt_first * QT_BK;
    const double* b_src = R.U + bn + (size_t)(kt_first * QT_BK + tid / (BN / 2)) * R.ldu + (tid % (BN / 2)) * 2;
    const size_t a_rstride = (size_t)A_ROWS * Q.NP, b_rstride = (size_t)B_ROWS * R.ldu, b_kstride = (size_t)QT_BK * R.ldu;
    const int a_dst = (tid >> 3) * QT_LDA + (tid & 7) * 2;
    const int b_dst = QT_BM * QT_LDA + (tid / (BN / 2)) * LDB + (tid % (BN / 2)) * 2;
    auto issue_a = [&](int stage) {                                // A: 64 rows x 8 chunks of 16 bytes
        double* S = sm + (size_t)stage * STAGE_DOUBLES + a_dst;
#pragma unroll
        for (int i = 0; i < A_CP; ++i) cp_async16(S + i * A_ROWS * QT_LDA, a_src + i * a_rstride);
    };
    auto issue_b = [&](int stage, int i0, int i1) {                // B: 16 rows x BN / 2 chunks
        double* S = sm + (size_t)stage * STAGE_DOUBLES + b_dst;
#pragma unroll
        for (int i = i0; i < i1; ++i) cp_async16(S + i * B_ROWS * LDB, b_src + i * b_rstride);
    };
    // operand fragments of k4 step ks of the tile in `stage`
    const int a_off = (wm * 32 + (lane >> 2)) * QT_LDA + (lane & 3);
    const int b_off = QT_BM * QT_LDA + (lane & 3) * LDB + wn * 32 + (lane >> 2);
    auto frags = [&](int stage, int ks, double (&af)[4], double (&bf)[4]) {
        const double* S = sm + (size_t)stage * STAGE_DOUBLES;
#pragma unroll
        for (int i = 0; i < 4; ++i) af[i] = S[a_off + i * 8 * QT_LDA + ks];
#pragma unroll
        for (int j = 0; j < 4; ++j) bf[j] = S[b_off + ks * LDB + j * 8];
    };
    double acc[4][4][2];
#pragma unroll
    for (int i = 0; i < 4; ++i)
#pragma unroll
        for (int j = 0; j < 4; ++j) acc[i][j][0] = acc[i][j][1] = 0.0;
    double tot[SPLIT ? 1 : 4][SPLIT ? 1 : 4][2];
    auto fold = [&](int chunk) {                                   // chunk boundary: tot = chunk sums added in order
        if (!SPLIT) {
#pragma unroll
            for (int i = 0; i < 4; ++i)
#pragma unroll
                for (int j = 0; j < 4; ++j) {
                    tot[SPLIT ? 0 : i][SPLIT ? 0 : j][0] = chunk == 0 ? acc[i][j][0] : tot[SPLIT ? 0 : i][SPLIT ? 0 : j][0] + acc[i][j][0];
                    tot[SPLIT ? 0 : i][SPLIT ? 0 : j][1] = chunk == 0 ? acc[i][j][1] : tot[SPLIT ? 0 : i][SPLIT ? 0 : j][1] + acc[i][j][1];
                    acc[i][j][0] = acc[i][j][1] = 0.0;
                }
        }
    };
    // ring of 4 stages.  The one barrier per tile sits in the MIDDLE of the tile's four k4 steps: behind it every warp has
    // left tile kt-1 (its stage is refilled with tile kt+3) and tile kt+1 has landed for everybody, so that the first
    // fragments of tile kt+1 are fetched during the last step of tile kt and the tensor pipe never waits at a tile boundary.
    // The fragments of the step behind the barrier are in registers when the barrier is reached, and the copies of tile
    // kt+3 are issued in three portions BEHIND the DMMAs of steps 1, 2 and 3, so that both warps of a scheduler never leave
    // the tensor pipe without work at the same time.
#pragma unroll
    for (int s = 0; s < QT_STAGES - 1; ++s) {
        if (kt_first + s < kt_last) { issue_a(s); issue_b(s, 0, B_CP); }
        cp_async_commit();
        a_src += QT_BK;
        b_src += b_kstride;
    }
    int chunk = 0;
    int chunk_end = SPLIT ? kt_last : (int)((long long)nkt / QT_CHUNKS);
    double af[2][4], bf[2][4];
    cp_async_wait<1>();                                            // tiles 0 and 1 of this thread ...
    __syncthreads();                                               // ... and of everybody
    if (kt_first < kt_last) frags(0, 0, af[0], bf[0]);
    for (int kt = kt_first; kt < kt_last; ++kt) {
        const int stage = (kt - kt_first) % QT_STAGES, nstage = (stage + QT_STAGES - 1) % QT_STAGES;
        const bool more = kt + QT_STAGES - 1 < kt_last;
        if (!SPLIT) {
            while (kt >= chunk_end) {
                fold(chunk);
                ++chunk;
                chunk_end = (int)((long long)nkt * (chunk + 1) / QT_CHUNKS);
            }
        }
#pragma unroll
        for (int s = 0; s < QT_BK / 4; ++s) {
            const int cur = s & 1;
            if (s == 1) {
                cp_async_wait<1>();                                // tile kt+1 has landed (tile kt+2 may still be in flight)
                __syncthreads();
            }
            if (s + 1 < QT_BK / 4) frags(stage, (s + 1) * 4, af[cur ^ 1], bf[cur ^ 1]);
            else if (kt + 1 < kt_last) frags((stage + 1) % QT_STAGES, 0, af[cur ^ 1], bf[cur ^ 1]);
#pragma unroll
            for (int i = 0; i < 4; ++i)
                if (i < mfv) {
#pragma unroll
                    for (int j = 0; j < 4; ++j) dmma884(acc[i][j][0], acc[i][j][1], af[cur][i], bf[cur][j]);
                }
            if (more) {
                if (s == 1) issue_a(nstage);
                if (s == 2) issue_b(nstage, 0, B_CP / 2);
                if (s == 3) issue_b(nstage, B_CP / 2, B_CP);
            }
        }
        cp_async_commit();
        a_src += QT_BK;
        b_src += b_kstride;
    }
    cp_async_wait<0>();
    if (SPLIT) {
        const int ch = (int)blockIdx.z % QT_CHUNKS;
        double* out = Tc + ((size_t)ch * B_pad + bm) * Q.NP + R.poff + bn;
#pragma unroll
        for (int i = 0; i < 4; ++i)
#pragma unroll
            for (int j = 0; j < 4; ++j) {
                const int row = wm * 32 + i * 8 + (lane >> 2), col = wn * 32 + j * 8 + (lane & 3) * 2;
                *reinterpret_cast<double2*>(out + (size_t)row * Q.NP + col) = make_double2(acc[i][j][0], acc[i][j][1]);
            }
        return;
    }
    // remaining chunk sums (the last chunk, and empty chunks of very short K ranges)
    for (; chunk < QT_CHUNKS; ++chunk) fold(chunk);
    // quad sums of this warp's 32 columns: partial[global column tile][warp column][row]
    double* pq = partial + (((size_t)(R.poff / QT_BN) + ct) * 4 + w0 + wn) * B;
#pragma unroll
    for (int i = 0; i < 4; ++i) {
        double s = 0.0;
#pragma unroll
        for (int j = 0; j < 4; ++j) {
            s = fma(tot[SPLIT ? 0 : i][SPLIT ? 0 : j][0], tot[SPLIT ? 0 : i][SPLIT ? 0 : j][0], s);
            s = fma(tot[SPLIT ? 0 : i][SPLIT ? 0 : j][1], tot[SPLIT ? 0 : i][SPLIT ? 0 : j][1], s);
        }
        s += __shfl_xor_sync(0xffffffffu, s, 1);
        s += __shfl_xor_sync(0xffffffffu, s, 2);
        const long long row = bm + wm * 32 + i * 8 + (lane >> 2);
        if ((lane & 3) == 0 && row < B) pq[row] = s;
    }
}

// ---------------------------------------------------------------------------------------------------
// skinny kernel: B <= 8 MF rows
// ---------------------------------------------------------------------------------------------------
template <int MF>
__global__ void __launch_bounds__(256) k_qf_skinny(const QfBatch Q, const double* __restrict__ Af, long long B_pad,
                                                  double* __restrict__ Tc) {
    const int resp = blockIdx.z, ch = blockIdx.y, t = blockIdx.x;
    const QfResp& R = Q.r[resp];
    if (t >= R.ntile) return;
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const int nkt = qf_ktiles(R, t);
    const int s0 = (int)((long long)nkt * ch / QT_CHUNKS) * (QT_BK / 4), s1 = (int)((long long)nkt * (ch + 1) / QT_CHUNKS) * (QT_BK / 4);
    const int col0 = t * QT_BN + warp * 16;
    // lane (kq = lane % 4, c = lane / 4) reads U[4 s + kq][col0 + 2 c .. + 1]: .x feeds the B fragment of the even columns,
    // .y the one of the odd columns (the assignment of columns to fragment positions is free)
    const double2* up = reinterpret_cast<const double2*>(R.U + (size_t)(lane & 3) * R.ldu + col0 + 2 * (lane >> 2));
    const size_t ustep = (size_t)4 * R.ldu / 2;                    // double2 per k4 step
    const double* ap = Af + ((size_t)(R.poff / 4) * MF) * 32 + lane;
    double acc[MF][2][2];
#pragma unroll
    for (int m = 0; m < MF; ++m) acc[m][0][0] = acc[m][0][1] = acc[m][1][0] = acc[m][1][1] = 0.0;
    // the panel is streamed in groups of UN k4 steps, two groups in flight: the loads of group g+1 are issued before the
    // DMMAs of group g, so that 8 to 16 sixteen-byte loads per lane are outstanding all the time (the kernel is bound by
    // the latency-bandwidth product of this one read of U)
    constexpr int UN = 8;
    auto load_group = [&](int s_, double2 (&u)[UN]) {
#pragma unroll
        for (int q = 0; q < UN; ++q) {
            const int sq = (s_ + q < s1) ? s_ + q : s1 - 1;           // the tail group re-reads the last step (its DMMAs are skipped)
            const double2* a = up + (size_t)sq * ustep;
            asm volatile("ld.global.nc.L1::no_allocate.v2.f64 {%0,%1}, [%2];" : "=d"(u[q].x), "=d"(u[q].y) : "l"(a));
        }
    };
    auto mma_group = [&](int s_, const double2 (&u)[UN]) {
#pragma unroll
        for (int q = 0; q < UN; ++q) {
            if (s_ + q < s1) {
#pragma unroll
                for (int m = 0; m < MF; ++m) {
                    const double a = __ldg(ap + ((size_t)(s_ + q) * MF + m) * 32);
                    dmma884(acc[m][0][0], acc[m][0][1], a, u[q].x);
                    dmma884(acc[m][1][0], acc[m][1][1], a, u[q].y);
                }
            }
        }
    };
    if (s0 < s1) {
        double2 ua[UN], ub[UN];
        load_group(s0, ua);
        for (int s = s0; s < s1; s += 2 * UN) {
            if (s + UN < s1) load_group(s + UN, ub);
            mma_group(s, ua);
            if (s + 2 * UN < s1) load_group(s + 2 * UN, ua);
            if (s + UN < s1) mma_group(s + UN, ub);
        }
    }
    // accumulator (row = lane / 4, fragment column 2 (lane % 4) + e) of fragment f is column col0 + 4 (lane % 4) + 2 e + f
    double* out = Tc + (size_t)ch * B_pad * Q.NP + R.poff + col0 + 4 * (lane & 3);
#pragma unroll
    for (int m = 0; m < MF; ++m) {
        double* o = out + (size_t)(m * 8 + (lane >> 2)) * Q.NP;
        *reinterpret_cast<double2*>(o) = make_double2(acc[m][0][0], acc[m][1][0]);
        *reinterpret_cast<double2*>(o + 2) = make_double2(acc[m][0][1], acc[m][1][1]);
    }
}

// chunk sums Tc[4][B_pad][NP] -> partial[global column tile][warp column][row]: combination of the chunks and the quad sums
// of k_qf_tile's epilogue.  grid (column tile, group of 8 rows, response), 128 threads = (warp column w, row, q)
__global__ void __launch_bounds__(128) k_qf_finish_tc(const QfBatch Q, const double* __restrict__ Tc, long long B, long long B_pad,
                                                     double* __restrict__ partial) {
    const int resp = blockIdx.z, t = blockIdx.x;
    const QfResp& R = Q.r[resp];
    if (t >= R.ntile) return;
    const int tid = threadIdx.x, w = tid >> 5, rsub = (tid & 31) >> 2, q = tid & 3;
    const long long row = (long long)blockIdx.y * 8 + rsub;
    const size_t cs = (size_t)B_pad * Q.NP;
    const double* base = Tc + (size_t)row * Q.NP + R.poff + t * QT_BN + w * 32 + q * 2;
    double s = 0.0;
#pragma unroll
    for (int j = 0; j < 4; ++j) {
        const double2 c0 = *reinterpret_cast<const double2*>(base + j * 8);
        const double2 c1 = *reinterpret_cast<const double2*>(base + cs + j * 8);
        const double2 c2 = *reinterpret_cast<const double2*>(base + 2 * cs + j * 8);
        const double2 c3 = *reinterpret_cast<const double2*>(base + 3 * cs + j * 8);
        const double x0 = ((c0.x + c1.x) + c2.x) + c3.x, x1 = ((c0.y + c1.y) + c2.y) + c3.y;
        s = fma(x0, x0, s);
        s = fma(x1, x1, s);
    }
    s += __shfl_xor_sync(0xffffffffu, s, 1);
    s += __shfl_xor_sync(0xffffffffu, s, 2);
    if (q == 0 && row < B) partial[(((size_t)(R.poff / QT_BN) + t) * 4 + w) * B + row] = s;
}

// ll_b += -0.5*(c_r + sum_t tile_t) for every full-covariance response in order, then the NaN convention and logP.
// One WARP per row: the lanes fetch and combine the quad sums of the column tiles in parallel (the loads are what takes
// time), lane 0 adds the tile sums in ascending tile order -- the fixed order every kernel shape shares.
__global__ void __launch_bounds__(256) k_qf_reduce_finish(const QfBatch Q, const double* __restrict__ partial, long long B,
                                                         const double* __restrict__ lp, double* __restrict__ ll,
                                                         double* __restrict__ logp) {
    __shared__ double s_tile[8][128];
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const long long b = (long long)blockIdx.x * 8 + warp;
    if (b >= B) return;
    double l = ll[b];
    for (int r = 0; r < Q.nresp; ++r) {
        const double* part = partial + (size_t)(Q.r[r].poff / QT_BN) * 4 * B + b;
        const int nt = Q.r[r].ntile;
        double m = 0.0;
        for (int t0 = 0; t0 < nt; t0 += 128) {                               // 128 tiles (16384 columns) per round
            const int nn = nt - t0 < 128 ? nt - t0 : 128;
            for (int t = lane; t < nn; t += 32) {
                const double* q = part + (size_t)(t0 + t) * 4 * B;
                s_tile[warp][t] = ((q[0] + q[B]) + q[2 * B]) + q[3 * B];
            }
            __syncwarp();
            if (lane == 0)
                for (int t = 0; t < nn; ++t) m += s_tile[warp][t];
            __syncwarp();
        }
        m = __shfl_sync(0xffffffffu, m, 0);
        l = l + (-0.5 * (Q.r[r].c + m));
    }
    if (lane == 0) {
        if (l != l) l = pq_neg_inf();            // NaN simulated values -> -inf (IP.py:2880-2882)
        ll[b] = l;
        if (logp) logp[b] = combine_logp(lp[b], l);
    }
}

namespace {
template <int MF>
void launch_skinny(const QfBatch& Q, int ntile_max, const double* sim, int ld_sim, const int* src, const double* y, long long B,
                   long long B_pad, double* Af, double* Tc, int sm_count, cudaStream_t st, const double* theta, int k,
                   const double* Phi) {
    const long long total = (long long)(Q.NP / 4) * MF * 32;
    const int rb = (int)std::min<long long>((total + 255) / 256, (long long)sm_count * 8);
    if (!sim) k_qf_residual_frag_lin<MF><<<rb, 256, 0, st>>>(Q, theta, k, Phi, y, B, Af);
    else k_qf_residual_frag<MF><<<rb, 256, 0, st>>>(Q, sim, ld_sim, src, y, B, Af);
    k_qf_skinny<MF><<<dim3((unsigned)ntile_max, QT_CHUNKS, (unsigned)Q.nresp), 256, 0, st>>>(Q, Af, B_pad, Tc);
}
}  // namespace

static int g_qf_wn = 4;      // warps along N of the tile kernel's CTA (PEUQSE_QF_WN=2: 64 x 64 CTAs of 4 warps, two per SM)

// rows [0, B) of one chunk: residuals, contraction, partial sums, likelihood
static int fullcov_rows(const peuqse_problem* p, const QfBatch& Q, int ntile_max, int ntiles, const double* d_sim, long long B,
                        const double* lp, double* ll, double* d_logp, cudaStream_t st, const double* d_theta) {
    static bool attr_done = false;
    if (!attr_done) {
        cudaFuncSetAttribute((const void*)k_qf_tile<true, 4>, cudaFuncAttributeMaxDynamicSharedMemorySize, QtCfg<4>::SMEM_BYTES);
        cudaFuncSetAttribute((const void*)k_qf_tile<false, 4>, cudaFuncAttributeMaxDynamicSharedMemorySize, QtCfg<4>::SMEM_BYTES);
        cudaFuncSetAttribute((const void*)k_qf_tile<true, 2>, cudaFuncAttributeMaxDynamicSharedMemorySize, QtCfg<2>::SMEM_BYTES);
        cudaFuncSetAttribute((const void*)k_qf_tile<false, 2>, cudaFuncAttributeMaxDynamicSharedMemorySize, QtCfg<2>::SMEM_BYTES);
        const char* e = getenv("PEUQSE_QF_WN");
        if (e) g_qf_wn = atoi(e) == 2 ? 2 : 4;
        attr_done = true;
    }
    const bool skinny = B <= 16;
    const int MF = B <= 8 ? 1 : 2;
    const long long B_pad = skinny ? 8LL * MF : ((B + QT_BM - 1) / QT_BM) * QT_BM;
    const int nrt = (int)(B_pad / QT_BM);
    // fewer tiles than two per SM: one CTA per (tile, K chunk) instead of a long K loop on a part of the device
    const bool split = !skinny && (long long)nrt * ntiles < 2LL * p->sm_count;
    const size_t n_part = (size_t)ntiles * 4 * B, n_res = (size_t)B_pad * Q.NP,
                 n_tc = (skinny || split) ? (size_t)QT_CHUNKS * B_pad * Q.NP : 0;
    double* scratch = nullptr;
    PQ_CUDA_OK(cudaMallocAsync((void**)&scratch, (n_part + n_res + n_tc + 8) * sizeof(double), st));
    double* part = scratch;
    double* res = (double*)(((uintptr_t)(part + n_part) + 15) & ~(uintptr_t)15);
    double* tc = res + n_res;
    const int ld_sim = p->d.n_total;
    if (skinny) {
        if (MF == 1) launch_skinny<1>(Q, ntile_max, d_sim, ld_sim, p->d.src, p->d.y, B, B_pad, res, tc, p->sm_count, st, d_theta, p->d.k, p->d.mconst);
        else launch_skinny<2>(Q, ntile_max, d_sim, ld_sim, p->d.src, p->d.y, B, B_pad, res, tc, p->sm_count, st, d_theta, p->d.k, p->d.mconst);
        k_qf_finish_tc<<<dim3((unsigned)ntile_max, (unsigned)(B_pad / 8), (unsigned)Q.nresp), 128, 0, st>>>(Q, tc, B, B_pad, part);
    } else {
        if (!d_sim)
            k_qf_residual_lin<<<dim3((unsigned)((B_pad + QR_ROWS - 1) / QR_ROWS), (unsigned)((Q.NP + 255) / 256)), 256, 0, st>>>(
                Q, d_theta, p->d.k, p->mconstT, p->d.n_total, p->d.y, B, B_pad, res);
        else
            k_qf_residual<<<dim3((unsigned)B_pad, (unsigned)std::min((Q.NP + 255) / 256, 16)), 256, 0, st>>>(Q, d_sim, ld_sim, p->d.src,
                                                                                                           p->d.y, B, B_pad, res);
        if (split) {
            if (g_qf_wn == 4)
                k_qf_tile<true, 4><<<dim3((unsigned)(ntile_max * nrt), 1, (unsigned)(Q.nresp * QT_CHUNKS)), 256, QtCfg<4>::SMEM_BYTES, st>>>(
                    Q, res, B, B_pad, nrt, part, tc);
            else
                k_qf_tile<true, 2><<<dim3((unsigned)(2 * ntile_max * nrt), 1, (unsigned)(Q.nresp * QT_CHUNKS)), 128, QtCfg<2>::SMEM_BYTES, st>>>(
                    Q, res, B, B_pad, nrt, part, tc);
            k_qf_finish_tc<<<dim3((unsigned)ntile_max, (unsigned)(B_pad / 8), (unsigned)Q.nresp), 128, 0, st>>>(Q, tc, B, B_pad, part);
        } else {
            if (g_qf_wn == 4)
                k_qf_tile<false, 4><<<dim3((unsigned)(ntile_max * nrt), 1, (unsigned)Q.nresp), 256, QtCfg<4>::SMEM_BYTES, st>>>(
                    Q, res, B, B_pad, nrt, part, nullptr);
            else
                k_qf_tile<false, 2><<<dim3((unsigned)(2 * ntile_max * nrt), 1, (unsigned)Q.nresp), 128, QtCfg<2>::SMEM_BYTES, st>>>(
                    Q, res, B, B_pad, nrt, part, nullptr);
        }
    }
    k_qf_reduce_finish<<<(unsigned)((B + 7) / 8), 256, 0, st>>>(Q, part, B, lp, ll, d_logp);
    const cudaError_t e = cudaGetLastError();
    cudaFreeAsync(scratch, st);
    if (e != cudaSuccess) return pq_set_cuda_error(e, "full-covariance contraction", __LINE__);
    return PEUQSE_OK;
}

// d_sim == nullptr: built-in LINEAR model whose responses all have full covariances -- the residual kernels simulate (d_theta
// is needed then)
int pq_fullcov_logp(const peuqse_problem* p, const double* d_theta, const double* d_sim, long long B, double* d_logp,
                    double* d_logprior, double* d_loglik, cudaStream_t st) {
    if (!d_sim && (!d_theta || p->d.model_id != PEUQSE_MODEL_LINEAR)) return pq_set_error(PEUQSE_ERR_INVALID, "fullcov: no simulated values");
    // the padded residual block places every response at a multiple of 128 columns (poff / 128 = its first global column tile)
    QfBatch Q;
    Q.nresp = 0;
    int poff = 0, ntile_max = 0, ntiles = 0;
    for (int r = 0; r < p->d.n_resp; ++r) {
        if (p->d.resp_mode[r] != PEUQSE_RESP_MVN_FULL) continue;
        QfResp& R = Q.r[Q.nresp++];
        R.U = p->d.U[r];
        R.n = p->d.resp_n[r];
        R.n_pad = p->U_rows[r];
        R.ldu = p->U_ld[r];
        R.off = p->d.resp_off[r];
        R.poff = poff;
        R.tri = p->U_tri[r];
        R.ntile = R.ldu / QT_BN;
        R.c = p->d.resp_c[r];
        poff += R.ldu;
        ntile_max = std::max(ntile_max, R.ntile);
        ntiles += R.ntile;
    }
    Q.NP = poff;
    double* ll = d_loglik;
    double* lp = d_logprior;
    double* tmp = nullptr;
    const bool need_lp = d_logp != nullptr || d_logprior != nullptr;
    if (!ll || (need_lp && !lp)) {
        PQ_CUDA_OK(cudaMallocAsync((void**)&tmp, (size_t)2 * B * sizeof(double), st));
        if (!ll) ll = tmp;
        if (need_lp && !lp) lp = tmp + B;
    }
    // prior and the diagonal / per-value responses; full responses contribute 0 here
    int rc = pq_logp_external(p, need_lp ? d_theta : nullptr, d_sim, B, nullptr, need_lp ? lp : nullptr, ll, st);
    // row chunks bound the residual block to 1 GiB (a row's value does not depend on the chunk it is evaluated in)
    long long rows = ((1LL << 27) / Q.NP) / QT_BM * QT_BM;
    if (rows < QT_BM) rows = QT_BM;
    for (long long b0 = 0; b0 < B && rc == PEUQSE_OK; b0 += rows) {
        const long long nb = std::min(rows, B - b0);
        rc = fullcov_rows(p, Q, ntile_max, ntiles, d_sim ? d_sim + (size_t)b0 * p->d.n_total : nullptr, nb, lp ? lp + b0 : nullptr,
                          ll + b0, d_logp ? d_logp + b0 : nullptr, st, d_theta ? d_theta + (size_t)b0 * p->d.k : nullptr);
    }
    if (tmp) cudaFreeAsync(tmp, st);
    return rc;
}
