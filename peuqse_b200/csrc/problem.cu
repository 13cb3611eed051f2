// Problem handle: validates the host descriptor and uploads the constants once (replaces the per-call
// re-derivation in the reference: scipy's _PSD eigendecomposition on every logpdf call, IP.py:2773, 2960).
#include <cstdio>
#include <cstring>
#include <string>
#include <vector>
#include "common.cuh"
#include "eval.cuh"

static thread_local std::string g_err;

int pq_set_error(int code, const char* msg) {
    g_err = msg;
    return code;
}
int pq_set_cuda_error(cudaError_t e, const char* what, int line) {
    char buf[512];
    snprintf(buf, sizeof buf, "CUDA error %d (%s) at %s [line %d]", (int)e, cudaGetErrorString(e), what, line);
    g_err = buf;
    return PEUQSE_ERR_CUDA;
}

extern "C" const char* peuqse_last_error(void) { return g_err.c_str(); }
extern "C" int peuqse_version(void) { return 100; }

extern "C" int peuqse_device_info(int32_t* out3) {
    int dev = 0;
    PQ_CUDA_OK(cudaGetDevice(&dev));
    cudaDeviceProp prop;
    PQ_CUDA_OK(cudaGetDeviceProperties(&prop, dev));
    out3[0] = prop.multiProcessorCount;
    out3[1] = (int32_t)prop.sharedMemPerBlockOptin;
    out3[2] = prop.major * 10 + prop.minor;
    return PEUQSE_OK;
}

size_t pq_obs_smem_bytes(const DevProblem& d) { return pq_obs_doubles(d) * sizeof(double); }

namespace {
struct BlobBuilder {
    std::vector<unsigned char> host;
    size_t add(const void* src, size_t bytes) {
        size_t off = (host.size() + 255) & ~(size_t)255;   // 256 B alignment for vector loads / TMA
        host.resize(off + bytes);
        if (bytes) memcpy(host.data() + off, src, bytes);
        return off;
    }
};
}  // namespace

extern "C" int peuqse_problem_create(const peuqse_problem_desc* h, peuqse_problem** out) {
    if (!h || !out) return pq_set_error(PEUQSE_ERR_INVALID, "null argument");
    if (h->k < 1 || h->k > PQ_MAXK) return pq_set_error(PEUQSE_ERR_INVALID, "k must be in [1, PEUQSE_MAX_K]");
    if (h->n_resp < 1 || h->n_resp > PQ_MAXR) return pq_set_error(PEUQSE_ERR_INVALID, "n_resp must be in [1, PEUQSE_MAX_RESP]");
    if (h->kp < 0 || h->kp > h->k) return pq_set_error(PEUQSE_ERR_INVALID, "kp out of range");
    if (h->prior_mode != PEUQSE_PRIOR_NONE && h->kp == 0) return pq_set_error(PEUQSE_ERR_INVALID, "prior_mode needs kp > 0");
    if (h->model_id < PEUQSE_MODEL_EXTERNAL || h->model_id > PEUQSE_MODEL_LANGMUIR_LOG)
        return pq_set_error(PEUQSE_ERR_INVALID, "unknown model_id");

    {
        // the stream-ordered scratch of the launch functions (cudaMallocAsync) comes from the device's default pool, whose
        // release threshold of 0 hands freed memory back to the driver at every synchronisation: samplers that synchronise
        // once per round (ensemble slice sampling) then paid ~80 us per allocation.  Keep the pool's memory.
        int dev = 0;
        cudaMemPool_t pool;
        if (cudaGetDevice(&dev) == cudaSuccess && cudaDeviceGetDefaultMemPool(&pool, dev) == cudaSuccess) {
            unsigned long long keep = ~0ull;
            cudaMemPoolSetAttribute(pool, cudaMemPoolAttrReleaseThreshold, &keep);
        }
        cudaGetLastError();
    }
    peuqse_problem* p = new peuqse_problem();
    memset(p, 0, sizeof *p);
    DevProblem& d = p->d;
    d.k = h->k;
    d.kp = h->kp;
    d.scaling_on = h->scaling_on;
    d.prior_mode = h->prior_mode;
    d.has_lb = h->has_lb;
    d.has_ub = h->has_ub;
    d.prior_c = h->prior_c;
    for (int i = 0; i < PQ_MAXK; ++i) {
        d.lb[i] = -INFINITY; d.ub[i] = INFINITY; d.scale[i] = 1.0; d.mus[i] = 0.0; d.gidx[i] = i; d.pw[i] = 0.0;
    }
    for (int i = 0; i < h->k; ++i) {
        if (h->has_lb) d.lb[i] = h->lb[i];
        if (h->has_ub) d.ub[i] = h->ub[i];
        if (h->scaling_on) d.scale[i] = h->scale[i];
        d.mus[i] = h->mu_scaled[i];
    }
    d.prior_identity = 1;
    for (int j = 0; j < h->kp; ++j) {
        d.gidx[j] = h->prior_gidx[j];
        if (d.gidx[j] < 0 || d.gidx[j] >= h->k) { delete p; return pq_set_error(PEUQSE_ERR_INVALID, "prior_gidx out of range"); }
        if (d.gidx[j] != j) d.prior_identity = 0;
        if (h->prior_mode == PEUQSE_PRIOR_DIAG) d.pw[j] = h->prior_w[j];
    }
    if (h->prior_mode == PEUQSE_PRIOR_DENSE && !d.prior_identity) {
        delete p;
        return pq_set_error(PEUQSE_ERR_INVALID, "dense prior needs all parameters Gaussian in natural order (IP.py:151-152)");
    }
    d.n_resp = h->n_resp;
    d.n_total = h->n_total;
    d.model_id = h->model_id;
    d.n_mconst = h->n_model_const;                       // tpd2site: + n reciprocals appended below
    int off = 0;
    for (int r = 0; r < h->n_resp; ++r) {
        d.resp_n[r] = h->resp_n[r];
        d.resp_mode[r] = h->resp_mode[r];
        d.resp_off[r] = off;
        d.resp_c[r] = h->resp_c[r];
        if (h->resp_mode[r] == PEUQSE_RESP_PER_VALUE) d.any_per_value = 1;
        if (h->resp_mode[r] == PEUQSE_RESP_MVN_FULL) d.any_full = 1;
        off += h->resp_n[r];
    }
    if (off != h->n_total) { delete p; return pq_set_error(PEUQSE_ERR_INVALID, "n_total != sum(resp_n)"); }
    // model constant sizes
    {
        long need = 0;
        if (h->model_id == PEUQSE_MODEL_QUADRATIC) need = h->n_total;
        if (h->model_id == PEUQSE_MODEL_TPD2SITE) need = 3L * h->n_total;
        if (h->model_id == PEUQSE_MODEL_LANGMUIR_LOG) need = 3L * h->n_total;
        if (h->model_id == PEUQSE_MODEL_LINEAR) need = (long)h->n_total * h->k;
        if (need != h->n_model_const) { delete p; return pq_set_error(PEUQSE_ERR_INVALID, "n_model_const does not match the model"); }
        if (h->model_id == PEUQSE_MODEL_QUADRATIC && h->k != 2) { delete p; return pq_set_error(PEUQSE_ERR_INVALID, "quadratic model has k = 2"); }
        if (h->model_id == PEUQSE_MODEL_LANGMUIR_LOG && h->k != 2) { delete p; return pq_set_error(PEUQSE_ERR_INVALID, "langmuir_log model has k = 2"); }
        if (h->model_id == PEUQSE_MODEL_TPD2SITE && h->k != 6) { delete p; return pq_set_error(PEUQSE_ERR_INVALID, "tpd2site model has k = 6"); }
    }

    BlobBuilder bb;
    const size_t nt = (size_t)h->n_total;
    std::vector<double> zeros(nt, 0.0);
    std::vector<int32_t> ident(nt);
    for (size_t i = 0; i < nt; ++i) ident[i] = (int32_t)i;
    const size_t o_y = bb.add(h->y, nt * 8);
    const size_t o_w = bb.add(h->w ? h->w : zeros.data(), nt * 8);
    const size_t o_c = bb.add(h->cval ? h->cval : zeros.data(), nt * 8);
    const size_t o_s = bb.add(h->src ? (const void*)h->src : (const void*)ident.data(), nt * 4);
    size_t o_m;
    if (h->model_id == PEUQSE_MODEL_TPD2SITE) {
        // kBT, cov1, cov2 as given, then the correctly rounded reciprocals 1/kBT that eval.cuh's pq_div_by needs
        std::vector<double> mc(4 * nt);
        for (size_t i = 0; i < 3 * nt; ++i) mc[i] = h->model_const[i];
        for (size_t i = 0; i < nt; ++i) mc[3 * nt + i] = 1.0 / h->model_const[i];
        o_m = bb.add(mc.data(), mc.size() * 8);
        d.n_mconst = (int)(4 * nt);
    } else {
        o_m = bb.add(h->model_const, (size_t)h->n_model_const * 8);
    }
    size_t o_mT = 0;
    if (h->model_id == PEUQSE_MODEL_LINEAR) {
        std::vector<double> t((size_t)h->k * nt);
        for (size_t j = 0; j < nt; ++j)
            for (int d = 0; d < h->k; ++d) t[(size_t)d * nt + j] = h->model_const[j * h->k + d];
        o_mT = bb.add(t.data(), t.size() * 8);
    }
    size_t o_pU = 0;
    if (h->prior_mode == PEUQSE_PRIOR_DENSE) o_pU = bb.add(h->prior_U, (size_t)h->kp * h->kp * 8);
    size_t o_U[PQ_MAXR] = {0};
    for (int r = 0; r < h->n_resp; ++r)
        if (h->resp_mode[r] == PEUQSE_RESP_MVN_FULL) {
            if (!h->resp_U || !h->resp_U[r]) { delete p; return pq_set_error(PEUQSE_ERR_INVALID, "MVN_FULL response without resp_U"); }
            const size_t n = (size_t)h->resp_n[r], rows = (n + 15) / 16 * 16, ld = (n + 127) / 128 * 128;
            std::vector<double> pad(rows * ld, 0.0);
            int tri = 1;
            for (size_t i = 0; i < n; ++i) {
                const double* row = h->resp_U[r] + i * n;
                memcpy(pad.data() + i * ld, row, n * 8);
                for (size_t j = 0; j < i && tri; ++j)
                    if (row[j] != 0.0) tri = 0;
            }
            o_U[r] = bb.add(pad.data(), pad.size() * 8);
            p->U_rows[r] = (int)rows;
            p->U_ld[r] = (int)ld;
            p->U_tri[r] = tri;
        }
    cudaError_t e = cudaMalloc(&p->blob, bb.host.size() ? bb.host.size() : 256);
    if (e != cudaSuccess) { delete p; return pq_set_cuda_error(e, "cudaMalloc(problem blob)", __LINE__); }
    p->blob_bytes = bb.host.size();
    e = cudaMemcpy(p->blob, bb.host.data(), bb.host.size(), cudaMemcpyHostToDevice);
    if (e != cudaSuccess) { cudaFree(p->blob); delete p; return pq_set_cuda_error(e, "cudaMemcpy(problem blob)", __LINE__); }
    unsigned char* base = (unsigned char*)p->blob;
    d.y = (const double*)(base + o_y);
    d.w = (const double*)(base + o_w);
    d.cval = (const double*)(base + o_c);
    d.src = (const int*)(base + o_s);
    d.mconst = (const double*)(base + o_m);
    p->mconstT = h->model_id == PEUQSE_MODEL_LINEAR ? (const double*)(base + o_mT) : nullptr;
    d.pU = h->prior_mode == PEUQSE_PRIOR_DENSE ? (const double*)(base + o_pU) : nullptr;
    for (int r = 0; r < h->n_resp; ++r)
        d.U[r] = (h->resp_mode[r] == PEUQSE_RESP_MVN_FULL) ? (const double*)(base + o_U[r]) : nullptr;

    pq_fill_mathc(d);
    // fast view
    d.fast_n = 0;
    for (int i = 0; i < PQ_MAXK; ++i) { d.gauss[i] = 0; d.pwn[i] = 0.0; }
    for (int j = 0; j < h->kp && h->prior_mode == PEUQSE_PRIOR_DIAG; ++j) { d.gauss[d.gidx[j]] = 1; d.pwn[d.gidx[j]] = d.pw[j]; }
    for (int i = 0; i < PQ_FAST_N; ++i) { d.fy[i] = 0.0; d.fw[i] = 0.0; }
    for (int i = 0; i < PQ_FAST_N * PQ_MAXK; ++i) d.fmc[i] = 0.0;
    if (h->n_resp == 1 && h->resp_mode[0] == PEUQSE_RESP_MVN_DIAG && h->n_total <= PQ_FAST_N &&
        h->prior_mode != PEUQSE_PRIOR_DENSE &&
        (h->model_id == PEUQSE_MODEL_QUADRATIC || h->model_id == PEUQSE_MODEL_LINEAR)) {
        d.fast_n = h->n_total;
        for (int j = 0; j < h->n_total; ++j) {
            d.fy[j] = h->y[j];
            d.fw[j] = h->w[j];
            if (h->model_id == PEUQSE_MODEL_QUADRATIC) d.fmc[j] = h->model_const[j];
            else for (int q = 0; q < h->k; ++q) d.fmc[j * PQ_MAXK + q] = h->model_const[(size_t)j * h->k + q];
        }
    }

    cudaGetDevice(&p->device);
    cudaDeviceProp prop;
    cudaGetDeviceProperties(&prop, p->device);
    p->sm_count = prop.multiProcessorCount;
    p->src_identity = 1;
    if (h->src)
        for (size_t i = 0; i < nt; ++i)
            if (h->src[i] != (int32_t)i) { p->src_identity = 0; break; }
    p->max_smem_optin = (int)prop.sharedMemPerBlockOptin;
    p->obs_smem_bytes = pq_obs_smem_bytes(d);
    *out = p;
    return PEUQSE_OK;
}

extern "C" int peuqse_problem_destroy(peuqse_problem* p) {
    if (!p) return PEUQSE_OK;
    if (p->blob) cudaFree(p->blob);
    delete p;
    return PEUQSE_OK;
}

// ---------------------------------------------------------------------------------------------------
// fp64 pipe microbenchmark (roofline denominator for the compute-bound kernels): every thread runs `iters`
// rounds of 8 independent chains; mode 0 = DFMA, mode 1 = DADD + DMUL pairs (what -fmad=false parity code issues)
// ---------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256) k_bench_fp64(int iters, int mode, double seed, double* __restrict__ out) {
    double a[8];
#pragma unroll
    for (int j = 0; j < 8; ++j) a[j] = seed + threadIdx.x * 1e-9 + j;
    const double m = 1.0000000001, c = 1e-12;
    if (mode == 0) {
        for (int i = 0; i < iters; ++i) {
#pragma unroll
            for (int j = 0; j < 8; ++j) a[j] = fma(a[j], m, c);
        }
    } else {
        for (int i = 0; i < iters; ++i) {
#pragma unroll
            for (int j = 0; j < 8; ++j) a[j] = __dadd_rn(__dmul_rn(a[j], m), c);
        }
    }
    double s = 0.0;
#pragma unroll
    for (int j = 0; j < 8; ++j) s += a[j];
    if (s == 123.456) out[0] = s;    // keep the chains alive
}

// issue-mix probe: 8 independent DFMA chains per thread with NI integer instructions (IMAD / LOP3 chains on their own
// registers) and NL shared-memory loads (8 bytes, conflict-free) per 8 DFMAs; result = DFMA instructions/s.  Answers whether the
// half-rate fp64 pipe leaves issue slots for the other pipes (config C3's loop issues 17 non-fp64 per 28 fp64 instructions).
template <int NI, int NL>
__global__ void __launch_bounds__(256) k_bench_fp64_mix(int iters, double seed, double* __restrict__ out) {
    __shared__ double sh[512];
    sh[threadIdx.x] = seed; sh[threadIdx.x + 256] = seed;
    __syncthreads();
    double a[8];
#pragma unroll
    for (int j = 0; j < 8; ++j) a[j] = seed + threadIdx.x * 1e-9 + j;
    const double m = 1.0000000001, c = 1e-12;
    unsigned x[4] = {threadIdx.x, threadIdx.x + 1u, threadIdx.x + 2u, threadIdx.x + 3u};
    double ld = 0.0;
    for (int i = 0; i < iters; ++i) {
#pragma unroll
        for (int j = 0; j < 8; ++j) {
            a[j] = fma(a[j], m, c);
            if (j * NI / 8 != (j + 1) * NI / 8) {
#pragma unroll
                for (int q = j * NI / 8; q < (j + 1) * NI / 8; ++q)
                    x[q & 3] = (q & 1) ? ((x[q & 3] ^ x[(q + 1) & 3]) & 0x7fffffffu) : x[q & 3] * 1664525u + 1013904223u;
            }
            if (j * NL / 8 != (j + 1) * NL / 8) {
#pragma unroll
                for (int q = j * NL / 8; q < (j + 1) * NL / 8; ++q)
                    ld += sh[(threadIdx.x + 32 * q + (x[0] & 0x100u)) & 511];
            }
        }
    }
    double s = ld + (double)(x[0] ^ x[1] ^ x[2] ^ x[3]);
#pragma unroll
    for (int j = 0; j < 8; ++j) s += a[j];
    if (s == 123.456) out[0] = s;    // keep the chains alive
}

// DMMA probe: every warp issues rounds of NACC independent mma.sync.m8n8k4.f64 accumulations from registers only (no shared
// memory, no loads): the ceiling for the dense-covariance contraction (csrc/fullcov.cu).  256 FMA = 512 flop per instruction.
template <int NACC>
__global__ void __launch_bounds__(256) k_bench_dmma(int iters, double seed, double* __restrict__ out) {
    double c[NACC][2];
#pragma unroll
    for (int j = 0; j < NACC; ++j) c[j][0] = c[j][1] = 0.0;
    const double a = seed + threadIdx.x * 1e-9, b = 1.0 - threadIdx.x * 1e-9;
    for (int i = 0; i < iters; ++i) {
#pragma unroll
        for (int j = 0; j < NACC; ++j)
            asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};"
                         : "+d"(c[j][0]), "+d"(c[j][1]) : "d"(a), "d"(b));
    }
    double s = 0.0;
#pragma unroll
    for (int j = 0; j < NACC; ++j) s += c[j][0] + c[j][1];
    if (s == 123.456) out[0] = s;
}

// DMMA probe in the register pattern of the contraction kernel (csrc/fullcov.cu): 16 accumulators per warp fed by 4 x 4
// distinct operand fragments, 8 warps per CTA; the grid decides how many CTAs share an SM
__global__ void __launch_bounds__(256) k_bench_dmma_tile(int iters, double seed, double* __restrict__ out) {
    double c[4][4][2], a[4], b[4];
#pragma unroll
    for (int i = 0; i < 4; ++i) {
        a[i] = seed + threadIdx.x * 1e-9 + i;
        b[i] = 1.0 - threadIdx.x * 1e-9 - i * 1e-3;
#pragma unroll
        for (int j = 0; j < 4; ++j) c[i][j][0] = c[i][j][1] = 0.0;
    }
    for (int it = 0; it < iters; ++it) {
#pragma unroll
        for (int i = 0; i < 4; ++i)
#pragma unroll
            for (int j = 0; j < 4; ++j)
                asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};"
                             : "+d"(c[i][j][0]), "+d"(c[i][j][1]) : "d"(a[i]), "d"(b[j]));
    }
    double s = 0.0;
#pragma unroll
    for (int i = 0; i < 4; ++i)
#pragma unroll
        for (int j = 0; j < 4; ++j) s += c[i][j][0] + c[i][j][1];
    if (s == 123.456) out[0] = s;
}

// returns through *instr_per_s the measured fp64 instruction rate (instructions, not flops) of the whole device
// latency probe: NCH dependent DFMA chains per thread, one warp per scheduler (128-thread CTA per SM)
template <int NCH>
__global__ void __launch_bounds__(128) k_bench_fp64_lat(int iters, double seed, double* __restrict__ out) {
    double a[NCH];
#pragma unroll
    for (int j = 0; j < NCH; ++j) a[j] = seed + threadIdx.x * 1e-9 + j;
    const double m = 1.0000000001, c = 1e-12;
    for (int i = 0; i < iters; ++i) {
#pragma unroll
        for (int r = 0; r < 8; ++r) {
#pragma unroll
            for (int j = 0; j < NCH; ++j) a[j] = fma(a[j], m, c);
        }
    }
    double s = 0.0;
#pragma unroll
    for (int j = 0; j < NCH; ++j) s += a[j];
    if (s == 123.456) out[0] = s;
}

// operand-pattern probe: Horner steps p = fma(r, p, c_i) with r, p and twelve c_i all in distinct registers (what an
// exp/log polynomial issues), 4 chains per thread, full occupancy -- against mode 0 whose multiplier and addend are
// two fixed registers
__global__ void __launch_bounds__(256) k_bench_fp64_horner(int iters, const double* __restrict__ coef, double* __restrict__ out) {
    double c[12], r[4], p[4];
#pragma unroll
    for (int i = 0; i < 12; ++i) c[i] = coef[i];
#pragma unroll
    for (int j = 0; j < 4; ++j) { r[j] = coef[12 + j] + threadIdx.x * 1e-12; p[j] = c[0]; }
    for (int it = 0; it < iters; ++it) {
#pragma unroll
        for (int i = 0; i < 12; ++i) {
#pragma unroll
            for (int j = 0; j < 4; ++j) p[j] = fma(r[j], p[j], c[i]);
        }
    }
    const double s = (p[0] + p[1]) + (p[2] + p[3]);
    if (s == 123.456) out[0] = s;
}

extern "C" int peuqse_bench_fp64(int32_t mode, double* instr_per_s, double* d_scratch, void* stream) {
    int dev = 0;
    PQ_CUDA_OK(cudaGetDevice(&dev));
    cudaDeviceProp prop;
    PQ_CUDA_OK(cudaGetDeviceProperties(&prop, dev));
    cudaStream_t st = (cudaStream_t)stream;
    const int blocks = prop.multiProcessorCount * 8, iters = 20000;
    cudaEvent_t e0, e1;
    PQ_CUDA_OK(cudaEventCreate(&e0));
    PQ_CUDA_OK(cudaEventCreate(&e1));
    if (mode == 2) {    // Horner operand pattern; d_scratch holds 16 doubles: 12 coefficients + 4 arguments
        const int nb = prop.multiProcessorCount * 8, it = 4000;
        k_bench_fp64_horner<<<nb, 256, 0, st>>>(50, d_scratch, d_scratch + 16);
        PQ_CUDA_OK(cudaEventRecord(e0, st));
        k_bench_fp64_horner<<<nb, 256, 0, st>>>(it, d_scratch, d_scratch + 16);
        PQ_CUDA_OK(cudaEventRecord(e1, st));
        PQ_CUDA_OK(cudaEventSynchronize(e1));
        float ms = 0.f;
        PQ_CUDA_OK(cudaEventElapsedTime(&ms, e0, e1));
        cudaEventDestroy(e0);
        cudaEventDestroy(e1);
        *instr_per_s = (double)nb * 256.0 * it * 48.0 / (ms * 1e-3);
        return PEUQSE_OK;
    }
    if (mode == 3 || mode == 4) {   // DMMA m8n8k4: 8 (mode 3) or 4 (mode 4) accumulators per warp, 8 CTAs of 8 warps per SM; result = flop/s
        const int nb = prop.multiProcessorCount * 8, it = 20000, nacc = mode == 3 ? 8 : 4;
        auto run = [&](int n_it) {
            if (mode == 3) k_bench_dmma<8><<<nb, 256, 0, st>>>(n_it, 1.0, d_scratch);
            else k_bench_dmma<4><<<nb, 256, 0, st>>>(n_it, 1.0, d_scratch);
        };
        run(100);
        PQ_CUDA_OK(cudaEventRecord(e0, st));
        run(it);
        PQ_CUDA_OK(cudaEventRecord(e1, st));
        PQ_CUDA_OK(cudaEventSynchronize(e1));
        float ms = 0.f;
        PQ_CUDA_OK(cudaEventElapsedTime(&ms, e0, e1));
        cudaEventDestroy(e0);
        cudaEventDestroy(e1);
        *instr_per_s = (double)nb * 8.0 * it * nacc * 512.0 / (ms * 1e-3);
        return PEUQSE_OK;
    }
    if (mode >= 5 && mode <= 8) {   // DMMA in the contraction kernel's register pattern: 1, 2, 4 or 8 CTAs of 8 warps per SM; result = flop/s
        const int per_sm = 1 << (mode - 5), nb = prop.multiProcessorCount * per_sm, it = 40000 / per_sm;
        k_bench_dmma_tile<<<nb, 256, 0, st>>>(100, 1.0, d_scratch);
        PQ_CUDA_OK(cudaEventRecord(e0, st));
        k_bench_dmma_tile<<<nb, 256, 0, st>>>(it, 1.0, d_scratch);
        PQ_CUDA_OK(cudaEventRecord(e1, st));
        PQ_CUDA_OK(cudaEventSynchronize(e1));
        float ms = 0.f;
        PQ_CUDA_OK(cudaEventElapsedTime(&ms, e0, e1));
        cudaEventDestroy(e0);
        cudaEventDestroy(e1);
        *instr_per_s = (double)nb * 8.0 * it * 16.0 * 512.0 / (ms * 1e-3);
        return PEUQSE_OK;
    }
    if (mode >= 20 && mode <= 25) {   // issue mix: DFMA with 4 / 8 / 16 integer instructions or 2 / 4 shared loads per 8 DFMAs
        const int nb = prop.multiProcessorCount * 8, it = 10000;
        auto run = [&](int n_it) {
            if (mode == 20) k_bench_fp64_mix<4, 0><<<nb, 256, 0, st>>>(n_it, 1.0, d_scratch);
            else if (mode == 21) k_bench_fp64_mix<8, 0><<<nb, 256, 0, st>>>(n_it, 1.0, d_scratch);
            else if (mode == 22) k_bench_fp64_mix<16, 0><<<nb, 256, 0, st>>>(n_it, 1.0, d_scratch);
            else if (mode == 23) k_bench_fp64_mix<0, 2><<<nb, 256, 0, st>>>(n_it, 1.0, d_scratch);
            else if (mode == 24) k_bench_fp64_mix<0, 4><<<nb, 256, 0, st>>>(n_it, 1.0, d_scratch);
            else k_bench_fp64_mix<5, 2><<<nb, 256, 0, st>>>(n_it, 1.0, d_scratch);
        };
        run(100);
        PQ_CUDA_OK(cudaEventRecord(e0, st));
        run(it);
        PQ_CUDA_OK(cudaEventRecord(e1, st));
        PQ_CUDA_OK(cudaEventSynchronize(e1));
        float ms = 0.f;
        PQ_CUDA_OK(cudaEventElapsedTime(&ms, e0, e1));
        cudaEventDestroy(e0);
        cudaEventDestroy(e1);
        *instr_per_s = (double)nb * 256.0 * it * 8.0 / (ms * 1e-3);
        return PEUQSE_OK;
    }
    if (mode >= 10) {   // 10 + log2(chains): dependent-chain probe, one warp per scheduler; result = DFMA/s of the device
        const int lg = mode - 10, nch = 1 << lg, it = 20000;
        const int nb = prop.multiProcessorCount;
        auto run = [&](int n_it) {
            if (lg == 0) k_bench_fp64_lat<1><<<nb, 128, 0, st>>>(n_it, 1.0, d_scratch);
            else if (lg == 1) k_bench_fp64_lat<2><<<nb, 128, 0, st>>>(n_it, 1.0, d_scratch);
            else if (lg == 2) k_bench_fp64_lat<4><<<nb, 128, 0, st>>>(n_it, 1.0, d_scratch);
            else if (lg == 3) k_bench_fp64_lat<8><<<nb, 128, 0, st>>>(n_it, 1.0, d_scratch);
            else k_bench_fp64_lat<16><<<nb, 128, 0, st>>>(n_it, 1.0, d_scratch);
        };
        run(100);
        PQ_CUDA_OK(cudaEventRecord(e0, st));
        run(it);
        PQ_CUDA_OK(cudaEventRecord(e1, st));
        PQ_CUDA_OK(cudaEventSynchronize(e1));
        float ms = 0.f;
        PQ_CUDA_OK(cudaEventElapsedTime(&ms, e0, e1));
        cudaEventDestroy(e0);
        cudaEventDestroy(e1);
        *instr_per_s = (double)nb * 128.0 * it * 8.0 * nch / (ms * 1e-3);
        return PEUQSE_OK;
    }
    k_bench_fp64<<<blocks, 256, 0, st>>>(200, mode, 1.0, d_scratch);
    PQ_CUDA_OK(cudaEventRecord(e0, st));
    k_bench_fp64<<<blocks, 256, 0, st>>>(iters, mode, 1.0, d_scratch);
    PQ_CUDA_OK(cudaEventRecord(e1, st));
    PQ_CUDA_OK(cudaEventSynchronize(e1));
    float ms = 0.f;
    PQ_CUDA_OK(cudaEventElapsedTime(&ms, e0, e1));
    cudaEventDestroy(e0);
    cudaEventDestroy(e1);
    const double n = (double)blocks * 256.0 * iters * 8.0 * (mode == 0 ? 1.0 : 2.0);
    *instr_per_s = n / (ms * 1e-3);
    return PEUQSE_OK;
}
