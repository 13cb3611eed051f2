// Peer-memory exchange for the three tiny collectives of a sharded multistart (best record, moment accumulator,
// histogram): ONE kernel per exchange pushes this rank's payload straight into every peer's mailbox over
// NVLink/NVSwitch (plain remote stores into memory mapped with CUDA IPC), raises a per-sender sequence flag, waits
// for the peers' flags and merges the gathered payloads -- instead of an NCCL collective followed by a merge kernel.
// Replaces the pickle-file rendezvous of consolidate_parallel_sampling_data (IP.py:1130-1155, 1209-1319).
//
// Mailbox of a rank (device memory it owns, zero-initialised):
//   line [2 slots][PQ_PEER_MAX_WORLD senders][PQ_PEER_PAYLOAD / 8 words] of 16 bytes {lo32, seq32, hi32, seq32}
// Low-latency protocol: every 64-bit payload word travels as ONE 16-byte store whose two 8-byte halves each carry 32 bits
// of data and the 32-bit exchange number, so the receiver polls the data itself (8-byte stores arrive whole) -- no
// separate flag, no fence between payload and flag: the latency of an exchange is one NVLink store, not a store, a
// round-trip fence and a second store.
// Exchange number q (1, 2, ... identical on all ranks: every rank issues the same exchanges in the same order) uses
// slot q & 1.  A rank can finish exchange q+1 only after every peer pushed its q+1 payload, which a peer does after it
// finished reading slot q & 1 in exchange q, so when a rank overwrites that slot in exchange q+2 nobody reads it any more:
// two slots suffice.
// The exchange number lives in DEVICE memory (a counter in the local mailbox that the kernel increments), so the exchange
// kernels take no host-side sequence argument and a whole sharded multistart can be captured in a CUDA graph and replayed.
// The wait is bounded (PEUQSE_PEER_TIMEOUT_S seconds, default 60): a peer that never arrives poisons the result (NaN / index
// -1) instead of hanging the GPU AND raises a sticky status word in mapped host memory that the host layer turns into an
// error on its next result read (peuqse_peer_status); after a time-out the mailbox protocol is void (a late peer may still
// be reading a slot) and the handle must not be used again.
#include <cstdlib>
#include <cstring>
#include "common.cuh"

#define PQ_PEER_MAX_WORLD 16
#define PQ_PEER_PAYLOAD 8192
#define PQ_PEER_DATA_BYTES (2LL * PQ_PEER_MAX_WORLD * PQ_PEER_PAYLOAD * 2)     /* 16-byte line per 8-byte word */
#define PQ_PEER_BYTES (PQ_PEER_DATA_BYTES + 64)

struct PeerView {
    int rank, world;
    unsigned char* box[PQ_PEER_MAX_WORLD];    // box[r]: rank r's mailbox as mapped into this process (box[rank] is local)
    unsigned long long* seq;                  // exchange counter of this rank (device memory, behind the flags of its mailbox)
    int* status;                              // mapped host word: 0 = fine, 1 = a wait timed out (sticky)
    unsigned long long timeout_ns;
};

struct peuqse_peer {
    PeerView v;
    void* local;
    int* h_status;
    bool opened[PQ_PEER_MAX_WORLD];
};

__device__ __forceinline__ uint4* peer_line(unsigned char* box, int slot, int sender) {
    return reinterpret_cast<uint4*>(box + ((size_t)slot * PQ_PEER_MAX_WORLD + sender) * (PQ_PEER_PAYLOAD * 2));
}
__device__ __forceinline__ void peer_store_line(uint4* p, unsigned long long w, unsigned int seq) {
    asm volatile("st.volatile.global.v4.u32 [%0], {%1, %2, %3, %4};" ::"l"(p), "r"((unsigned int)w), "r"(seq),
                 "r"((unsigned int)(w >> 32)), "r"(seq)
                 : "memory");
}
__device__ __forceinline__ uint4 peer_load_line(const uint4* p) {
    uint4 v;
    asm volatile("ld.volatile.global.v4.u32 {%0, %1, %2, %3}, [%4];" : "=r"(v.x), "=r"(v.y), "=r"(v.z), "=r"(v.w) : "l"(p) : "memory");
    return v;
}
__device__ __forceinline__ unsigned long long peer_now() {
    unsigned long long t;
    asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t));
    return t;
}

enum { PEER_OP_BEST = 0, PEER_OP_ACC = 1, PEER_OP_SUM_I64 = 2 };

// io: this rank's payload (n8 8-byte words) in, the merged result out
template <int OP>
__global__ void __launch_bounds__(256) k_peer_exchange(const __grid_constant__ PeerView V, unsigned long long* __restrict__ io,
                                                       int n8, int k) {
    extern __shared__ unsigned long long s_got[];             // [world][n8] gathered payloads
    __shared__ int s_ok;
    __shared__ unsigned long long s_seq;
    if (threadIdx.x == 0) {
        s_ok = 1;
        s_seq = *V.seq + 1ULL;                                 // every rank issues the same exchanges in the same order
        *V.seq = s_seq;
    }
    __syncthreads();
    const unsigned long long seq = s_seq;
    const unsigned int seq32 = (unsigned int)seq;
    const int slot = (int)(seq & 1ULL), me = V.rank, W = V.world;
    // push my payload into every mailbox (my own included): one 16-byte line per word, data and exchange number together
    for (int idx = threadIdx.x; idx < W * n8; idx += blockDim.x) {
        const int r = idx / n8, w = idx - r * n8;
        peer_store_line(peer_line(V.box[r], slot, me) + w, io[w], seq32);
    }
    // gather: poll every sender's lines in my mailbox until both halves carry this exchange's number
    unsigned long long* got = s_got;                          // [W][n8] when it fits, else re-read in the merge below
    const unsigned long long t0 = peer_now();
    for (int idx = threadIdx.x; idx < W * n8; idx += blockDim.x) {
        const int r = idx / n8, w = idx - r * n8;
        const uint4* ln = peer_line(V.box[me], slot, r) + w;
        uint4 v = peer_load_line(ln);
        while (v.y != seq32 || v.w != seq32) {
            if (peer_now() - t0 > V.timeout_ns) { s_ok = 0; break; }
            v = peer_load_line(ln);
        }
        got[idx] = ((unsigned long long)v.z << 32) | (unsigned long long)v.x;
    }
    __syncthreads();
    if (!s_ok) {                                               // a peer never arrived: poison the result
        for (int w = threadIdx.x; w < n8; w += blockDim.x) io[w] = 0xFFF8000000000000ULL;
        if (OP == PEER_OP_BEST && threadIdx.x == 0) io[1] = (unsigned long long)(-1LL);
        if (threadIdx.x == 0) { *(volatile int*)V.status = 1; __threadfence_system(); }
        return;
    }
    if (OP == PEER_OP_BEST) {                                  // highest logP, lowest index among equals (IP.py:933)
        if (threadIdx.x < 32) {
            Best b = best_none();
            for (int r = threadIdx.x; r < W; r += 32) {
                const unsigned long long* d = got + (size_t)r * n8;
                Best o;
                o.v = __longlong_as_double((long long)d[0]);
                o.i = (long long)d[1];
                if (o.i < 0) { o.i = 0x7fffffffffffffffLL; o.v = pq_neg_inf(); }
                b = best_merge(b, o);
            }
            b = warp_best(b);
            if (threadIdx.x == 0) {
                io[0] = (unsigned long long)__double_as_longlong(b.v);
                io[1] = (unsigned long long)((b.i == 0x7fffffffffffffffLL) ? -1LL : b.i);
            }
        }
    } else if (OP == PEER_OP_ACC) {                            // sums / minima / maxima by slot, fixed rank order
        const int mn0 = PEUQSE_ACC_MIN(k), mx0 = PEUQSE_ACC_MAX(k);
        for (int a = threadIdx.x; a < n8; a += blockDim.x) {
            double r = __longlong_as_double((long long)got[a]);
            for (int q = 1; q < W; ++q) {
                const double x = __longlong_as_double((long long)got[(size_t)q * n8 + a]);
                if (a >= mx0) r = fmax(r, x);
                else if (a >= mn0) r = fmin(r, x);
                else r += x;
            }
            io[a] = (unsigned long long)__double_as_longlong(r);
        }
    } else {                                                   // integer sums (histograms)
        for (int a = threadIdx.x; a < n8; a += blockDim.x) {
            unsigned long long r = 0ULL;
            for (int q = 0; q < W; ++q) r += got[(size_t)q * n8 + a];
            io[a] = r;
        }
    }
}

extern "C" int peuqse_peer_alloc(void** d_local, void* handle64) {
    if (!d_local || !handle64) return pq_set_error(PEUQSE_ERR_INVALID, "null argument");
    static_assert(sizeof(cudaIpcMemHandle_t) == 64, "IPC handle size");
    void* p = nullptr;
    PQ_CUDA_OK(cudaMalloc(&p, PQ_PEER_BYTES));
    PQ_CUDA_OK(cudaMemset(p, 0, PQ_PEER_BYTES));
    cudaIpcMemHandle_t h;
    cudaError_t e = cudaIpcGetMemHandle(&h, p);
    if (e != cudaSuccess) { cudaFree(p); return pq_set_cuda_error(e, "cudaIpcGetMemHandle", __LINE__); }
    memcpy(handle64, &h, 64);
    *d_local = p;
    return PEUQSE_OK;
}

extern "C" int peuqse_peer_create(void* d_local, const void* handles, int32_t rank, int32_t world, peuqse_peer** out) {
    if (!d_local || !handles || !out || world < 1 || world > PQ_PEER_MAX_WORLD || rank < 0 || rank >= world)
        return pq_set_error(PEUQSE_ERR_INVALID, "bad peer arguments");
    peuqse_peer* p = new peuqse_peer();
    p->v.rank = rank;
    p->v.world = world;
    p->v.seq = reinterpret_cast<unsigned long long*>((unsigned char*)d_local + PQ_PEER_DATA_BYTES);
    const char* ts = getenv("PEUQSE_PEER_TIMEOUT_S");
    const double tsec = ts ? atof(ts) : 60.0;
    p->v.timeout_ns = (unsigned long long)((tsec > 0 ? tsec : 60.0) * 1e9);
    p->local = d_local;
    p->h_status = nullptr;
    {
        cudaError_t e = cudaHostAlloc((void**)&p->h_status, sizeof(int), cudaHostAllocMapped);
        if (e == cudaSuccess) {
            *p->h_status = 0;
            e = cudaHostGetDevicePointer((void**)&p->v.status, p->h_status, 0);
        }
        if (e != cudaSuccess) {
            if (p->h_status) cudaFreeHost(p->h_status);
            delete p;
            return pq_set_cuda_error(e, "cudaHostAlloc(peer status)", __LINE__);
        }
    }
    for (int r = 0; r < PQ_PEER_MAX_WORLD; ++r) { p->v.box[r] = nullptr; p->opened[r] = false; }
    for (int r = 0; r < world; ++r) {
        if (r == rank) { p->v.box[r] = (unsigned char*)d_local; continue; }
        cudaIpcMemHandle_t h;
        memcpy(&h, (const char*)handles + (size_t)r * 64, 64);
        void* q = nullptr;
        cudaError_t e = cudaIpcOpenMemHandle(&q, h, cudaIpcMemLazyEnablePeerAccess);
        if (e != cudaSuccess) {
            for (int s = 0; s < r; ++s) if (p->opened[s]) cudaIpcCloseMemHandle(p->v.box[s]);
            cudaFreeHost(p->h_status);
            delete p;
            return pq_set_cuda_error(e, "cudaIpcOpenMemHandle", __LINE__);
        }
        p->v.box[r] = (unsigned char*)q;
        p->opened[r] = true;
    }
    // the gathered payloads of the largest exchange (world x PQ_PEER_PAYLOAD bytes) live in dynamic shared memory
    const int smem = PQ_PEER_MAX_WORLD * PQ_PEER_PAYLOAD;
    cudaFuncSetAttribute((const void*)k_peer_exchange<PEER_OP_BEST>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
    cudaFuncSetAttribute((const void*)k_peer_exchange<PEER_OP_ACC>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
    cudaFuncSetAttribute((const void*)k_peer_exchange<PEER_OP_SUM_I64>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
    *out = p;
    return PEUQSE_OK;
}

extern "C" int peuqse_peer_destroy(peuqse_peer* p) {
    if (!p) return PEUQSE_OK;
    for (int r = 0; r < p->v.world; ++r) if (p->opened[r]) cudaIpcCloseMemHandle(p->v.box[r]);
    if (p->local) cudaFree(p->local);
    if (p->h_status) cudaFreeHost(p->h_status);
    delete p;
    return PEUQSE_OK;
}

// 0 = every exchange so far saw all peers; 1 = a wait timed out (read after synchronising the stream of the exchanges)
extern "C" int peuqse_peer_status(const peuqse_peer* p) {
    if (!p || !p->h_status) return 0;
    return *(volatile int*)p->h_status;
}

extern "C" int64_t peuqse_peer_max_bytes(void) { return PQ_PEER_PAYLOAD; }

// d_best: this rank's record in, the merged record out (every rank ends with the same record)
extern "C" int peuqse_peer_merge_best(peuqse_peer* p, peuqse_best* d_best, void* stream) {
    if (!p || !d_best) return pq_set_error(PEUQSE_ERR_INVALID, "bad peer_merge_best arguments");
    k_peer_exchange<PEER_OP_BEST><<<1, 64, (size_t)p->v.world * 2 * 8, (cudaStream_t)stream>>>(p->v, (unsigned long long*)d_best, 2, 0);
    PQ_CUDA_OK(cudaGetLastError());
    return PEUQSE_OK;
}

// d_acc: PEUQSE_ACC_DOUBLES(k) doubles, combined over ranks like peuqse_merge_acc
extern "C" int peuqse_peer_merge_acc(peuqse_peer* p, double* d_acc, int32_t k, void* stream) {
    if (!p || !d_acc || k < 1 || k > PQ_MAXK) return pq_set_error(PEUQSE_ERR_INVALID, "bad peer_merge_acc arguments");
    k_peer_exchange<PEER_OP_ACC><<<1, 256, (size_t)p->v.world * PEUQSE_ACC_DOUBLES(k) * 8, (cudaStream_t)stream>>>(p->v, (unsigned long long*)d_acc, PEUQSE_ACC_DOUBLES(k), k);
    PQ_CUDA_OK(cudaGetLastError());
    return PEUQSE_OK;
}

// in-place sum over ranks of count int64 counters (count*8 <= peuqse_peer_max_bytes())
extern "C" int peuqse_peer_sum_i64(peuqse_peer* p, void* d_buf, int64_t count, void* stream) {
    if (!p || !d_buf || count < 0 || count * 8 > PQ_PEER_PAYLOAD) return pq_set_error(PEUQSE_ERR_INVALID, "bad peer_sum_i64 arguments");
    if (count == 0) return PEUQSE_OK;
    k_peer_exchange<PEER_OP_SUM_I64><<<1, 256, (size_t)p->v.world * count * 8, (cudaStream_t)stream>>>(p->v, (unsigned long long*)d_buf, (int)count, 0);
    PQ_CUDA_OK(cudaGetLastError());
    return PEUQSE_OK;
}
