// k5_peer.cu -- the exchange step of the row-sharded search as ONE kernel over NVLink peer memory:
// the all-gather of the per-shard top-k lists is fused into the k-way merge.
//
// No reference counterpart (the reference's retrieval is single-process CPU FAISS,
// retrieve/retrieve_faiss.py:62-74); this is SURVEY.md section 8e / north_star (4), built the way the
// hardware offers it: one process per GPU on one NVSwitch node, every rank exports a small device buffer
// through CUDA IPC, every rank maps the buffers of all peers.
//
// Export buffer of a rank (cudaMalloc'ed here, so it can be IPC-shared):
//     flags[MAXW]            u64   flags[g] = number of exchanges rank g has published INTO THIS RANK
//     2 slots x { D[max_entries] f32 ; I[max_entries] i64 }      this rank's [nq, k] result lists
// One exchange (all on the caller's stream, no host synchronisation, no NCCL):
//     1. this rank's lists -> its own slot s = step & 1                              (device copy)
//     2. publish kernel: __threadfence_system, then store `step+1` into flags[rank] of EVERY peer (remote
//        st.release.sys over NVLink) -- "my slot s is readable"
//     3. merge kernel: a CTA per query waits until the local flags of all ranks reached step+1
//        (ld.acquire.sys on LOCAL memory: no polling over the link), then loads the G lists of its query straight
//        from the peers' slots (ld.volatile.global on mapped peer pointers), sorts the G*k keys in shared memory and
//        writes the merged top-k.  (score desc, id asc) as everywhere; shards hold ascending disjoint id ranges.
// Slot reuse is safe without a barrier: a rank overwrites slot s at step i+2, after its merge i+1 has seen every
// peer's publish i+1, and a peer publishes i+1 only after its merge i (same stream) has finished reading slot s.
#include <float.h>
#include <string.h>

#include <new>

#include "common.cuh"

namespace trx {

constexpr int kMaxWorld = 16;

struct PeerView {
    const float* D[kMaxWorld];
    const int64_t* I[kMaxWorld];
};
struct PeerFlags {
    unsigned long long* f[kMaxWorld];   // flags array of every rank (index [g] is where rank g publishes)
};

__global__ void k5_publish_kernel(PeerFlags peers, int world, int rank, unsigned long long value) {
    const int p = threadIdx.x;
    if (p >= world) return;
    __threadfence_system();   // the slot copy issued earlier on this stream is visible system-wide before the flag
    asm volatile("st.release.sys.global.u64 [%0], %1;" ::"l"(peers.f[p] + rank), "l"(value) : "memory");
}

__device__ __forceinline__ unsigned long long ld_acquire_sys(const unsigned long long* p) {
    unsigned long long v;
    asm volatile("ld.acquire.sys.global.u64 %0, [%1];" : "=l"(v) : "l"(p) : "memory");
    return v;
}

__global__ void __launch_bounds__(256) k5_peer_merge_kernel(int metric, PeerView pv, const unsigned long long* my_flags,
                                                            unsigned long long want, int G, int64_t q0, int k,
                                                            float* __restrict__ D, int64_t* __restrict__ I) {
    extern __shared__ __align__(16) unsigned char smem_raw[];
    const int64_t q = q0 + blockIdx.x;      // query of the exchanged lists; output row blockIdx.x (a slice starts at q0)
    const int64_t oq = blockIdx.x;
    const int total = G * k;
    int P = 2;
    while (P < total) P <<= 1;
    // layout: keys[P] u64 | ids[total] i64 | scores[total] f32   (every peer entry crosses NVLink exactly once)
    uint64_t* keys = reinterpret_cast<uint64_t*>(smem_raw);
    int64_t* sI = reinterpret_cast<int64_t*>(keys + P);
    float* sD = reinterpret_cast<float*>(sI + total);
    // wait for every rank's publish of this exchange (flags live in LOCAL memory; peers store into them)
    if (threadIdx.x < G) {
        // a peer that never publishes (crashed rank) must not hang this GPU: trap after 30 s
        uint64_t t0 = 0;
        uint32_t polls = 0;
        while (ld_acquire_sys(my_flags + threadIdx.x) < want) {
            __nanosleep(64);
            if ((++polls & 0x3ffu) == 0) {
                uint64_t t;
                asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t));
                if (t0 == 0) t0 = t;
                else if (t - t0 > 30ull * 1000 * 1000 * 1000) __trap();
            }
        }
    }
    __syncthreads();
    for (int i = threadIdx.x; i < P; i += blockDim.x) {
        uint64_t key = KEY_SENTINEL;
        if (i < total) {
            const int g = i / k, j = i - g * k;
            const int64_t src = q * k + j;
            const int64_t id = *reinterpret_cast<const volatile int64_t*>(pv.I[g] + src);     // peer load over NVLink
            const float v = *reinterpret_cast<const volatile float*>(pv.D[g] + src);
            sI[i] = id; sD[i] = v;
            if (id >= 0) key = pack_key(metric == TRX_METRIC_L2 ? -v : v, (uint32_t)i);
        }
        keys[i] = key;
    }
    bitonic_sort_u64(keys, P);
    const float fill = metric == TRX_METRIC_L2 ? FLT_MAX : -FLT_MAX;
    for (int j = threadIdx.x; j < k; j += blockDim.x) {
        const uint64_t key = keys[j];
        if (key == KEY_SENTINEL) { D[oq * k + j] = fill; I[oq * k + j] = -1; }
        else {
            const int pos = (int)key_id(key);
            D[oq * k + j] = sD[pos];
            I[oq * k + j] = sI[pos];
        }
    }
}

// Bounds exchange of the two-phase search: every rank published payload[q] = { its nb best PREFILTER scores of query
// q, eps[q] }.  The k-th largest of the G*nb scores is a lower bound on the k-th largest prefilter score over the whole
// corpus (it is the k-th largest of a subset); the rows behind those k scores have exact scores >= score - eps_max, so
// the global k-th EXACT score is >= kth - eps_max, and a row whose prefilter score is below  kth - 2 eps_max  has an
// exact score below that: it cannot be in the global top-k.  floor[q] = kth - 2 eps_max (minus rounding slack), or
// -inf when the shards hold fewer than k candidates between them.
struct PeerPayload {
    const float* p[kMaxWorld];
};

__global__ void __launch_bounds__(256) k5_peer_floor_kernel(PeerPayload pp, const unsigned long long* my_flags,
                                                            unsigned long long want, int G, int nb, int k,
                                                            float* __restrict__ floor_out) {
    extern __shared__ __align__(16) unsigned char smem_raw[];
    uint64_t* keys = reinterpret_cast<uint64_t*>(smem_raw);
    __shared__ float s_eps[kMaxWorld];
    const int64_t q = blockIdx.x;
    const int total = G * nb;
    int P = 2;
    while (P < total) P <<= 1;
    if (threadIdx.x < G) {
        uint64_t t0 = 0;
        uint32_t polls = 0;
        while (ld_acquire_sys(my_flags + threadIdx.x) < want) {
            __nanosleep(64);
            if ((++polls & 0x3ffu) == 0) {
                uint64_t t;
                asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t));
                if (t0 == 0) t0 = t;
                else if (t - t0 > 30ull * 1000 * 1000 * 1000) __trap();
            }
        }
    }
    __syncthreads();
    for (int i = threadIdx.x; i < P; i += blockDim.x) {
        uint64_t key = KEY_SENTINEL;
        if (i < total) {
            const int g = i / nb, j = i - g * nb;
            const float v = *reinterpret_cast<const volatile float*>(pp.p[g] + q * (nb + 1) + j);   // peer load
            if (v > -INFINITY) key = pack_key(v, (uint32_t)i);
        }
        keys[i] = key;
    }
    if (threadIdx.x < G) s_eps[threadIdx.x] = *reinterpret_cast<const volatile float*>(pp.p[threadIdx.x] + q * (nb + 1) + nb);
    bitonic_sort_u64(keys, P);
    if (threadIdx.x == 0) {
        float fl = -INFINITY;
        if (k <= total && keys[k - 1] != KEY_SENTINEL) {
            float em = 0.f;
            for (int g = 0; g < G; g++) em = fmaxf(em, s_eps[g]);
            fl = key_score(keys[k - 1]) - 2.f * em * 1.00001f;
            fl -= fabsf(fl) * 1e-6f + 1e-30f;
        }
        floor_out[q] = fl;
    }
}

}  // namespace trx

using namespace trx;

struct trx_exchange {
    int device = 0, rank = 0, world = 1;
    int64_t max_entries = 0;
    unsigned char* base = nullptr;              // this rank's export buffer
    void* peer_base[kMaxWorld] = {nullptr};     // mapped export buffers (peer_base[rank] == base)
    bool connected = false;
    unsigned long long step = 0;
};

static size_t slot_bytes(int64_t max_entries) { return (size_t)max_entries * 12; }
static size_t flags_bytes() { return 256; }     // kMaxWorld u64, padded
static size_t export_bytes(int64_t max_entries) { return flags_bytes() + 2 * slot_bytes(max_entries); }
static const float* slot_D(const void* base, int64_t max_entries, int s) {
    return reinterpret_cast<const float*>((const unsigned char*)base + flags_bytes() + (size_t)s * slot_bytes(max_entries) +
                                          (size_t)max_entries * 8);
}
static const int64_t* slot_I(const void* base, int64_t max_entries, int s) {
    return reinterpret_cast<const int64_t*>((const unsigned char*)base + flags_bytes() + (size_t)s * slot_bytes(max_entries));
}

struct ExDeviceGuard {
    int prev = -1;
    explicit ExDeviceGuard(int dev) { cudaGetDevice(&prev); if (prev != dev) cudaSetDevice(dev); else prev = -1; }
    ~ExDeviceGuard() { if (prev >= 0) cudaSetDevice(prev); }
};

extern "C" {

int trx_exchange_create(int device, int rank, int world, int64_t max_entries, trx_exchange** out) {
    if (!out) { set_error("out is null"); return TRX_EINVAL; }
    *out = nullptr;
    if (world < 1 || world > kMaxWorld || rank < 0 || rank >= world || max_entries <= 0) {
        set_error("exchange: bad rank/world/size (world <= %d)", kMaxWorld);
        return TRX_EINVAL;
    }
    trx_exchange* ex = new (std::nothrow) trx_exchange();
    if (!ex) { set_error("host allocation failed"); return TRX_ENOMEM; }
    ex->device = device; ex->rank = rank; ex->world = world; ex->max_entries = max_entries;
    ExDeviceGuard g(device);
    cudaError_t e = cudaMalloc((void**)&ex->base, export_bytes(max_entries));
    if (e != cudaSuccess) {
        cudaGetLastError();
        set_error("exchange: cudaMalloc of %zu bytes failed: %s", export_bytes(max_entries), cudaGetErrorString(e));
        delete ex;
        return TRX_ENOMEM;
    }
    if (cudaMemset(ex->base, 0, export_bytes(max_entries)) != cudaSuccess) {
        set_error("exchange: memset failed"); cudaFree(ex->base); delete ex; return TRX_ECUDA;
    }
    ex->peer_base[rank] = ex->base;
    *out = ex;
    return TRX_OK;
}

int trx_exchange_handle(trx_exchange* ex, unsigned char* handle64) {
    if (!ex || !handle64) { set_error("bad argument"); return TRX_EINVAL; }
    static_assert(sizeof(cudaIpcMemHandle_t) == 64, "IPC handle size");
    ExDeviceGuard g(ex->device);
    cudaIpcMemHandle_t h;
    TRX_CUDA(cudaIpcGetMemHandle(&h, ex->base));
    memcpy(handle64, &h, 64);
    return TRX_OK;
}

int trx_exchange_connect(trx_exchange* ex, const unsigned char* handles) {
    if (!ex || !handles) { set_error("bad argument"); return TRX_EINVAL; }
    ExDeviceGuard g(ex->device);
    for (int p = 0; p < ex->world; p++) {
        if (p == ex->rank) continue;
        cudaIpcMemHandle_t h;
        memcpy(&h, handles + (size_t)p * 64, 64);
        cudaError_t e = cudaIpcOpenMemHandle(&ex->peer_base[p], h, cudaIpcMemLazyEnablePeerAccess);
        if (e != cudaSuccess) {
            cudaGetLastError();
            set_error("exchange: cannot map the export buffer of rank %d: %s", p, cudaGetErrorString(e));
            return TRX_ECUDA;
        }
    }
    ex->connected = true;
    return TRX_OK;
}

int trx_exchange_merge(trx_exchange* ex, int metric, const float* D_local, const int64_t* I_local, int64_t nq, int k,
                       float* D, int64_t* I, void* cuda_stream) {
    return trx_exchange_merge_slice(ex, metric, D_local, I_local, nq, k, 0, nq, D, I, cuda_stream);
}

int trx_exchange_merge_slice(trx_exchange* ex, int metric, const float* D_local, const int64_t* I_local, int64_t nq, int k,
                             int64_t q0, int64_t nq_out, float* D, int64_t* I, void* cuda_stream) {
    if (!ex || !D_local || !I_local || nq <= 0 || k <= 0 || q0 < 0 || nq_out < 0 || q0 + nq_out > nq ||
        (nq_out > 0 && (!D || !I))) { set_error("bad argument"); return TRX_EINVAL; }
    if (!ex->connected && ex->world > 1) { set_error("exchange: not connected"); return TRX_EINVAL; }
    if (nq * k > ex->max_entries) { set_error("exchange: %lld entries exceed the export buffer (%lld)", (long long)(nq * k), (long long)ex->max_entries); return TRX_EINVAL; }
    ExDeviceGuard g(ex->device);
    cudaStream_t st = (cudaStream_t)cuda_stream;
    const int s = (int)(ex->step & 1ull);
    const size_t n = (size_t)nq * k;
    TRX_CUDA(cudaMemcpyAsync(const_cast<float*>(slot_D(ex->base, ex->max_entries, s)), D_local, n * 4, cudaMemcpyDeviceToDevice, st));
    TRX_CUDA(cudaMemcpyAsync(const_cast<int64_t*>(slot_I(ex->base, ex->max_entries, s)), I_local, n * 8, cudaMemcpyDeviceToDevice, st));
    PeerFlags pf{};
    PeerView pv{};
    for (int p = 0; p < ex->world; p++) {
        pf.f[p] = reinterpret_cast<unsigned long long*>(ex->peer_base[p]);
        pv.D[p] = slot_D(ex->peer_base[p], ex->max_entries, s);
        pv.I[p] = slot_I(ex->peer_base[p], ex->max_entries, s);
    }
    const unsigned long long want = ex->step + 1;
    k5_publish_kernel<<<1, 32, 0, st>>>(pf, ex->world, ex->rank, want);
    count_launch();
    const int64_t total = (int64_t)ex->world * k;
    int P = 2;
    while (P < total) P <<= 1;
    const size_t smem = (size_t)P * 8 + (size_t)total * 12;
    if (smem > 200 * 1024) { set_error("exchange: world*k=%lld too large", (long long)total); return TRX_EINVAL; }
    TRX_CUDA(cudaFuncSetAttribute(k5_peer_merge_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    if (nq_out > 0) {     // (a rank with an empty slice still publishes its lists)
        k5_peer_merge_kernel<<<(unsigned)nq_out, 256, smem, st>>>(metric, pv, reinterpret_cast<const unsigned long long*>(ex->base),
                                                                  want, ex->world, q0, k, D, I);
        count_launch();
    }
    TRX_CUDA(cudaGetLastError());
    ex->step++;
    return TRX_OK;
}

int trx_exchange_floor(trx_exchange* ex, const float* payload, int64_t nq, int nb, int k, float* floor_out,
                       void* cuda_stream) {
    if (!ex || !payload || !floor_out || nq <= 0 || nb <= 0 || k <= 0) { set_error("bad argument"); return TRX_EINVAL; }
    if (!ex->connected && ex->world > 1) { set_error("exchange: not connected"); return TRX_EINVAL; }
    const size_t bytes = (size_t)nq * (nb + 1) * 4;
    if (bytes > slot_bytes(ex->max_entries)) { set_error("exchange: bounds payload of %zu bytes exceeds the export slot", bytes); return TRX_EINVAL; }
    ExDeviceGuard g(ex->device);
    cudaStream_t st = (cudaStream_t)cuda_stream;
    const int s = (int)(ex->step & 1ull);
    // the payload travels in the same double-buffered export slots (and under the same flag protocol) as the lists
    TRX_CUDA(cudaMemcpyAsync(const_cast<int64_t*>(slot_I(ex->base, ex->max_entries, s)), payload, bytes, cudaMemcpyDeviceToDevice, st));
    PeerFlags pf{};
    PeerPayload pp{};
    for (int p = 0; p < ex->world; p++) {
        pf.f[p] = reinterpret_cast<unsigned long long*>(ex->peer_base[p]);
        pp.p[p] = reinterpret_cast<const float*>(slot_I(ex->peer_base[p], ex->max_entries, s));
    }
    const unsigned long long want = ex->step + 1;
    k5_publish_kernel<<<1, 32, 0, st>>>(pf, ex->world, ex->rank, want);
    count_launch();
    int P = 2;
    while (P < ex->world * nb) P <<= 1;
    const size_t smem = (size_t)P * 8;
    if (smem > 200 * 1024) { set_error("exchange: world*nb=%d too large", ex->world * nb); return TRX_EINVAL; }
    TRX_CUDA(cudaFuncSetAttribute(k5_peer_floor_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    k5_peer_floor_kernel<<<(unsigned)nq, 256, smem, st>>>(pp, reinterpret_cast<const unsigned long long*>(ex->base), want,
                                                          ex->world, nb, k, floor_out);
    count_launch();
    TRX_CUDA(cudaGetLastError());
    ex->step++;
    return TRX_OK;
}

void trx_exchange_destroy(trx_exchange* ex) {
    if (!ex) return;
    ExDeviceGuard g(ex->device);
    cudaDeviceSynchronize();
    for (int p = 0; p < ex->world; p++)
        if (p != ex->rank && ex->peer_base[p]) cudaIpcCloseMemHandle(ex->peer_base[p]);
    if (ex->base) cudaFree(ex->base);
    delete ex;
}

}  // extern "C"
