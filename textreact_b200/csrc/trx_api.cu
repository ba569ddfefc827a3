// trx_api.cu -- the C ABI (include/trx.h) and the host-side search pipeline.
//
// Pipeline of one trx_search batch (B <= max_batch queries):
//
//   K1 query_prep   fp32 queries -> bf16 (+|q|^2), certificate slack eps
//   pass 0          K2 (tcgen05) over the 1/32 row sample -> per-query threshold thr (target: ~T rows
//                   of the corpus score above it)
//   main pass       scorer over the whole bf16 corpus, candidates with score > thr appended
//                   (K2 tcgen05 for batches, K3 CUDA-core streaming for tiny batches)
//   K4 rescore      exact fp32 score of the best candidates, certificate, final top-k
//   fallback        queries without a certificate: K3 fp32 scan + exact radix top-k
//
// The fallback is an exact GPU path, not a CPU one: nothing here ever computes on the host.
#include <float.h>
#include <limits.h>
#include <stdarg.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

#include <algorithm>
#include <atomic>
#include <mutex>
#include <new>
#include <string>
#include <vector>

#include "common.cuh"

namespace trx {

static thread_local std::string g_err;
static std::atomic<int64_t> g_launches{0};

void set_error(const char* fmt, ...) {
    char buf[1024];
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(buf, sizeof(buf), fmt, ap);
    va_end(ap);
    g_err = buf;
}
void count_launch(int n) { g_launches.fetch_add(n, std::memory_order_relaxed); }

static bool is_device_ptr(const void* p) {
    if (!p) return false;
    cudaPointerAttributes at;
    if (cudaPointerGetAttributes(&at, p) != cudaSuccess) { cudaGetLastError(); return false; }
    return at.type == cudaMemoryTypeDevice || at.type == cudaMemoryTypeManaged;
}

struct DeviceGuard {
    int prev = -1;
    explicit DeviceGuard(int dev) { cudaGetDevice(&prev); if (prev != dev) cudaSetDevice(dev); else prev = -1; }
    ~DeviceGuard() { if (prev >= 0) cudaSetDevice(prev); }
};

template <typename T>
static int dmalloc(T** p, size_t count) {
    *p = nullptr;
    if (count == 0) return TRX_OK;
    cudaError_t e = cudaMalloc((void**)p, count * sizeof(T));
    if (e != cudaSuccess) {
        cudaGetLastError();
        set_error("cudaMalloc of %zu bytes failed: %s", count * sizeof(T), cudaGetErrorString(e));
        return TRX_ENOMEM;
    }
    return TRX_OK;
}
template <typename T>
static void dfree(T*& p) { if (p) { cudaFree(p); p = nullptr; } }

__global__ void gather_rows_kernel(const float* __restrict__ src, const int32_t* __restrict__ list, int d,
                                   float* __restrict__ dst) {
    const int64_t s = list[blockIdx.x];
    for (int c = threadIdx.x; c < d; c += blockDim.x) dst[(int64_t)blockIdx.x * d + c] = src[s * d + c];
}
__global__ void add_bias_kernel(float* __restrict__ v, int n, float b) {
    int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) v[i] += b;
}
// Second prefilter pass of the queries without a certificate (finish_batch): query i of the pass is batch row list[i].
// thr2[i] comes in as the k-th best EXACT score K4 saw for it (minus the fp32 summation slack) and leaves as the
// prefilter-domain threshold  s_k (+|q|^2 for L2) - eps - eps_acc:  a row whose bf16 score is not above it cannot
// reach s_k, so the rows above it are a complete candidate list.
__global__ void second_pass_prep_kernel(const int32_t* __restrict__ list, int n, int Kp, int l2,
                                        const __nv_bfloat16* __restrict__ q16, const float* __restrict__ qnorm2,
                                        const float* __restrict__ eps, const float* __restrict__ eps_acc,
                                        __nv_bfloat16* __restrict__ q16o, float* __restrict__ epso,
                                        float* __restrict__ eps_acco, float* __restrict__ thr2,
                                        uint32_t* __restrict__ cand_cnt) {
    const int i = blockIdx.x;
    if (i >= n) return;
    const int64_t q = list[i];
    const uint4* src = reinterpret_cast<const uint4*>(q16 + q * Kp);
    uint4* dst = reinterpret_cast<uint4*>(q16o + (int64_t)i * Kp);
    for (int c = threadIdx.x; c < Kp / 8; c += blockDim.x) dst[c] = src[c];
    if (threadIdx.x == 0) {
        const float e = eps[q], ea = eps_acc[q];
        epso[i] = e; eps_acco[i] = ea;
        thr2[i] = thr2[i] + (l2 ? qnorm2[q] : 0.f) - e - ea;
        cand_cnt[i] = 0u;
    }
}

// two-phase search on the exact path: no bound from this shard (payload row = -inf scores, eps 0)
__global__ void fill_payload_kernel(float* __restrict__ payload, int64_t nq, int nb) {
    const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i < nq * (nb + 1)) payload[i] = (i % (nb + 1)) == nb ? 0.f : -INFINITY;
}

__global__ void gather_i32_kernel(const int32_t* __restrict__ src, const int32_t* __restrict__ list, int n,
                                  int32_t* __restrict__ dst) {
    int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) dst[i] = src[list[i]];
}

}  // namespace trx

using namespace trx;

// Everything one in-flight batch owns.  Two of them let batch i+1 be queued (queries uploaded on the copy
// stream, kernels behind batch i on the compute stream) before the host waits for batch i.
struct BatchWs {
    int batch = 0, cap = 0, k = 0;                         // allocated for
    float* q32 = nullptr; __nv_bfloat16* q16 = nullptr; float* qnorm2 = nullptr;
    float* eps = nullptr; float* eps_acc = nullptr; float* thr = nullptr; int32_t* excl = nullptr;
    Cand* cand = nullptr; uint32_t* cand_cnt = nullptr;
    float* slots = nullptr; size_t slots_elems = 0;
    HitRec* hitlog = nullptr; uint32_t* hitlog_cnt = nullptr; size_t hitlog_elems = 0; int hitlog_n = 0;
    float* Dd = nullptr; int64_t* Id = nullptr;
    int32_t* fb_list = nullptr; float* fb_thr = nullptr; uint32_t* fb_count = nullptr;
    uint32_t* h_nfb = nullptr;                             // pinned: fallback count of this batch
    float* hD = nullptr; int64_t* hI = nullptr; size_t h_elems = 0;   // pinned staging for pageable outputs
    cudaEvent_t q_ready = nullptr, done = nullptr, ev[4] = {nullptr, nullptr, nullptr, nullptr};
    // launch plan of the prefilter pipeline (prepare_prefilter) and its cached CUDA graph (small batches)
    bool pair = false; int S = 0, T = 0, r = 0, log_cap = 0;
    cudaGraphExec_t graph = nullptr; int graph_nodes = 0; uint64_t graph_gen = 0; int64_t graph_B = -1; int graph_k = 0, graph_path = 0;
    bool graph_excl = false; int32_t graph_attr = 0; int graph_dedup = 0;
    // the batch in flight
    int64_t B = 0; int path = 0;
    bool timed = false;                                    // ev[0..3] were recorded for the batch in flight
    bool two_phase = false; int tp_k = 0; int32_t tp_attr = INT32_MAX; int tp_dedup = 0;   // between search_begin and _finish
    const float* qdev = nullptr; const int32_t* exdev = nullptr;
    float* D = nullptr; int64_t* I = nullptr; bool out_dev = false, out_pinned = false;
};

struct trx_index {
    int d = 0, metric = 0, device = 0, Kp = 0, sm_count = 148;
    int64_t ntotal = 0, capacity = 0, id_offset = 0;
    float* x32 = nullptr;            // [capacity, d]   the FAISS-equivalent fp32 store
    __nv_bfloat16* x16 = nullptr;    // [capacity, Kp]  TMA / streaming copy
    float* xnorm2 = nullptr;         // [capacity]
    int32_t* groups = nullptr;       // [capacity] (valid when has_groups)
    bool has_groups = false;
    int32_t* attr = nullptr;         // [ntotal] per-row attribute (e.g. year), valid when has_attr
    bool has_attr = false;
    int32_t attr_below = INT32_MAX;  // rows with attr >= attr_below are ineligible (INT32_MAX: no filter)
    std::vector<int32_t> attr_sorted; // host copy, sorted: eligible fraction of a bound in O(log N)
    int dedup = 0;                   // distinct-groups mode: only the best row of a group is returned
    int group_max = 0; double group_avg = 1.0;   // largest / mean group size (computed on first use)
    float* xD = nullptr; int64_t* xI = nullptr; size_t x_elems = 0;   // exact path + dedup: widened result rows
    int32_t* x_nfound = nullptr;     // ... and the round state of scans wider than one selection
    uint32_t* norm2_max = nullptr;   // [2] float bits: max |x|^2, max |x - bf16(x)|^2 over the stored rows
    // 1/rate row sample for threshold estimation
    __nv_bfloat16* xs16 = nullptr; int64_t ns = 0, ns_cap = 0; bool sample_dirty = true;
    // options
    int opt_path = TRX_PATH_AUTO;
    int max_batch = 8192;
    int target = 768;       // expected candidates per query
    int sample_rate = 32;
    int stream_max_batch = 1;  // AUTO: batches <= this use the K3 streaming prefilter (measured crossover,
                               // profiles/r1i_summary.md: K3 wins at batch 1, K2 from batch 2 on)
    int timing = 0;
    int umma_pair = 1;      // allow the CTA-pair (cta_group::2) tiling
    int pair_min_batch = 129;  // ... for batches of at least this many queries (measured crossover)
    int pipeline = 1;       // overlap the upload / launch of batch i+1 with batch i when a call has several
    int thr_margin = 1;     // the sampled threshold is kept 2.5 eps(q) under the sample's estimate of the k-th score (0: as sampled)
    int second_pass = 1;    // queries without a certificate: batched second tcgen05 pass with a threshold that makes
                            // the candidate list complete (0: one fp32 streaming sweep per 4 queries instead)
    int graphs = 1;         // replay the prefilter pipeline of small batches (<= graph_max_batch) as one CUDA graph
    int graph_max_batch = 256;
    uint64_t gen = 1;       // bumped by everything that changes what a captured graph would do
    int64_t graph_replays = 0, graph_captures = 0;
    float thr_bias = 0.f;   // experiments only: added to every estimated threshold
    BatchWs ws[2];
    // fallback-only buffers (fallbacks run synchronously, one batch at a time)
    int fb_batch = 0;
    int32_t* fb_list2 = nullptr; float* thr2 = nullptr; float* neg_inf = nullptr;
    float* qfb = nullptr; int32_t* exfb = nullptr;
    __nv_bfloat16* q16fb = nullptr; float* epsfb = nullptr; float* eaccfb = nullptr;   // second prefilter pass
    float* xscores = nullptr; size_t xscores_elems = 0;  // exact-path score rows
    uint64_t* counters = nullptr;
    cudaStream_t own_stream = nullptr, copy_stream = nullptr;
    trx_stats_t st{};
    std::mutex mu;          // one call at a time per index: the workspaces are shared (FAISS indexes are
                            // searchable from several threads; here concurrent callers simply take turns)
};

static void free_batch_ws(BatchWs& w) {
    dfree(w.q32); dfree(w.q16); dfree(w.qnorm2); dfree(w.eps); dfree(w.eps_acc); dfree(w.thr); dfree(w.excl);
    dfree(w.cand); dfree(w.cand_cnt); dfree(w.slots); dfree(w.hitlog); dfree(w.hitlog_cnt);
    dfree(w.Dd); dfree(w.Id); dfree(w.fb_list); dfree(w.fb_thr); dfree(w.fb_count);
    if (w.h_nfb) { cudaFreeHost(w.h_nfb); w.h_nfb = nullptr; }
    if (w.hD) { cudaFreeHost(w.hD); w.hD = nullptr; }
    if (w.hI) { cudaFreeHost(w.hI); w.hI = nullptr; }
    if (w.q_ready) { cudaEventDestroy(w.q_ready); w.q_ready = nullptr; }
    if (w.done) { cudaEventDestroy(w.done); w.done = nullptr; }
    for (int i = 0; i < 4; i++) if (w.ev[i]) { cudaEventDestroy(w.ev[i]); w.ev[i] = nullptr; }
    if (w.graph) { cudaGraphExecDestroy(w.graph); w.graph = nullptr; w.graph_B = -1; }
    w.batch = w.cap = w.k = 0; w.slots_elems = w.hitlog_elems = w.h_elems = 0; w.hitlog_n = 0;
}

static void free_ws(trx_index* ix) {
    free_batch_ws(ix->ws[0]); free_batch_ws(ix->ws[1]);
    dfree(ix->fb_list2); dfree(ix->thr2); dfree(ix->neg_inf); dfree(ix->qfb); dfree(ix->exfb); dfree(ix->xscores);
    dfree(ix->q16fb); dfree(ix->epsfb); dfree(ix->eaccfb);
    dfree(ix->xD); dfree(ix->xI); dfree(ix->x_nfound); ix->x_elems = 0;
    ix->fb_batch = 0; ix->xscores_elems = 0;
}

static void free_store(trx_index* ix) {
    dfree(ix->x32); dfree(ix->x16); dfree(ix->xnorm2); dfree(ix->groups); dfree(ix->xs16); dfree(ix->attr);
    ix->capacity = 0; ix->ntotal = 0; ix->ns = ix->ns_cap = 0; ix->has_groups = false; ix->sample_dirty = true;
    ix->has_attr = false; ix->attr_sorted.clear();
}

static int grow(trx_index* ix, int64_t need) {
    if (need <= ix->capacity) return TRX_OK;
    if (need >= (int64_t)1 << 31) { set_error("ntotal %lld exceeds 2^31-1 rows per index", (long long)need); return TRX_EINVAL; }
    int64_t cap = ix->capacity == 0 ? need : std::max(need, ix->capacity + ix->capacity / 2);
    float* nx32; __nv_bfloat16* nx16; float* nn2; int32_t* ng;
    TRX_TRY(dmalloc(&nx32, (size_t)cap * ix->d));
    if (dmalloc(&nx16, (size_t)cap * ix->Kp) != TRX_OK) { cudaFree(nx32); return TRX_ENOMEM; }
    if (dmalloc(&nn2, (size_t)cap) != TRX_OK) { cudaFree(nx32); cudaFree(nx16); return TRX_ENOMEM; }
    if (dmalloc(&ng, (size_t)cap) != TRX_OK) { cudaFree(nx32); cudaFree(nx16); cudaFree(nn2); return TRX_ENOMEM; }
    if (ix->ntotal > 0) {
        TRX_CUDA(cudaMemcpy(nx32, ix->x32, (size_t)ix->ntotal * ix->d * 4, cudaMemcpyDeviceToDevice));
        TRX_CUDA(cudaMemcpy(nx16, ix->x16, (size_t)ix->ntotal * ix->Kp * 2, cudaMemcpyDeviceToDevice));
        TRX_CUDA(cudaMemcpy(nn2, ix->xnorm2, (size_t)ix->ntotal * 4, cudaMemcpyDeviceToDevice));
        if (ix->has_groups) TRX_CUDA(cudaMemcpy(ng, ix->groups, (size_t)ix->ntotal * 4, cudaMemcpyDeviceToDevice));
    }
    dfree(ix->x32); dfree(ix->x16); dfree(ix->xnorm2); dfree(ix->groups);
    ix->x32 = nx32; ix->x16 = nx16; ix->xnorm2 = nn2; ix->groups = ng; ix->capacity = cap;
    return TRX_OK;
}

// CTA-pair tiling (256-query tiles) or single-CTA tiling (128-query tiles) for a batch of B queries.  Pairs win from
// 129 queries on -- except where the last pair tile would be half empty and there are few tiles: 257..384 queries run
// as three 128-row tiles (measured at 4M x 768: 1.95 ms vs 2.08 ms as two pair tiles; profiles/round2_p_sweep.json).
static bool use_pair_tiling(const trx_index* ix, int64_t B) {
    if (!ix->umma_pair || B < ix->pair_min_batch) return false;
    return !(B > 256 && B <= 384);
}

// Anything that changes the stored rows or their side data abandons a pending two-phase search: its candidate lists
// refer to what was there at trx_search_begin (trx_search_finish then fails cleanly instead of reading freed rows).
static void abandon_two_phase(trx_index* ix) { ix->ws[0].two_phase = false; ix->ws[1].two_phase = false; }

static bool attr_active(const trx_index* ix) { return ix->has_attr && ix->attr_below != INT32_MAX; }

// fraction of rows that pass the attribute filter
static double eligible_fraction(const trx_index* ix) {
    if (!attr_active(ix) || ix->attr_sorted.empty()) return 1.0;
    auto it = std::lower_bound(ix->attr_sorted.begin(), ix->attr_sorted.end(), ix->attr_below);
    return (double)(it - ix->attr_sorted.begin()) / (double)ix->attr_sorted.size();
}

// Candidates per query the prefilter aims for.  The threshold is sized on ALL rows; with an attribute filter
// only a fraction f of the candidates is eligible, so the target grows by 1/f (bounded by K4's list size).
static bool dedup_active(const trx_index* ix) { return ix->dedup && ix->has_groups; }

static int effective_target(const trx_index* ix, int k) {
    int T = std::max(ix->target, 4 * k);
    double scale = 1.0 / std::max(eligible_fraction(ix), 1e-6);
    if (dedup_active(ix)) scale *= std::max(1.0, ix->group_avg);   // k distinct groups need ~k * (group size) rows
    if (scale > 1.0) T = (int)std::min<double>(2048.0, std::ceil(T * scale));
    return T;
}

// largest and mean group size: sizes the candidate target and the widened exact scan of the distinct-groups mode
static int ensure_group_stats(trx_index* ix) {
    if (!dedup_active(ix) || ix->group_max > 0) return TRX_OK;
    std::vector<int32_t> g;
    try { g.resize((size_t)ix->ntotal); } catch (...) { set_error("host allocation failed"); return TRX_ENOMEM; }
    TRX_CUDA(cudaMemcpy(g.data(), ix->groups, (size_t)ix->ntotal * 4, cudaMemcpyDeviceToHost));
    std::sort(g.begin(), g.end());
    int64_t ngroups = 0; int best = 0, run = 0;
    for (size_t i = 0; i < g.size(); i++) {
        if (i == 0 || g[i] != g[i - 1]) { ngroups++; run = 0; }
        if (++run > best) best = run;
    }
    ix->group_max = std::max(best, 1);
    ix->group_avg = ngroups > 0 ? (double)g.size() / (double)ngroups : 1.0;
    return TRX_OK;
}

// largest k the prefilter paths serve: the candidate target is 4k and K4 keeps 4x the target in shared memory
constexpr int kMaxPrefilterK = 512;

static int candidate_cap(const trx_index* ix, int k) {
    int T = effective_target(ix, k);
    int cap = 4 * T;
    int p = 1024;
    while (p < cap) p <<= 1;
    return p;
}

static int ensure_ws(trx_index* ix, BatchWs& w, int B, int k, int cap) {
    if (!w.done) {
        TRX_CUDA(cudaEventCreateWithFlags(&w.q_ready, cudaEventDisableTiming));
        TRX_CUDA(cudaEventCreateWithFlags(&w.done, cudaEventDisableTiming));
        for (int i = 0; i < 4; i++) TRX_CUDA(cudaEventCreate(&w.ev[i]));
        TRX_CUDA(cudaHostAlloc((void**)&w.h_nfb, 64, cudaHostAllocDefault));
        TRX_TRY(dmalloc(&w.fb_count, 1));
    }
    if (B > w.batch || cap > w.cap) {
        int nb = std::max(B, w.batch), nc = std::max(cap, w.cap);
        dfree(w.q32); dfree(w.q16); dfree(w.qnorm2); dfree(w.eps); dfree(w.eps_acc); dfree(w.thr); dfree(w.excl);
        dfree(w.cand); dfree(w.cand_cnt); dfree(w.fb_list); dfree(w.fb_thr);
        dfree(w.Dd); dfree(w.Id); w.k = 0;
        TRX_TRY(dmalloc(&w.q32, (size_t)nb * ix->d));
        TRX_TRY(dmalloc(&w.q16, (size_t)(nb + 8) * ix->Kp));   // K2 moves round8(nq) query rows
        TRX_CUDA(cudaMemset(w.q16, 0, (size_t)(nb + 8) * ix->Kp * 2));
        TRX_TRY(dmalloc(&w.qnorm2, (size_t)nb));
        TRX_TRY(dmalloc(&w.eps, (size_t)nb));
        TRX_TRY(dmalloc(&w.eps_acc, (size_t)nb));
        TRX_TRY(dmalloc(&w.thr, (size_t)nb));
        TRX_TRY(dmalloc(&w.excl, (size_t)nb));
        TRX_TRY(dmalloc(&w.cand, (size_t)nb * nc));
        TRX_TRY(dmalloc(&w.cand_cnt, (size_t)nb));
        TRX_TRY(dmalloc(&w.fb_list, (size_t)nb));
        TRX_TRY(dmalloc(&w.fb_thr, (size_t)nb));
        w.batch = nb; w.cap = nc; ix->gen++;
    }
    if (k > w.k) {
        ix->gen++;
        dfree(w.Dd); dfree(w.Id);
        TRX_TRY(dmalloc(&w.Dd, (size_t)w.batch * k));
        TRX_TRY(dmalloc(&w.Id, (size_t)w.batch * k));
        w.k = k;
    }
    return TRX_OK;
}

static int ensure_fallback_ws(trx_index* ix, int B) {
    if (B <= ix->fb_batch) return TRX_OK;
    dfree(ix->fb_list2); dfree(ix->thr2); dfree(ix->neg_inf); dfree(ix->qfb); dfree(ix->exfb);
    dfree(ix->q16fb); dfree(ix->epsfb); dfree(ix->eaccfb);
    TRX_TRY(dmalloc(&ix->q16fb, (size_t)(B + 8) * ix->Kp));
    TRX_CUDA(cudaMemset(ix->q16fb, 0, (size_t)(B + 8) * ix->Kp * 2));
    TRX_TRY(dmalloc(&ix->epsfb, (size_t)B));
    TRX_TRY(dmalloc(&ix->eaccfb, (size_t)B));
    TRX_TRY(dmalloc(&ix->fb_list2, (size_t)B));
    TRX_TRY(dmalloc(&ix->thr2, (size_t)B));
    TRX_TRY(dmalloc(&ix->neg_inf, (size_t)B));
    TRX_TRY(dmalloc(&ix->qfb, (size_t)B * ix->d));
    TRX_TRY(dmalloc(&ix->exfb, (size_t)B));
    std::vector<float> ninf((size_t)B, -INFINITY);
    TRX_CUDA(cudaMemcpy(ix->neg_inf, ninf.data(), (size_t)B * 4, cudaMemcpyHostToDevice));
    ix->fb_batch = B;
    return TRX_OK;
}

static int ensure_sample(trx_index* ix, cudaStream_t st) {
    if (!ix->sample_dirty) return TRX_OK;
    int64_t ns = (ix->ntotal + ix->sample_rate - 1) / ix->sample_rate;
    if (ns > ix->ns_cap) {
        dfree(ix->xs16);
        TRX_TRY(dmalloc(&ix->xs16, (size_t)ns * ix->Kp));
        ix->ns_cap = ns;
    }
    ix->ns = ns;
    TRX_TRY(launch_sample_gather(ix->x16, ix->ntotal, ix->Kp, ix->sample_rate, ix->xs16, ns, st));
    ix->sample_dirty = false;
    ix->gen++;
    return TRX_OK;
}

// exact path for nq queries whose fp32 rows are qdev[nq][d]; results to rows qmap[i] (or i) of Dd/Id.
static int run_exact(trx_index* ix, const float* qdev, const int32_t* excl_dev, const int32_t* qmap_dev, int64_t nq,
                     int k, float* Dd, int64_t* Id, cudaStream_t st) {
    const int64_t N = ix->ntotal;
    size_t budget = (size_t)64 << 20;  // floats (256 MB)
    int64_t rows = (int64_t)std::max<size_t>(8, std::min<size_t>(1024, budget / (size_t)N));
    rows = std::min<int64_t>(rows, std::max<int64_t>(nq, 1));
    if ((size_t)rows * N > ix->xscores_elems) {
        dfree(ix->xscores);
        TRX_TRY(dmalloc(&ix->xscores, (size_t)rows * N));
        ix->xscores_elems = (size_t)rows * N;
    }
    // distinct-groups mode: the top k * (largest group) rows always contain k group leaders.  When that exceeds the
    // widest exact selection (2048 rows) the scan runs in rounds: leaders of the best 2048 rows, every row of the groups
    // seen so far masked out of the score rows, next 2048 ... until k leaders are found or no row is left.
    const bool dd = dedup_active(ix);
    int kx = k;
    bool rounds = false;
    if (dd) {
        const int64_t want = (int64_t)k * ix->group_max;
        rounds = want > 2048;
        kx = (int)std::min<int64_t>(want, 2048);
        if ((size_t)rows * kx > ix->x_elems) {
            dfree(ix->xD); dfree(ix->xI);
            TRX_TRY(dmalloc(&ix->xD, (size_t)rows * kx));
            TRX_TRY(dmalloc(&ix->xI, (size_t)rows * kx));
            ix->x_elems = (size_t)rows * kx;
        }
        if (!ix->x_nfound) TRX_TRY(dmalloc(&ix->x_nfound, 1025));   // [<= 1024] leaders found so far, [1024]: unfinished count
    }
    for (int64_t q0 = 0; q0 < nq; q0 += rows) {
        int64_t nb = std::min(rows, nq - q0);
        StreamArgs a{};
        a.x = ix->x32; a.pitch = ix->d; a.n = N; a.d = ix->d;
        a.q32 = qdev + q0 * ix->d; a.q_pitch = ix->d; a.nq = nb;
        a.groups = ix->has_groups ? ix->groups : nullptr;
        a.excl = (excl_dev && ix->has_groups) ? excl_dev + q0 : nullptr;
        a.attr = attr_active(ix) ? ix->attr : nullptr; a.attr_below = ix->attr_below;
        a.out = ix->xscores; a.out_ld = N;
        a.metric = ix->metric; a.bf16 = false; a.append = false;
        TRX_TRY(launch_stream(a, ix->sm_count, st));
        const bool l2 = ix->metric == TRX_METRIC_L2;
        if (!dd) {
            TRX_TRY(launch_exact_topk(ix->xscores, N, N, nb, k, l2, ix->id_offset, qmap_dev ? qmap_dev + q0 : nullptr,
                                      qmap_dev ? Dd : Dd + q0 * k, qmap_dev ? Id : Id + q0 * k, st));
        } else {
            int32_t* nfound = rounds ? ix->x_nfound : nullptr;
            uint32_t* unfinished = rounds ? reinterpret_cast<uint32_t*>(ix->x_nfound + 1024) : nullptr;
            if (rounds) TRX_CUDA(cudaMemsetAsync(ix->x_nfound, 0, 1025 * 4, st));
            for (;;) {
                TRX_TRY(launch_exact_topk(ix->xscores, N, N, nb, kx, l2, ix->id_offset, nullptr, ix->xD, ix->xI, st));
                TRX_TRY(launch_dedup_rows(ix->xD, ix->xI, kx, ix->groups, ix->id_offset, k, l2,
                                          qmap_dev ? qmap_dev + q0 : nullptr, nb, qmap_dev ? Dd : Dd + q0 * k,
                                          qmap_dev ? Id : Id + q0 * k, nfound, unfinished, st));
                if (!rounds) break;
                uint32_t left = 0;    // (the rounds path is host-synchronous: rare, and only ever on the exact path)
                TRX_CUDA(cudaMemcpyAsync(&left, unfinished, 4, cudaMemcpyDeviceToHost, st));
                TRX_CUDA(cudaStreamSynchronize(st));
                if (left == 0) break;
                TRX_CUDA(cudaMemsetAsync(unfinished, 0, 4, st));
                TRX_TRY(launch_mask_seen_groups(ix->xscores, N, N, ix->xI, kx, ix->groups, ix->id_offset, nfound, k, nb, st));
            }
        }
    }
    ix->st.queries_exact += nq;
    return TRX_OK;
}

static bool is_pinned_host_ptr(const void* p) {
    cudaPointerAttributes at;
    if (cudaPointerGetAttributes(&at, p) != cudaSuccess) { cudaGetLastError(); return false; }
    return at.type == cudaMemoryTypeHost;
}

static RescoreArgs rescore_args(const trx_index* ix, const BatchWs& w, int k) {
    RescoreArgs ra{};
    ra.cand = w.cand; ra.cand_cnt = w.cand_cnt; ra.cap = w.cap; ra.thr = w.thr; ra.eps = w.eps;
    ra.x32 = ix->x32; ra.d = ix->d; ra.n = ix->ntotal; ra.q32 = w.qdev; ra.nq = w.B;
    ra.groups = ix->has_groups ? ix->groups : nullptr; ra.excl = w.exdev;
    ra.attr = attr_active(ix) ? ix->attr : nullptr; ra.attr_below = ix->attr_below;
    ra.dedup = dedup_active(ix) ? 1 : 0;
    ra.k = k; ra.metric = ix->metric; ra.id_offset = ix->id_offset;
    ra.D = w.Dd; ra.I = w.Id; ra.fb_list = w.fb_list; ra.fb_count = w.fb_count;
    ra.fb_thr = w.fb_thr; ra.eps_acc = w.eps_acc; ra.qmap = nullptr; ra.counters = ix->counters;
    return ra;
}

// results of batch w -> the caller's buffers (device: D2D; pinned host: D2H in place; pageable host: D2H into
// pinned staging, copied out by finish_batch)
static int send_results(trx_index* ix, BatchWs& w, int k, cudaStream_t st) {
    const size_t n = (size_t)w.B * k;
    if (w.out_dev || w.out_pinned) {
        const cudaMemcpyKind kind = w.out_dev ? cudaMemcpyDeviceToDevice : cudaMemcpyDeviceToHost;
        TRX_CUDA(cudaMemcpyAsync(w.D, w.Dd, n * 4, kind, st));
        TRX_CUDA(cudaMemcpyAsync(w.I, w.Id, n * 8, kind, st));
    } else {
        if (n > w.h_elems) {
            if (w.hD) cudaFreeHost(w.hD);
            if (w.hI) cudaFreeHost(w.hI);
            w.hD = nullptr; w.hI = nullptr; w.h_elems = 0;
            const size_t want = std::max(n, (size_t)w.batch * k);
            TRX_CUDA(cudaHostAlloc((void**)&w.hD, want * 4, cudaHostAllocDefault));
            TRX_CUDA(cudaHostAlloc((void**)&w.hI, want * 8, cudaHostAllocDefault));
            w.h_elems = want;
        }
        TRX_CUDA(cudaMemcpyAsync(w.hD, w.Dd, n * 4, cudaMemcpyDeviceToHost, st));
        TRX_CUDA(cudaMemcpyAsync(w.hI, w.Id, n * 8, cudaMemcpyDeviceToHost, st));
    }
    return TRX_OK;
}

static int ensure_hitlog(trx_index* ix, BatchWs& w, int nlogs, int log_cap) {
    const size_t need = (size_t)nlogs * log_cap;
    if (need > w.hitlog_elems || nlogs > w.hitlog_n) {
        const size_t elems = std::max(need, w.hitlog_elems);
        const int n = std::max(nlogs, w.hitlog_n);
        dfree(w.hitlog); dfree(w.hitlog_cnt);
        TRX_TRY(dmalloc(&w.hitlog, elems));
        TRX_TRY(dmalloc(&w.hitlog_cnt, (size_t)n));
        w.hitlog_elems = elems; w.hitlog_n = n; ix->gen++;
    }
    return TRX_OK;
}

// Sizes every buffer the prefilter pipeline of batch w needs and fixes its launch plan (no stream work here, so
// that enqueue_prefilter can run inside a stream capture).
static int prepare_prefilter(trx_index* ix, BatchWs& w, int k) {
    const int64_t B = w.B, N = ix->ntotal;
    w.T = effective_target(ix, k);
    w.r = std::max(1, (w.T + ix->sample_rate / 2) / ix->sample_rate);
    w.pair = use_pair_tiling(ix, B);
    w.S = umma_num_slices(ix->ns, B, ix->sm_count, w.pair);
    size_t need = (size_t)B * w.S * 32;
    if (need > w.slots_elems) { dfree(w.slots); TRX_TRY(dmalloc(&w.slots, need)); w.slots_elems = need; ix->gen++; }
    if (w.path == TRX_PATH_UMMA) {   // private hit logs: 3x the expected hits per epilogue thread, at least 256 entries
        const int grid = umma_grid(B, N, ix->sm_count, w.pair, false);
        const int nlogs = grid * 128;
        double expect = 1.15 * (double)B * (double)w.T / (double)nlogs;
        w.log_cap = std::max(256, (int)(3.0 * expect) + 64);
        TRX_TRY(ensure_hitlog(ix, w, nlogs, w.log_cap));
    }
    return TRX_OK;
}

// The prefilter pipeline of one batch, stream work only:
//   K1 batch begin -> K2<SLOTMAX> on the sample -> thresholds -> main pass (K2<THRESH> + scatter | K3) -> K4 -> count
static int enqueue_prefilter(trx_index* ix, BatchWs& w, int k, cudaStream_t st, bool with_rescore = true) {
    const int64_t B = w.B, N = ix->ntotal;
    TRX_TRY(launch_query_prep(w.qdev, B, ix->d, ix->Kp, ix->metric, w.q16, w.qnorm2, ix->norm2_max, w.eps,
                              w.eps_acc, w.cand_cnt, w.fb_count, st));
    // pass 0 (both prefilter paths): tcgen05 scores of the batch against the 1/32 row sample, slot maxima,
    // r-th largest -> per-query threshold that ~T corpus rows are expected to beat
    UmmaArgs u{};
    u.q16 = w.q16; u.nq = B; u.x16 = ix->xs16; u.n = ix->ns; u.Kp = ix->Kp;
    u.pair = w.pair;
    u.mode = 2; u.out = w.slots;
    TRX_TRY(launch_umma(u, ix->sm_count, st));
    TRX_TRY(launch_slot_thr(w.slots, B, w.S, std::min(w.r, 32 * w.S), w.thr, st, ix->thr_margin ? w.eps : nullptr,
                            (k + ix->sample_rate - 1) / ix->sample_rate));
    if (ix->thr_bias != 0.f) {
        add_bias_kernel<<<(unsigned)((B + 255) / 256), 256, 0, st>>>(w.thr, (int)B, ix->thr_bias);
        count_launch();
    }
    if (w.timed) TRX_CUDA(cudaEventRecord(w.ev[1], st));

    if (w.path == TRX_PATH_UMMA) {
        u.x16 = ix->x16; u.n = N; u.mode = 1; u.out = nullptr;
        u.thr = w.thr; u.cand = w.cand; u.cand_cnt = w.cand_cnt; u.cap = w.cap;
        u.log = w.hitlog; u.log_cnt = w.hitlog_cnt; u.log_cap = w.log_cap;
        TRX_TRY(launch_umma(u, ix->sm_count, st));
    } else {
        // main pass on the CUDA cores: one sweep over the bf16 corpus per 4 queries, hits appended directly
        StreamArgs a{};
        a.x = ix->x16; a.pitch = ix->Kp; a.n = N; a.d = ix->Kp;
        a.q16 = w.q16; a.q_pitch = ix->Kp; a.nq = B;
        a.metric = TRX_METRIC_INNER_PRODUCT; a.bf16 = true; a.append = true;
        a.thr = w.thr; a.cand = w.cand; a.cand_cnt = w.cand_cnt; a.cap = w.cap;
        TRX_TRY(launch_stream(a, ix->sm_count, st));
    }
    if (w.timed) TRX_CUDA(cudaEventRecord(w.ev[2], st));
    if (!with_rescore) return TRX_OK;       // two-phase search: K4 runs in trx_search_finish, with the exchanged floor

    TRX_TRY(launch_rescore(rescore_args(ix, w, k), st));
    TRX_CUDA(cudaMemcpyAsync(w.h_nfb, w.fb_count, 4, cudaMemcpyDeviceToHost, st));
    return TRX_OK;
}

// Queue one batch: query upload (copy stream), every kernel, the fallback count and the results (compute
// stream).  Nothing here waits for the GPU.
static int launch_batch(trx_index* ix, BatchWs& w, const float* xq, bool xq_dev, int64_t B, int k, const int32_t* excl,
                        bool excl_dev, float* D, int64_t* I, bool out_dev, cudaStream_t st, int bounds_nb = 0,
                        float* payload = nullptr) {
    const bool first_half = payload != nullptr;     // trx_search_begin: stop after the candidate lists + bounds
    const int64_t N = ix->ntotal;
    const int cap = candidate_cap(ix, k);
    TRX_TRY(ensure_ws(ix, w, (int)B, k, cap));
    w.B = B; w.D = D; w.I = I; w.out_dev = out_dev;
    w.out_pinned = !first_half && !out_dev && is_pinned_host_ptr(D) && is_pinned_host_ptr(I);
    w.two_phase = false;

    // queries / exclusion list on the device: uploaded on the copy stream so that a (host-blocking) copy from
    // pageable memory runs while the previous batch computes
    w.qdev = xq;
    w.exdev = nullptr;
    bool uploaded = false;
    if (!xq_dev) {
        TRX_CUDA(cudaMemcpyAsync(w.q32, xq, (size_t)B * ix->d * 4, cudaMemcpyHostToDevice, ix->copy_stream));
        w.qdev = w.q32; uploaded = true;
    }
    if (excl && ix->has_groups) {
        if (excl_dev) w.exdev = excl;
        else {
            TRX_CUDA(cudaMemcpyAsync(w.excl, excl, (size_t)B * 4, cudaMemcpyHostToDevice, ix->copy_stream));
            w.exdev = w.excl; uploaded = true;
        }
    }
    if (uploaded) {
        TRX_CUDA(cudaEventRecord(w.q_ready, ix->copy_stream));
        TRX_CUDA(cudaStreamWaitEvent(st, w.q_ready, 0));
    }

    int path = ix->opt_path;
    if (path == TRX_PATH_AUTO) {
        if (N <= std::max<int64_t>(8192, 2 * (int64_t)cap) || k > kMaxPrefilterK) path = TRX_PATH_EXACT;
        else if (B <= ix->stream_max_batch) path = TRX_PATH_STREAM;
        else path = TRX_PATH_UMMA;
    } else if (path != TRX_PATH_EXACT && (N <= 2 * (int64_t)cap || k > kMaxPrefilterK)) {
        path = TRX_PATH_EXACT;  // prefilter needs a corpus larger than the candidate list
    }
    // a filter that keeps less than ~1/8 of the rows would starve the candidate lists: scan exactly instead
    if (path != TRX_PATH_EXACT && eligible_fraction(ix) * 2048.0 < (double)std::max(ix->target, 4 * k) * 0.33)
        path = TRX_PATH_EXACT;
    // very wide rows (d near 16384 with a large candidate list): K4's shared-memory budget would not hold
    if (path != TRX_PATH_EXACT && k4_smem_bytes(cap, ix->d, k, dedup_active(ix)) > kK4MaxSmem) path = TRX_PATH_EXACT;
    w.path = path;
    ix->st.last_path = path;
    // Small batches replay a captured graph (below); every other batch carries four event records, so that the
    // per-stage device times of the batches a caller times are available afterwards (trx_stats sums).
    const bool graph_ok = path != TRX_PATH_EXACT && ix->graphs && B <= ix->graph_max_batch && !ix->timing && !first_half &&
                          st != nullptr && st != cudaStreamLegacy && st != cudaStreamPerThread;
    w.timed = !graph_ok;
    if (w.timed) TRX_CUDA(cudaEventRecord(w.ev[0], st));

    if (first_half) {
        // the caller's query / mask buffers need not outlive this call: the second half reads the workspace copies
        if (w.qdev != w.q32) {
            TRX_CUDA(cudaMemcpyAsync(w.q32, w.qdev, (size_t)B * ix->d * 4, cudaMemcpyDeviceToDevice, st));
            w.qdev = w.q32;
        }
        if (w.exdev != nullptr && w.exdev != w.excl) {
            TRX_CUDA(cudaMemcpyAsync(w.excl, w.exdev, (size_t)B * 4, cudaMemcpyDeviceToDevice, st));
            w.exdev = w.excl;
        }
        if (path == TRX_PATH_EXACT) {
            const int64_t n = B * (bounds_nb + 1);
            fill_payload_kernel<<<(unsigned)((n + 255) / 256), 256, 0, st>>>(payload, B, bounds_nb);
            count_launch();
            TRX_CUDA(cudaGetLastError());
            if (w.timed) { TRX_CUDA(cudaEventRecord(w.ev[1], st)); TRX_CUDA(cudaEventRecord(w.ev[2], st)); }
        } else {
            TRX_TRY(ensure_sample(ix, st));
            TRX_TRY(prepare_prefilter(ix, w, k));
            TRX_TRY(enqueue_prefilter(ix, w, k, st, false));
            BoundsArgs ba{};
            ba.cand = w.cand; ba.cand_cnt = w.cand_cnt; ba.cap = w.cap; ba.nq = B;
            ba.groups = ix->has_groups ? ix->groups : nullptr; ba.excl = w.exdev;
            ba.attr = attr_active(ix) ? ix->attr : nullptr; ba.attr_below = ix->attr_below;
            ba.eps = w.eps; ba.nb = bounds_nb; ba.payload = payload;
            TRX_TRY(launch_bounds(ba, st));
        }
        w.two_phase = true; w.tp_k = k; w.tp_attr = ix->attr_below; w.tp_dedup = ix->dedup;
        return TRX_OK;
    }

    if (path == TRX_PATH_EXACT) {
        TRX_TRY(run_exact(ix, w.qdev, w.exdev, nullptr, B, k, w.Dd, w.Id, st));
        *w.h_nfb = 0;
    } else {
        TRX_TRY(ensure_sample(ix, st));
        TRX_TRY(prepare_prefilter(ix, w, k));
        bool replayed = false;
        // Small batches are launch-latency sensitive: the pipeline of a given (batch size, k, path) is captured once
        // and replayed as one graph.  Not on the legacy / NULL stream (not capturable) and not while timing.
        if (graph_ok) {
            // the graph reads its inputs from the workspace: bring device-resident inputs there first
            if (w.qdev != w.q32) {
                TRX_CUDA(cudaMemcpyAsync(w.q32, w.qdev, (size_t)B * ix->d * 4, cudaMemcpyDeviceToDevice, st));
                w.qdev = w.q32;
            }
            if (w.exdev != nullptr && w.exdev != w.excl) {
                TRX_CUDA(cudaMemcpyAsync(w.excl, w.exdev, (size_t)B * 4, cudaMemcpyDeviceToDevice, st));
                w.exdev = w.excl;
            }
            const bool hit = w.graph != nullptr && w.graph_gen == ix->gen && w.graph_B == B && w.graph_k == k &&
                             w.graph_path == path && w.graph_excl == (w.exdev != nullptr) &&
                             w.graph_attr == ix->attr_below && w.graph_dedup == ix->dedup;
            if (!hit) {
                if (w.graph) { cudaGraphExecDestroy(w.graph); w.graph = nullptr; }
                cudaGraph_t g = nullptr;
                if (cudaStreamBeginCapture(st, cudaStreamCaptureModeThreadLocal) == cudaSuccess) {
                    const int64_t l0 = g_launches.load();
                    int rc = enqueue_prefilter(ix, w, k, st);
                    cudaError_t e = cudaStreamEndCapture(st, &g);
                    w.graph_nodes = (int)(g_launches.load() - l0);
                    if (rc == TRX_OK && e == cudaSuccess && g != nullptr &&
                        cudaGraphInstantiate(&w.graph, g, 0) == cudaSuccess) {
                        ix->graph_captures++;
                        w.graph_gen = ix->gen; w.graph_B = B; w.graph_k = k; w.graph_path = path;
                        w.graph_excl = w.exdev != nullptr;
                        w.graph_attr = ix->attr_below; w.graph_dedup = ix->dedup;
                        count_launch(-w.graph_nodes);      // captured, not launched
                    } else {
                        w.graph = nullptr;
                        ix->graphs = 0;                    // capture is not possible here: plain launches from now on
                        count_launch(-w.graph_nodes);
                    }
                    if (g) cudaGraphDestroy(g);
                    cudaGetLastError();
                } else {
                    cudaGetLastError();
                    ix->graphs = 0;
                }
            }
            if (w.graph != nullptr) {
                TRX_CUDA(cudaGraphLaunch(w.graph, st));
                count_launch(w.graph_nodes);
                ix->graph_replays++;
                replayed = true;
            }
        }
        if (!replayed) TRX_TRY(enqueue_prefilter(ix, w, k, st));
    }
    // Common case (every query certified): the results leave with the same synchronisation that reads the
    // fallback count.  Otherwise finish_batch fills the missing rows and sends them again.
    if (w.timed) TRX_CUDA(cudaEventRecord(w.ev[3], st));
    TRX_TRY(send_results(ix, w, k, st));
    TRX_CUDA(cudaEventRecord(w.done, st));
    return TRX_OK;
}

// Wait for batch w; run the exact fallbacks of the queries without a certificate; hand the results over.
static int finish_batch(trx_index* ix, BatchWs& w, int k, cudaStream_t st) {
    const int64_t N = ix->ntotal;
    const int64_t B = w.B;
    TRX_CUDA(cudaEventSynchronize(w.done));
    const uint32_t nfb = w.path == TRX_PATH_EXACT ? 0u : *w.h_nfb;
    if (nfb > 0) {
        // Queries without a certificate.  Second chance, still exact: K4 left the k-th best exact score s_k it saw;
        // every true top-k row scores at least that, hence has a bf16 score above s_k - eps.  ONE more tcgen05 pass
        // over the bf16 corpus, batched over all such queries, with that threshold yields a COMPLETE candidate
        // list, which K4 then finishes (its certificate holds by construction once the list is rescored).
        // Only queries with no usable bound (overflow / fewer than k candidates) take the generic fp32 scan +
        // radix select, 4 queries per corpus sweep.
        TRX_TRY(ensure_fallback_ws(ix, w.batch));
        std::vector<int32_t> h_list(nfb);
        std::vector<float> h_thr(nfb);
        TRX_CUDA(cudaMemcpyAsync(h_list.data(), w.fb_list, (size_t)nfb * 4, cudaMemcpyDeviceToHost, st));
        TRX_CUDA(cudaMemcpyAsync(h_thr.data(), w.fb_thr, (size_t)nfb * 4, cudaMemcpyDeviceToHost, st));
        TRX_CUDA(cudaStreamSynchronize(st));
        std::vector<int32_t> lite, gen;
        std::vector<float> lite_thr;
        for (uint32_t i = 0; i < nfb; i++) {
            if (h_thr[i] > -INFINITY) { lite.push_back(h_list[i]); lite_thr.push_back(h_thr[i]); }
            else gen.push_back(h_list[i]);
        }
        if (!lite.empty()) {
            const int nl = (int)lite.size();
            TRX_CUDA(cudaMemcpyAsync(ix->fb_list2, lite.data(), (size_t)nl * 4, cudaMemcpyHostToDevice, st));
            TRX_CUDA(cudaMemcpyAsync(ix->thr2, lite_thr.data(), (size_t)nl * 4, cudaMemcpyHostToDevice, st));
            gather_rows_kernel<<<nl, 128, 0, st>>>(w.qdev, ix->fb_list2, ix->d, ix->qfb);
            count_launch();
            const int32_t* exfb = nullptr;
            if (w.exdev) {
                gather_i32_kernel<<<(nl + 255) / 256, 256, 0, st>>>(w.exdev, ix->fb_list2, nl, ix->exfb);
                count_launch();
                exfb = ix->exfb;
            }
            TRX_CUDA(cudaGetLastError());
            TRX_CUDA(cudaMemsetAsync(w.fb_count, 0, 4, st));
            RescoreArgs rb = rescore_args(ix, w, k);
            rb.q32 = ix->qfb; rb.nq = nl; rb.excl = exfb; rb.qmap = ix->fb_list2;
            if (ix->second_pass) {
                // ONE more tcgen05 pass over the bf16 corpus for all of them: threshold s_k - eps (prefilter domain)
                second_pass_prep_kernel<<<nl, 128, 0, st>>>(ix->fb_list2, nl, ix->Kp, ix->metric == TRX_METRIC_L2, w.q16,
                                                            w.qnorm2, w.eps, w.eps_acc, ix->q16fb, ix->epsfb, ix->eaccfb,
                                                            ix->thr2, w.cand_cnt);
                count_launch();
                TRX_CUDA(cudaGetLastError());
                UmmaArgs u{};
                u.q16 = ix->q16fb; u.nq = nl; u.x16 = ix->x16; u.n = N; u.Kp = ix->Kp;
                u.pair = use_pair_tiling(ix, nl);
                u.mode = 1; u.thr = ix->thr2; u.cand = w.cand; u.cand_cnt = w.cand_cnt; u.cap = w.cap;
                const int nlogs = umma_grid(nl, N, ix->sm_count, u.pair, false) * 128;
                // the band [s_k - eps, s_k] can hold several times the first pass's target: size the logs for a full list
                const int log_cap = std::max(256, (int)(2.0 * (double)nl * (double)w.cap / (double)nlogs) + 64);
                TRX_TRY(ensure_hitlog(ix, w, nlogs, log_cap));
                u.log = w.hitlog; u.log_cnt = w.hitlog_cnt; u.log_cap = (int)std::min<size_t>(w.hitlog_elems / nlogs, 1u << 20);
                TRX_TRY(launch_umma(u, ix->sm_count, st));
                rb.thr = ix->thr2; rb.eps = ix->epsfb; rb.eps_acc = ix->eaccfb;
            } else {
                TRX_CUDA(cudaMemsetAsync(w.cand_cnt, 0, (size_t)nl * 4, st));
                StreamArgs a{};
                a.x = ix->x32; a.pitch = ix->d; a.n = N; a.d = ix->d;
                a.q32 = ix->qfb; a.q_pitch = ix->d; a.nq = nl;
                a.metric = ix->metric; a.bf16 = false; a.append = true;
                a.thr = ix->thr2; a.cand = w.cand; a.cand_cnt = w.cand_cnt; a.cap = w.cap;
                TRX_TRY(launch_stream(a, ix->sm_count, st));
                rb.thr = ix->neg_inf;      // complete list: certified once everything is rescored
            }
            TRX_TRY(launch_rescore(rb, st));
            TRX_CUDA(cudaMemcpyAsync(w.h_nfb, w.fb_count, 4, cudaMemcpyDeviceToHost, st));
            TRX_CUDA(cudaStreamSynchronize(st));
            const uint32_t nfb2 = *w.h_nfb;
            if (nfb2 > 0) {
                size_t g0 = gen.size();
                gen.resize(g0 + nfb2);
                TRX_CUDA(cudaMemcpy(gen.data() + g0, w.fb_list, (size_t)nfb2 * 4, cudaMemcpyDeviceToHost));
            }
            if (ix->second_pass) ix->st.queries_second_pass += nl - (int64_t)nfb2;
            else ix->st.queries_exact += nl - (int64_t)nfb2;
        }
        if (!gen.empty()) {
            const int ng = (int)gen.size();
            TRX_CUDA(cudaMemcpyAsync(ix->fb_list2, gen.data(), (size_t)ng * 4, cudaMemcpyHostToDevice, st));
            gather_rows_kernel<<<ng, 128, 0, st>>>(w.qdev, ix->fb_list2, ix->d, ix->qfb);
            count_launch();
            const int32_t* exfb = nullptr;
            if (w.exdev) {
                gather_i32_kernel<<<(ng + 255) / 256, 256, 0, st>>>(w.exdev, ix->fb_list2, ng, ix->exfb);
                count_launch();
                exfb = ix->exfb;
            }
            TRX_CUDA(cudaGetLastError());
            TRX_TRY(run_exact(ix, ix->qfb, exfb, ix->fb_list2, ng, k, w.Dd, w.Id, st));
            TRX_CUDA(cudaStreamSynchronize(st));   // `gen` (host) must outlive the async upload
        }
        if (w.timed) TRX_CUDA(cudaEventRecord(w.ev[3], st));
        TRX_TRY(send_results(ix, w, k, st));
        TRX_CUDA(cudaStreamSynchronize(st));
    }
    if (!w.out_dev && !w.out_pinned) {
        memcpy(w.D, w.hD, (size_t)B * k * 4);
        memcpy(w.I, w.hI, (size_t)B * k * 8);
    }
    if (w.timed) {
        float ms = 0.f;
        TRX_CUDA(cudaEventSynchronize(w.ev[3]));
        TRX_CUDA(cudaEventElapsedTime(&ms, w.ev[0], w.ev[3]));
        ix->st.last_total_ms = ms;
        ix->st.sum_total_ms += ms;
        ix->st.timed_batches++;
        if (w.path != TRX_PATH_EXACT) {
            TRX_CUDA(cudaEventElapsedTime(&ms, w.ev[1], w.ev[2]));
            ix->st.last_prefilter_ms = ms;
            ix->st.sum_prefilter_ms += ms;
            TRX_CUDA(cudaEventElapsedTime(&ms, w.ev[0], w.ev[1]));
            ix->st.sum_sample_ms += ms;
            if (nfb == 0) {   // ev[3] moves behind the fallbacks when there are any
                TRX_CUDA(cudaEventElapsedTime(&ms, w.ev[2], w.ev[3]));
                ix->st.sum_rescore_ms += ms;
            }
        } else ix->st.last_prefilter_ms = 0.0;
    }
    ix->st.queries += B;
    return TRX_OK;
}

// ---------------------------------------------------------------------------------------------
// C ABI
// ---------------------------------------------------------------------------------------------
extern "C" {

const char* trx_last_error(void) { return g_err.c_str(); }
const char* trx_version(void) { return "textreact_b200 libtrx 0.1 (sm_100a)"; }

int trx_create(int d, int metric, int device, trx_index** out) {
    if (!out) { set_error("out is null"); return TRX_EINVAL; }
    *out = nullptr;
    if (d <= 0 || d > 16384) { set_error("bad dimension %d", d); return TRX_EINVAL; }
    if (metric != TRX_METRIC_INNER_PRODUCT && metric != TRX_METRIC_L2) { set_error("bad metric %d", metric); return TRX_EINVAL; }
    int ndev = 0;
    if (cudaGetDeviceCount(&ndev) != cudaSuccess || ndev == 0) {
        cudaGetLastError();
        set_error("no CUDA device: this engine has no CPU fallback");
        return TRX_ENODEV;
    }
    if (device < 0 || device >= ndev) { set_error("device %d out of range (%d devices)", device, ndev); return TRX_EINVAL; }
    cudaDeviceProp prop;
    TRX_CUDA(cudaGetDeviceProperties(&prop, device));
    if (prop.major != 10) {
        set_error("device %d is sm_%d%d; libtrx is built for sm_100a (B200) only", device, prop.major, prop.minor);
        return TRX_ENODEV;
    }
    trx_index* ix = new (std::nothrow) trx_index();
    if (!ix) { set_error("host allocation failed"); return TRX_ENOMEM; }
    ix->d = d; ix->metric = metric; ix->device = device; ix->sm_count = prop.multiProcessorCount;
    if (getenv("TRX_NO_GRAPHS")) ix->graphs = 0;
    int kcols = d + (metric == TRX_METRIC_L2 ? 3 : 0);
    ix->Kp = (kcols + 63) / 64 * 64;
    DeviceGuard g(device);
    int rc = TRX_OK;
    do {
        if ((rc = dmalloc(&ix->norm2_max, 2)) != TRX_OK) break;
        if ((rc = dmalloc(&ix->counters, 4)) != TRX_OK) break;
        if (cudaMemset(ix->norm2_max, 0, 8) != cudaSuccess || cudaMemset(ix->counters, 0, 32) != cudaSuccess ||
            cudaStreamCreateWithFlags(&ix->copy_stream, cudaStreamNonBlocking) != cudaSuccess ||
            cudaStreamCreateWithFlags(&ix->own_stream, cudaStreamNonBlocking) != cudaSuccess) {
            set_error("device init failed: %s", cudaGetErrorString(cudaGetLastError()));
            rc = TRX_ECUDA; break;
        }
    } while (0);
    if (rc != TRX_OK) { trx_destroy(ix); return rc; }
    ix->st.sm_count = ix->sm_count;
    *out = ix;
    return TRX_OK;
}

void trx_destroy(trx_index* ix) {
    if (!ix) return;
    DeviceGuard g(ix->device);
    cudaDeviceSynchronize();
    free_ws(ix); free_store(ix);
    dfree(ix->norm2_max); dfree(ix->counters);
    if (ix->own_stream) cudaStreamDestroy(ix->own_stream);
    if (ix->copy_stream) cudaStreamDestroy(ix->copy_stream);
    delete ix;
}

int trx_reset(trx_index* ix) {
    if (!ix) { set_error("null index"); return TRX_EINVAL; }
    std::lock_guard<std::mutex> lock(ix->mu);
    abandon_two_phase(ix);
    ix->gen++;
    DeviceGuard g(ix->device);
    TRX_CUDA(cudaDeviceSynchronize());
    free_store(ix);
    TRX_CUDA(cudaMemset(ix->norm2_max, 0, 8));
    return TRX_OK;
}

int64_t trx_ntotal(const trx_index* ix) { return ix ? ix->ntotal : -1; }
int trx_dim(const trx_index* ix) { return ix ? ix->d : -1; }
int trx_metric(const trx_index* ix) { return ix ? ix->metric : -1; }

int trx_reserve(trx_index* ix, int64_t n) {
    if (!ix || n < 0) { set_error("bad argument"); return TRX_EINVAL; }
    std::lock_guard<std::mutex> lock(ix->mu);
    abandon_two_phase(ix);
    DeviceGuard g(ix->device);
    ix->gen++;
    return grow(ix, n);
}

int trx_add(trx_index* ix, const float* x, int64_t n) {
    if (!ix || n < 0 || (n > 0 && !x)) { set_error("bad argument"); return TRX_EINVAL; }
    if (n == 0) return TRX_OK;
    std::lock_guard<std::mutex> lock(ix->mu);
    abandon_two_phase(ix);
    ix->gen++;
    DeviceGuard g(ix->device);
    TRX_TRY(grow(ix, ix->ntotal + n));
    cudaStream_t st = ix->own_stream;
    float* dst = ix->x32 + ix->ntotal * ix->d;
    TRX_CUDA(cudaMemcpyAsync(dst, x, (size_t)n * ix->d * 4, is_device_ptr(x) ? cudaMemcpyDeviceToDevice : cudaMemcpyHostToDevice, st));
    TRX_TRY(launch_ingest(dst, n, ix->d, ix->Kp, ix->metric, ix->x16 + ix->ntotal * ix->Kp, ix->xnorm2 + ix->ntotal,
                          ix->norm2_max, st));
    TRX_CUDA(cudaStreamSynchronize(st));
    ix->ntotal += n;
    ix->sample_dirty = true;
    ix->has_groups = false;  // groups / attributes must cover every row: set them again after the last add
    ix->has_attr = false;
    return TRX_OK;
}

int trx_add_typed(trx_index* ix, const void* x, int64_t n, int dtype) {
    if (dtype == TRX_DTYPE_F32) return trx_add(ix, (const float*)x, n);
    if (!ix || n < 0 || (n > 0 && !x)) { set_error("bad argument"); return TRX_EINVAL; }
    const int es = dtype_size(dtype);
    if (es == 0) { set_error("add_typed: unknown dtype %d", dtype); return TRX_EINVAL; }
    if (n == 0) return TRX_OK;
    std::lock_guard<std::mutex> lock(ix->mu);
    abandon_two_phase(ix);
    ix->gen++;
    DeviceGuard g(ix->device);
    TRX_TRY(grow(ix, ix->ntotal + n));
    cudaStream_t st = ix->own_stream;
    const bool dev = is_device_ptr(x);
    // raw rows cross PCIe in chunks of <= 256 MB through one staging buffer, widened straight into the fp32 store
    const int64_t rows_per_chunk = std::max<int64_t>(1, ((int64_t)256 << 20) / ((int64_t)ix->d * es));
    void* stage = nullptr;
    if (!dev) {
        cudaError_t e = cudaMalloc(&stage, (size_t)std::min(rows_per_chunk, n) * ix->d * es);
        if (e != cudaSuccess) { cudaGetLastError(); set_error("cudaMalloc of the ingest staging buffer failed"); return TRX_ENOMEM; }
    }
    int rc = TRX_OK;
    for (int64_t r0 = 0; r0 < n && rc == TRX_OK; r0 += rows_per_chunk) {
        const int64_t nr = std::min(rows_per_chunk, n - r0);
        const char* src = (const char*)x + (size_t)r0 * ix->d * es;
        const void* dsrc = src;
        if (!dev) {
            if (cudaMemcpyAsync(stage, src, (size_t)nr * ix->d * es, cudaMemcpyHostToDevice, st) != cudaSuccess) {
                set_error("add_typed: H2D copy failed: %s", cudaGetErrorString(cudaGetLastError())); rc = TRX_ECUDA; break;
            }
            dsrc = stage;
        }
        float* dst = ix->x32 + (ix->ntotal + r0) * ix->d;
        rc = launch_widen(dsrc, dtype, nr * ix->d, dst, st);
        if (rc == TRX_OK)
            rc = launch_ingest(dst, nr, ix->d, ix->Kp, ix->metric, ix->x16 + (ix->ntotal + r0) * ix->Kp,
                               ix->xnorm2 + ix->ntotal + r0, ix->norm2_max, st);
    }
    cudaError_t e = cudaStreamSynchronize(st);
    if (stage) cudaFree(stage);
    if (rc != TRX_OK) return rc;
    if (e != cudaSuccess) { set_error("add_typed: %s", cudaGetErrorString(e)); return TRX_ECUDA; }
    ix->ntotal += n;
    ix->sample_dirty = true;
    ix->has_groups = false;
    ix->has_attr = false;
    return TRX_OK;
}

int trx_set_groups(trx_index* ix, const int32_t* gsrc, int64_t n) {
    if (!ix) { set_error("null index"); return TRX_EINVAL; }
    std::lock_guard<std::mutex> lock(ix->mu);
    abandon_two_phase(ix);
    ix->group_max = 0; ix->group_avg = 1.0; ix->gen++;
    if (gsrc == nullptr) { ix->has_groups = false; return TRX_OK; }
    if (n != ix->ntotal) { set_error("set_groups: n=%lld != ntotal=%lld", (long long)n, (long long)ix->ntotal); return TRX_EINVAL; }
    if (n == 0) return TRX_OK;
    DeviceGuard g(ix->device);
    TRX_CUDA(cudaMemcpy(ix->groups, gsrc, (size_t)n * 4, is_device_ptr(gsrc) ? cudaMemcpyDeviceToDevice : cudaMemcpyHostToDevice));
    ix->has_groups = true;
    return TRX_OK;
}

int trx_set_row_attr(trx_index* ix, const int32_t* asrc, int64_t n) {
    if (!ix) { set_error("null index"); return TRX_EINVAL; }
    std::lock_guard<std::mutex> lock(ix->mu);
    abandon_two_phase(ix);
    ix->gen++;
    if (asrc == nullptr) { ix->has_attr = false; ix->attr_sorted.clear(); return TRX_OK; }
    if (n != ix->ntotal) { set_error("set_row_attr: n=%lld != ntotal=%lld", (long long)n, (long long)ix->ntotal); return TRX_EINVAL; }
    if (n == 0) return TRX_OK;
    DeviceGuard g(ix->device);
    ix->has_attr = false;            // stays off if anything below fails
    ix->attr_sorted.clear();
    dfree(ix->attr);
    TRX_TRY(dmalloc(&ix->attr, (size_t)n));
    const bool dev = is_device_ptr(asrc);
    TRX_CUDA(cudaMemcpy(ix->attr, asrc, (size_t)n * 4, dev ? cudaMemcpyDeviceToDevice : cudaMemcpyHostToDevice));
    try { ix->attr_sorted.resize((size_t)n); } catch (...) { set_error("host allocation failed"); return TRX_ENOMEM; }
    if (dev) TRX_CUDA(cudaMemcpy(ix->attr_sorted.data(), asrc, (size_t)n * 4, cudaMemcpyDeviceToHost));
    else memcpy(ix->attr_sorted.data(), asrc, (size_t)n * 4);
    std::sort(ix->attr_sorted.begin(), ix->attr_sorted.end());
    ix->has_attr = true;
    return TRX_OK;
}

static int search_impl(trx_index* ix, const float* xq, int64_t nq, int k, const trx_search_params_t* sp, float* D,
                       int64_t* I, void* cuda_stream);

int trx_search_self(trx_index* ix, int64_t row0, int64_t nq, int k, const int32_t* excl, float* D, int64_t* I,
                    void* cuda_stream) {
    if (!ix) { set_error("null index"); return TRX_EINVAL; }
    if (row0 < 0) { set_error("search_self: negative row"); return TRX_EINVAL; }
    trx_search_params_t sp;
    sp.exclude = excl; sp.attr_below = ix->attr_below; sp.dedup_groups = ix->dedup; sp.self_row0 = row0;
    return search_impl(ix, nullptr, nq, k, &sp, D, I, cuda_stream);
}

int trx_reconstruct(trx_index* ix, int64_t row0, int64_t n, float* out) {
    if (!ix || !out || row0 < 0 || n < 0 || row0 + n > ix->ntotal) { set_error("reconstruct: bad rows [%lld, %lld)", (long long)row0, (long long)(row0 + n)); return TRX_EINVAL; }
    if (n == 0) return TRX_OK;
    std::lock_guard<std::mutex> lock(ix->mu);
    DeviceGuard g(ix->device);
    TRX_CUDA(cudaMemcpy(out, ix->x32 + row0 * ix->d, (size_t)n * ix->d * 4,
                        is_device_ptr(out) ? cudaMemcpyDeviceToDevice : cudaMemcpyDeviceToHost));
    return TRX_OK;
}

int trx_set_id_offset(trx_index* ix, int64_t offset) {
    if (!ix) { set_error("null index"); return TRX_EINVAL; }
    std::lock_guard<std::mutex> lock(ix->mu);
    abandon_two_phase(ix);
    ix->id_offset = offset; ix->gen++;
    return TRX_OK;
}

// Restores the per-call modes (attr_below, dedup) that trx_search_ex overrides for the duration of one call.
struct ModeGuard {
    trx_index* ix; int32_t attr; int dedup;
    explicit ModeGuard(trx_index* i) : ix(i), attr(i->attr_below), dedup(i->dedup) {}
    ~ModeGuard() { ix->attr_below = attr; ix->dedup = dedup; }
};

static int search_impl(trx_index* ix, const float* xq, int64_t nq, int k, const trx_search_params_t* sp, float* D,
                       int64_t* I, void* cuda_stream) {
    if (!ix) { set_error("null index"); return TRX_EINVAL; }
    if (nq < 0 || k <= 0) { set_error("bad nq=%lld or k=%d", (long long)nq, k); return TRX_EINVAL; }
    if (nq == 0) return TRX_OK;
    std::lock_guard<std::mutex> lock(ix->mu);
    if (ix->ws[0].two_phase) { set_error("a two-phase search is pending (trx_search_begin without trx_search_finish)"); return TRX_EINVAL; }
    ModeGuard modes(ix);
    const int32_t* excl = nullptr;
    if (sp) {
        excl = sp->exclude;
        ix->attr_below = sp->attr_below;
        ix->dedup = sp->dedup_groups != 0;
        if (sp->self_row0 >= 0) {
            if (sp->self_row0 + nq > ix->ntotal) {
                set_error("search_self: rows [%lld, %lld) outside [0, %lld)", (long long)sp->self_row0,
                          (long long)(sp->self_row0 + nq), (long long)ix->ntotal);
                return TRX_EINVAL;
            }
            xq = ix->x32 + sp->self_row0 * ix->d;
        }
    }
    if (!xq || !D || !I) { set_error("null buffer"); return TRX_EINVAL; }
    if (k > 2048) { set_error("k=%d exceeds the supported maximum 2048", k); return TRX_EINVAL; }
    if (excl && !ix->has_groups) { set_error("exclude given but no groups set (trx_set_groups)"); return TRX_EINVAL; }
    if (ix->dedup && !ix->has_groups) { set_error("dedup_groups set but no groups set (trx_set_groups)"); return TRX_EINVAL; }
    if (ix->attr_below != INT32_MAX && !ix->has_attr) { set_error("attr_below set but no row attributes (trx_set_row_attr)"); return TRX_EINVAL; }
    DeviceGuard g(ix->device);
    const bool xq_dev = is_device_ptr(xq), out_dev = is_device_ptr(D), excl_dev = is_device_ptr(excl);
    if (out_dev != is_device_ptr(I)) { set_error("D and I must both be host or both be device pointers"); return TRX_EINVAL; }
    cudaStream_t st = cuda_stream ? (cudaStream_t)cuda_stream : ix->own_stream;
    ix->st.searches++;
    TRX_TRY(ensure_group_stats(ix));
    if (ix->ntotal == 0) {  // FAISS: empty index -> all -1
        std::vector<float> dfill((size_t)nq * k, ix->metric == TRX_METRIC_L2 ? FLT_MAX : -FLT_MAX);
        std::vector<int64_t> ifill((size_t)nq * k, -1);
        TRX_CUDA(cudaMemcpy(D, dfill.data(), dfill.size() * 4, out_dev ? cudaMemcpyHostToDevice : cudaMemcpyHostToHost));
        TRX_CUDA(cudaMemcpy(I, ifill.data(), ifill.size() * 8, out_dev ? cudaMemcpyHostToDevice : cudaMemcpyHostToHost));
        return TRX_OK;
    }
    // Software pipeline over the batches of the call: batch i+1 is queued (upload on the copy stream, kernels
    // behind batch i) before the host waits for batch i, so the GPU never idles between batches and host
    // copies to / from pageable memory overlap compute.
    const int64_t mb = ix->max_batch;
    const int64_t nb = (nq + mb - 1) / mb;
    auto launch = [&](int64_t i) {
        const int64_t q0 = i * mb, B = std::min<int64_t>(mb, nq - q0);
        return launch_batch(ix, ix->ws[i & 1], xq + q0 * ix->d, xq_dev, B, k, excl ? excl + q0 : nullptr, excl_dev,
                            D + q0 * k, I + q0 * k, out_dev, st);
    };
    int rc = launch(0);
    for (int64_t i = 0; rc == TRX_OK && i < nb; i++) {
        if (ix->pipeline && i + 1 < nb) rc = launch(i + 1);
        int rf = finish_batch(ix, ix->ws[i & 1], k, st);      // always drain what was queued
        if (rc == TRX_OK) rc = rf;
        if (rc == TRX_OK && !ix->pipeline && i + 1 < nb) rc = launch(i + 1);
    }
    if (rc != TRX_OK) cudaStreamSynchronize(st);
    return rc;
}

int trx_search(trx_index* ix, const float* xq, int64_t nq, int k, const int32_t* excl, float* D, int64_t* I,
               void* cuda_stream) {
    if (!ix) { set_error("null index"); return TRX_EINVAL; }
    trx_search_params_t sp;
    sp.exclude = excl; sp.attr_below = ix->attr_below; sp.dedup_groups = ix->dedup; sp.self_row0 = -1;
    return search_impl(ix, xq, nq, k, &sp, D, I, cuda_stream);
}

int trx_search_ex(trx_index* ix, const float* xq, int64_t nq, int k, const trx_search_params_t* params, float* D,
                  int64_t* I, void* cuda_stream) {
    if (!params) return trx_search(ix, xq, nq, k, nullptr, D, I, cuda_stream);
    return search_impl(ix, xq, nq, k, params, D, I, cuda_stream);
}

// ---- two-phase search for the row-sharded mode -------------------------------------------------------------------
// begin:  prefilter of ONE batch up to the (masked, sorted) candidate lists; payload[q] = the nb best prefilter scores
//         of the query on this shard + its eps.  The shards exchange the payloads (trx_exchange_floor) ...
// finish: ... and K4 rescores only what can still reach the GLOBAL top-k (rows at or above floor[q]): at G shards
//         that is ~1/G of what the local top-k would need.  Results may hold fewer than k rows per query (padded);
//         the merge of the shards' results is the exact global top-k.
int trx_search_begin(trx_index* ix, const float* xq, int64_t nq, int k, const trx_search_params_t* sp, int nb,
                     float* payload, void* cuda_stream) {
    if (!ix) { set_error("null index"); return TRX_EINVAL; }
    if (nq <= 0 || k <= 0 || nb <= 0 || nb > 4096 || !payload) { set_error("search_begin: bad nq=%lld, k=%d or nb=%d", (long long)nq, k, nb); return TRX_EINVAL; }
    std::lock_guard<std::mutex> lock(ix->mu);
    if (nq > ix->max_batch) { set_error("search_begin takes one batch (nq=%lld > max_batch=%d)", (long long)nq, ix->max_batch); return TRX_EINVAL; }
    if (ix->ntotal == 0) { set_error("search_begin on an empty index"); return TRX_EINVAL; }
    if (ix->ws[0].two_phase) { set_error("search_begin: the previous two-phase search was not finished"); return TRX_EINVAL; }
    ModeGuard modes(ix);
    const int32_t* excl = nullptr;
    if (sp) {
        excl = sp->exclude;
        ix->attr_below = sp->attr_below;
        ix->dedup = sp->dedup_groups != 0;
        if (sp->self_row0 >= 0) {
            if (sp->self_row0 + nq > ix->ntotal) { set_error("search_begin: self rows out of range"); return TRX_EINVAL; }
            xq = ix->x32 + sp->self_row0 * ix->d;
        }
    }
    if (!xq) { set_error("null buffer"); return TRX_EINVAL; }
    if (k > 2048) { set_error("k=%d exceeds the supported maximum 2048", k); return TRX_EINVAL; }
    if (excl && !ix->has_groups) { set_error("exclude given but no groups set (trx_set_groups)"); return TRX_EINVAL; }
    if (ix->dedup && !ix->has_groups) { set_error("dedup_groups set but no groups set (trx_set_groups)"); return TRX_EINVAL; }
    if (ix->attr_below != INT32_MAX && !ix->has_attr) { set_error("attr_below set but no row attributes (trx_set_row_attr)"); return TRX_EINVAL; }
    if (!is_device_ptr(payload)) { set_error("search_begin: payload must be a device pointer"); return TRX_EINVAL; }
    DeviceGuard g(ix->device);
    cudaStream_t st = cuda_stream ? (cudaStream_t)cuda_stream : ix->own_stream;
    ix->st.searches++;
    TRX_TRY(ensure_group_stats(ix));
    return launch_batch(ix, ix->ws[0], xq, is_device_ptr(xq), nq, k, excl, is_device_ptr(excl), nullptr, nullptr, true, st,
                        nb, payload);
}

int trx_search_finish(trx_index* ix, const float* floor, float* D, int64_t* I, void* cuda_stream) {
    if (!ix || !D || !I) { set_error("bad argument"); return TRX_EINVAL; }
    std::lock_guard<std::mutex> lock(ix->mu);
    BatchWs& w = ix->ws[0];
    if (!w.two_phase) { set_error("search_finish without search_begin"); return TRX_EINVAL; }
    w.two_phase = false;
    if (floor && !is_device_ptr(floor)) { set_error("search_finish: floor must be a device pointer"); return TRX_EINVAL; }
    const bool out_dev = is_device_ptr(D);
    if (out_dev != is_device_ptr(I)) { set_error("D and I must both be host or both be device pointers"); return TRX_EINVAL; }
    ModeGuard modes(ix);
    ix->attr_below = w.tp_attr; ix->dedup = w.tp_dedup;
    DeviceGuard g(ix->device);
    cudaStream_t st = cuda_stream ? (cudaStream_t)cuda_stream : ix->own_stream;
    const int k = w.tp_k;
    w.D = D; w.I = I; w.out_dev = out_dev;
    w.out_pinned = !out_dev && is_pinned_host_ptr(D) && is_pinned_host_ptr(I);
    if (w.path == TRX_PATH_EXACT) {
        TRX_TRY(run_exact(ix, w.qdev, w.exdev, nullptr, w.B, k, w.Dd, w.Id, st));
        *w.h_nfb = 0;
    } else {
        RescoreArgs ra = rescore_args(ix, w, k);
        ra.presorted = 1; ra.floor = floor;
        TRX_TRY(launch_rescore(ra, st));
        TRX_CUDA(cudaMemcpyAsync(w.h_nfb, w.fb_count, 4, cudaMemcpyDeviceToHost, st));
    }
    if (w.timed) TRX_CUDA(cudaEventRecord(w.ev[3], st));
    TRX_TRY(send_results(ix, w, k, st));
    TRX_CUDA(cudaEventRecord(w.done, st));
    int rc = finish_batch(ix, w, k, st);
    if (rc != TRX_OK) cudaStreamSynchronize(st);
    return rc;
}

int trx_set_option(trx_index* ix, const char* key, double v) {
    if (!ix || !key) { set_error("bad argument"); return TRX_EINVAL; }
    std::lock_guard<std::mutex> lock(ix->mu);
    {   // a cached graph stays valid when the option keeps its value (per-call modes toggle options on and off)
        double old = 0.0;
        if (trx_get_option(ix, key, &old) != TRX_OK || old != v) ix->gen++;
    }
    if (!strcmp(key, "path")) {
        if (v < 0 || v > 3) { set_error("path must be 0..3"); return TRX_EINVAL; }
        ix->opt_path = (int)v;
    } else if (!strcmp(key, "max_batch")) {
        if (v < 1 || v > 65536) { set_error("max_batch out of range"); return TRX_EINVAL; }
        ix->max_batch = (int)v;
    } else if (!strcmp(key, "target_candidates")) {
        // K4 keeps 4x the target (rounded up to a power of two) as 16-byte keys in shared memory: 2048 -> 8192 entries
        if (v < 32 || v > 2048) { set_error("target_candidates out of range (32..2048)"); return TRX_EINVAL; }
        ix->target = (int)v;
    } else if (!strcmp(key, "sample_rate")) {
        if (v < 2 || v > 1024) { set_error("sample_rate out of range"); return TRX_EINVAL; }
        ix->sample_rate = (int)v; ix->sample_dirty = true;
    } else if (!strcmp(key, "stream_max_batch")) {
        ix->stream_max_batch = (int)v;
    } else if (!strcmp(key, "timing")) {
        ix->timing = v != 0;
    } else if (!strcmp(key, "pipeline")) {
        ix->pipeline = v != 0;
    } else if (!strcmp(key, "dedup_groups")) {
        ix->dedup = v != 0;
    } else if (!strcmp(key, "second_pass")) {
        ix->second_pass = v != 0;
    } else if (!strcmp(key, "thr_margin")) {
        ix->thr_margin = v != 0;
    } else if (!strcmp(key, "graphs")) {
        ix->graphs = v != 0;
    } else if (!strcmp(key, "graph_max_batch")) {
        ix->graph_max_batch = (int)v;
    } else if (!strcmp(key, "attr_below")) {
        ix->attr_below = v >= 2147483647.0 ? INT32_MAX : (v <= -2147483648.0 ? INT32_MIN : (int32_t)v);
    } else if (!strcmp(key, "thr_bias")) {
        ix->thr_bias = (float)v;
    } else if (!strcmp(key, "umma_pair")) {
        ix->umma_pair = v != 0;
    } else if (!strcmp(key, "pair_min_batch")) {
        if (v < 1) { set_error("pair_min_batch out of range"); return TRX_EINVAL; }
        ix->pair_min_batch = (int)v;
    } else { set_error("unknown option '%s'", key); return TRX_EINVAL; }
    return TRX_OK;
}

int trx_get_option(const trx_index* ix, const char* key, double* v) {
    if (!ix || !key || !v) { set_error("bad argument"); return TRX_EINVAL; }
    if (!strcmp(key, "path")) *v = ix->opt_path;
    else if (!strcmp(key, "max_batch")) *v = ix->max_batch;
    else if (!strcmp(key, "target_candidates")) *v = ix->target;
    else if (!strcmp(key, "sample_rate")) *v = ix->sample_rate;
    else if (!strcmp(key, "stream_max_batch")) *v = ix->stream_max_batch;
    else if (!strcmp(key, "timing")) *v = ix->timing;
    else if (!strcmp(key, "umma_pair")) *v = ix->umma_pair;
    else if (!strcmp(key, "pair_min_batch")) *v = ix->pair_min_batch;
    else if (!strcmp(key, "pipeline")) *v = ix->pipeline;
    else if (!strcmp(key, "attr_below")) *v = ix->attr_below;
    else if (!strcmp(key, "dedup_groups")) *v = ix->dedup;
    else if (!strcmp(key, "second_pass")) *v = ix->second_pass;
    else if (!strcmp(key, "thr_margin")) *v = ix->thr_margin;
    else if (!strcmp(key, "graphs")) *v = ix->graphs;
    else if (!strcmp(key, "graph_max_batch")) *v = ix->graph_max_batch;
    else if (!strcmp(key, "graph_replays")) *v = (double)ix->graph_replays;
    else if (!strcmp(key, "graph_captures")) *v = (double)ix->graph_captures;
    else { set_error("unknown option '%s'", key); return TRX_EINVAL; }
    return TRX_OK;
}

int trx_stats(const trx_index* ix, trx_stats_t* out) {
    if (!ix || !out) { set_error("bad argument"); return TRX_EINVAL; }
    DeviceGuard g(ix->device);
    uint64_t c[4];
    TRX_CUDA(cudaMemcpy(c, ix->counters, 32, cudaMemcpyDeviceToHost));
    *out = ix->st;
    out->rescored = (int64_t)c[0]; out->queries_uncert = (int64_t)c[1]; out->queries_overflow = (int64_t)c[2];
    out->candidates = (int64_t)c[3];
    out->launches = g_launches.load();
    return TRX_OK;
}

int trx_merge_topk(int metric, const float* Dg, const int64_t* Ig, int G, int64_t nq, int k, float* D, int64_t* I,
                   void* cuda_stream) {
    if (G <= 0 || nq < 0 || k <= 0 || !Dg || !Ig || !D || !I) { set_error("bad argument"); return TRX_EINVAL; }
    if (!is_device_ptr(Dg) || !is_device_ptr(Ig) || !is_device_ptr(D) || !is_device_ptr(I)) {
        set_error("trx_merge_topk takes device pointers");
        return TRX_EINVAL;
    }
    cudaPointerAttributes at;        // launch on the device that owns the lists, whatever the caller's current device is
    TRX_CUDA(cudaPointerGetAttributes(&at, Dg));
    DeviceGuard g(at.device);
    return launch_merge(metric, Dg, Ig, G, nq, k, D, I, (cudaStream_t)cuda_stream);
}

// Plain device buffers for callers that have no CUDA runtime of their own (C programs, numpy-only test drivers).
int trx_device_malloc(int device, size_t bytes, void** out) {
    if (!out) { set_error("out is null"); return TRX_EINVAL; }
    DeviceGuard g(device);
    cudaError_t e = cudaMalloc(out, bytes);
    if (e != cudaSuccess) { cudaGetLastError(); *out = nullptr; set_error("cudaMalloc of %zu bytes failed: %s", bytes, cudaGetErrorString(e)); return TRX_ENOMEM; }
    return TRX_OK;
}
int trx_device_free(void* p) {
    if (p) TRX_CUDA(cudaFree(p));
    return TRX_OK;
}
int trx_device_copy(void* dst, const void* src, size_t bytes) {   // host <-> device in any direction, synchronous
    if (bytes && (!dst || !src)) { set_error("null buffer"); return TRX_EINVAL; }
    TRX_CUDA(cudaDeviceSynchronize());     // the index's streams are non-blocking: wait for whatever produces `src`
    TRX_CUDA(cudaMemcpy(dst, src, bytes, cudaMemcpyDefault));
    return TRX_OK;
}

int trx_debug_scores_umma(trx_index* ix, const float* xq, int64_t nq, int64_t row0, int64_t n, float* out,
                          void* cuda_stream) {
    if (!ix || !xq || !out || nq <= 0 || n <= 0 || row0 < 0 || row0 + n > ix->ntotal) { set_error("bad argument"); return TRX_EINVAL; }
    if (!is_device_ptr(xq) || !is_device_ptr(out)) { set_error("debug_scores takes device pointers"); return TRX_EINVAL; }
    std::lock_guard<std::mutex> lock(ix->mu);
    DeviceGuard g(ix->device);
    cudaStream_t st = cuda_stream ? (cudaStream_t)cuda_stream : ix->own_stream;
    BatchWs& w = ix->ws[0];
    TRX_TRY(ensure_ws(ix, w, (int)nq, 1, candidate_cap(ix, 1)));
    TRX_TRY(launch_query_prep(xq, nq, ix->d, ix->Kp, ix->metric, w.q16, w.qnorm2, nullptr, nullptr, nullptr, nullptr,
                              nullptr, st));
    UmmaArgs u{};
    u.q16 = w.q16; u.nq = nq; u.x16 = ix->x16 + row0 * ix->Kp; u.n = n; u.Kp = ix->Kp;
    u.mode = 0; u.out = out; u.out_ld = n;
    u.pair = ix->umma_pair && nq >= ix->pair_min_batch;
    TRX_TRY(launch_umma(u, ix->sm_count, st));
    TRX_CUDA(cudaStreamSynchronize(st));
    return TRX_OK;
}

}  // extern "C"
