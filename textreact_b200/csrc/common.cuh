// common.cuh -- shared declarations for libtrx.so (B200 / sm_100a only).
#pragma once

#include <cuda.h>
#include <cuda_bf16.h>
#include <cuda_runtime.h>
#include <stdint.h>

#include "../../include/trx.h"

namespace trx {

void set_error(const char* fmt, ...);

#define TRX_CUDA(call)                                                                      \
    do {                                                                                    \
        cudaError_t e_ = (call);                                                            \
        if (e_ != cudaSuccess) {                                                            \
            ::trx::set_error("%s:%d %s: %s", __FILE__, __LINE__, #call, cudaGetErrorString(e_)); \
            return (e_ == cudaErrorMemoryAllocation) ? TRX_ENOMEM : TRX_ECUDA;              \
        }                                                                                   \
    } while (0)

#define TRX_TRY(call)              \
    do {                           \
        int rc_ = (call);          \
        if (rc_ != TRX_OK) return rc_; \
    } while (0)

// Monotone float -> uint32: a > b  <=>  f2key(a) > f2key(b)  (-0 < +0, NaNs at the ends).
__host__ __device__ __forceinline__ uint32_t f2key(float f) {
#ifdef __CUDA_ARCH__
    uint32_t u = __float_as_uint(f);
#else
    union { float f; uint32_t u; } c; c.f = f; uint32_t u = c.u;
#endif
    return (u & 0x80000000u) ? ~u : (u | 0x80000000u);
}
__host__ __device__ __forceinline__ float key2f(uint32_t k) {
    uint32_t u = (k & 0x80000000u) ? (k & 0x7fffffffu) : ~k;
#ifdef __CUDA_ARCH__
    return __uint_as_float(u);
#else
    union { float f; uint32_t u; } c; c.u = u; return c.f;
#endif
}

// 64-bit sort key: ascending order == (score descending, id ascending).
__host__ __device__ __forceinline__ uint64_t pack_key(float score, uint32_t id) {
    return ((uint64_t)(~f2key(score)) << 32) | (uint64_t)id;
}
__host__ __device__ __forceinline__ float key_score(uint64_t k) { return key2f(~(uint32_t)(k >> 32)); }
__host__ __device__ __forceinline__ uint32_t key_id(uint64_t k) { return (uint32_t)k; }
static constexpr uint64_t KEY_SENTINEL = 0xffffffffffffffffull;  // sorts last

#ifdef __CUDACC__
// In-place ascending bitonic sort of P (power of two) 64-bit keys in shared memory, by the whole CTA.
__device__ __forceinline__ void bitonic_sort_u64(uint64_t* keys, int P) {
    for (int size = 2; size <= P; size <<= 1) {
        for (int stride = size >> 1; stride > 0; stride >>= 1) {
            __syncthreads();
            for (int t = threadIdx.x; t < (P >> 1); t += blockDim.x) {
                int i = 2 * t - (t & (stride - 1));
                int j = i + stride;
                uint64_t a = keys[i], b = keys[j];
                bool up = (i & size) == 0;
                if ((a > b) == up) { keys[i] = b; keys[j] = a; }
            }
        }
    }
    __syncthreads();
}
#endif

// A prefilter candidate: bf16-pipeline score + local row.
struct __align__(8) Cand {
    float score;
    int32_t row;
};

// One threshold hit logged by a K2 epilogue thread.
struct __align__(16) HitRec {
    int32_t q;     // query (row of the batch)
    int32_t row;   // corpus row (local)
    float score;
    int32_t pad;
};

// ---- host launchers (each returns TRX_*) ---------------------------------------------------

// K1: fp32 rows -> bf16 rows (pitch Kp, zero padded; L2: |x|^2 split in 3 bf16 at cols d..d+2),
// squared norms; norm2_max_bits[0] = running max of |x|^2, [1] = running max of |x - bf16(x)|^2 (float bits).
int launch_ingest(const float* x, int64_t n, int d, int Kp, int metric, __nv_bfloat16* x16,
                  float* xnorm2, uint32_t* norm2_max_bits, cudaStream_t st);
// Raw typed array (TRX_DTYPE_*) -> fp32, elementwise, on the device (same values as numpy's astype(float32)).
int launch_widen(const void* src, int dtype, int64_t count, float* dst, cudaStream_t st);
int dtype_size(int dtype);  // bytes per element, 0 for an unknown code
// One pseudo-randomly chosen row out of every `rate` consecutive rows -> xs16 [ns, Kp].
int launch_sample_gather(const __nv_bfloat16* x16, int64_t n, int Kp, int rate, __nv_bfloat16* xs16,
                         int64_t ns, cudaStream_t st);
// Start of a batch: queries fp32 [B,d] -> bf16 [B,Kp] (L2: 2q and -1,-1,-1 in the norm columns) + |q|^2;
// when given (nullable as a group): certificate slack eps[q] = c * |q| * max|x| (IP; L2 doubles it and adds
// norm slack) and the fp32 summation slack eps_acc[q]; cand_cnt[q] = 0; *fb_count = 0.
int launch_query_prep(const float* q, int64_t B, int d, int Kp, int metric, __nv_bfloat16* q16,
                      float* qnorm2, const uint32_t* norm2_max_bits, float* eps, float* eps_acc,
                      uint32_t* cand_cnt, uint32_t* fb_count, cudaStream_t st);

// K3: CUDA-core streaming scorer.
//   fp32 mode : exact scores (IP: q.x ; L2: -sum (q-x)^2), rows masked by group get -inf.
//   bf16 mode : prefilter scores over the Kp-wide augmented rows (same formula as K2).
// scores mode writes out[q*out_ld + row]; append mode emits candidates with score > thr[q].
struct StreamArgs {
    const void* x; int64_t pitch; int64_t n; int d;   // d = number of columns to reduce over
    const float* q32; const __nv_bfloat16* q16; int64_t q_pitch; int64_t nq;
    const int32_t* groups; const int32_t* excl;
    const int32_t* attr; int32_t attr_below;          // scores mode: rows with attr >= attr_below are ineligible
    float* out; int64_t out_ld;                       // scores mode
    const float* thr; Cand* cand; uint32_t* cand_cnt; int cap;  // append mode
    int metric; bool bf16; bool append;
};
int launch_stream(const StreamArgs& a, int sm_count, cudaStream_t st);

// r-th largest value of each row of s[nq][ld] (first n entries) -> thr[nq]; rows with fewer
// than r finite entries get -inf.
int launch_row_kth(const float* s, int64_t ld, int64_t n, int64_t nq, int r, float* thr,
                   cudaStream_t st, const float* margin = nullptr, int ks = 0);

// Exact top-k of materialised scores s[nq][ld] ("larger is better", -inf = ineligible):
// (score desc, id asc), padded with id -1.  negate_out: D = -score (L2).  qmap (nullable):
// output row of query i is qmap[i].
int launch_exact_topk(const float* s, int64_t ld, int64_t n, int64_t nq, int k, bool negate_out,
                      int64_t id_offset, const int32_t* qmap, float* D, int64_t* I, cudaStream_t st);

// K4: candidates -> exact fp32 rescore -> certificate -> final top-k.
struct RescoreArgs {
    const Cand* cand; const uint32_t* cand_cnt; int cap;
    const float* thr;          // prefilter threshold per query (-inf: every row is a candidate)
    const float* eps;          // certificate slack per query (score units)
    const float* x32; int d; int64_t n;
    const float* q32; int64_t nq;
    const int32_t* groups; const int32_t* excl;
    const int32_t* attr; int32_t attr_below;   // nullable: rows with attr[row] >= attr_below are ineligible
    int dedup;                                 // 1: distinct-groups mode -- only the best row of a group is returned
    int k; int metric; int64_t id_offset;
    float* D; int64_t* I;
    int32_t* fb_list; uint32_t* fb_count;  // queries that need the exact scan
    float* fb_thr;      // per fb_list entry: exact score every true top-k row must reach (-inf: unknown)
    const float* eps_acc;  // fp32 summation slack per query (subtracted from fb_thr)
    const int32_t* qmap;   // nullable: output row / fb_list value of query i is qmap[i]
    uint64_t* counters;                     // [0] rescored rows [1] uncertified [2] overflow [3] candidates
    // two-phase (row-sharded) search: the lists were filtered and sorted by launch_bounds (no masks / sort here), and
    // floor[q] (nullable, prefilter domain) is a score no row of the GLOBAL top-k can fall below: rows under it are
    // dropped, and a list that is rescored down to the floor is complete -- possibly with fewer than k results.
    int presorted; const float* floor;
};
int launch_rescore(const RescoreArgs& a, cudaStream_t st);
// First half of a two-phase search: per query, the candidate list with the masks applied, sorted by prefilter score
// (written back in place, cand_cnt = eligible count), and payload[q] = { its nb best prefilter scores (-inf padded),
// eps[q] } -- what the shards exchange to bound the global k-th score.  Overflowed lists are left alone (-inf row).
struct BoundsArgs {
    Cand* cand; uint32_t* cand_cnt; int cap; int64_t nq;
    const int32_t* groups; const int32_t* excl; const int32_t* attr; int32_t attr_below;
    const float* eps; int nb; float* payload;   // [nq][nb + 1]
};
int launch_bounds(const BoundsArgs& a, cudaStream_t st);
// dynamic shared memory one K4 CTA needs; the prefilter paths are only taken when it fits kK4MaxSmem
constexpr size_t kK4MaxSmem = 200 * 1024;
inline size_t k4_smem_bytes(int cap, int d, int k, bool dedup) {
    size_t smem = (size_t)cap * 16 + (size_t)((d + 3) & ~3) * 4;
    if (dedup) smem += (size_t)cap * 4 + (size_t)k * 4 + (size_t)cap;
    return smem;
}

// K2: tcgen05 scoring GEMM.  A = queries bf16 [nq, Kp], B = rows bf16 [n, Kp].
struct UmmaArgs {
    const __nv_bfloat16* q16; int64_t nq;          // q16 must be allocated for round_up(nq, 8) rows
    const __nv_bfloat16* x16; int64_t n; int Kp;
    // mode 0: store fp32 scores out[q*out_ld + row]; mode 1: append candidates > thr[q];
    // mode 2: slot maxima -> out[(q*S + slice)*32 + slot]
    int mode;
    float* out; int64_t out_ld;
    const float* thr; Cand* cand; uint32_t* cand_cnt; int cap;
    HitRec* log; uint32_t* log_cnt; int log_cap;   // mode 1: [umma_grid*128][log_cap] private hit logs
    bool pair;                                     // CTA-pair tiling (tcgen05 cta_group::2, 256-query tiles)
};
int launch_umma(const UmmaArgs& a, int sm_count, cudaStream_t st);
int umma_grid(int64_t nq, int64_t n, int sm_count, bool pair, bool slotmax);  // CTAs a launch will use
int umma_init();  // resolves cuTensorMapEncodeTiled
int umma_num_slices(int64_t n, int64_t nq, int sm_count, bool pair);  // S of the SLOTMAX mode (out = slots[nq][S][32])
// r-th largest of the S*32 slot maxima of each query -> thr (any S; S <= 8 stays in one warp's registers)
// margin (nullable, per query: eps(q)) and ks = ceil(k / sample rate): the ks-th largest sample score estimates the
// query's k-th best score s_k, and the certificate needs  s_k > thr + eps.  The threshold is therefore kept at least
// 2.5 eps under that estimate:  thr = min(r-th largest, ks-th largest - 2.5 eps).  Where the score spread dwarfs eps
// (Gaussian data: the r-th largest is ~10 eps under the ks-th) this changes nothing; where eps is about one sigma of
// the scores (clustered unit-norm data) it ends the misses of a threshold that the sample put too high.
int launch_slot_thr(const float* slots, int64_t nq, int S, int r, float* thr, cudaStream_t st,
                    const float* margin = nullptr, int ks = 0);

// Exact path, distinct-groups mode: Dx/Ix [nq, kx] sorted results (global ids) -> first k group leaders per row.
// nfound / unfinished (nullable together): round state of the widened scan (see k4_select.cu).
int launch_dedup_rows(const float* Dx, const int64_t* Ix, int kx, const int32_t* groups, int64_t id_offset, int k,
                      bool l2, const int32_t* qmap, int64_t nq, float* D, int64_t* I, int32_t* nfound,
                      uint32_t* unfinished, cudaStream_t st);
int launch_mask_seen_groups(float* scores, int64_t ld, int64_t n, const int64_t* Ix, int kx, const int32_t* groups,
                            int64_t id_offset, const int32_t* nfound, int k, int64_t nq, cudaStream_t st);

// K5: merge G sorted lists per query.
int launch_merge(int metric, const float* Dg, const int64_t* Ig, int G, int64_t nq, int k, float* D,
                 int64_t* I, cudaStream_t st);

void count_launch(int n = 1);

}  // namespace trx
