// k4_select.cu -- selection kernels: threshold estimation, exact fp32 rescore + certificate,
// exact radix top-k over materialised scores, and the k-way shard merge.
//
// Together with the scorers (k2/k3) these replace the "heap / reservoir" half of faiss
// `index.search` (retrieve/retrieve_faiss.py:71).  Ordering everywhere: score descending
// ("larger is better": IP score, or minus squared distance for L2), ties by ascending id --
// the (val, id) comparison of FAISS heaps.
#include <float.h>
#include <limits.h>

#include <algorithm>

#include "common.cuh"

// Gather shape of the K4 rescore loop, compile-time knobs for measurement.  Round 2 measured unroll 2/3/6 x RW 2/3/4 x
// min-blocks 2/3 at C2: K4 stays at 0.73-0.75 ms whatever the shape -- the gather is not what bounds it (the candidate
// sort and the per-CTA prologue are); the defaults are round 1's.
#ifndef K4_UNROLL
#define K4_UNROLL 2
#endif
#ifndef K4_RW
#define K4_RW 3
#endif
#ifndef K4_MINBLOCKS
#define K4_MINBLOCKS 3
#endif

namespace trx {

// ---------------------------------------------------------------------------------------------
// block-wide helpers
// ---------------------------------------------------------------------------------------------

__device__ __forceinline__ int next_pow2(int v) {
    int p = 2;
    while (p < v) p <<= 1;
    return p;
}

// r-th largest (1-based) key among s[0..n) by 4 x 8-bit MSB-first radix passes.
// hist: 256 x uint32 shared; bc: 4 x uint32 shared.  Returns through shared broadcast.
// cnt_gt = number of elements with key strictly greater than the result.
__device__ void block_radix_select(const float* __restrict__ s, int64_t n, uint32_t r, uint32_t* hist,
                                   uint32_t* bc, uint32_t& kth_key, uint32_t& cnt_gt) {
    uint32_t prefix = 0, mask = 0, remaining = r, gt = 0;
    for (int shift = 24; shift >= 0; shift -= 8) {
        for (int i = threadIdx.x; i < 256; i += blockDim.x) hist[i] = 0;
        __syncthreads();
        for (int64_t i = threadIdx.x; i < n; i += blockDim.x) {
            uint32_t key = f2key(s[i]);
            bool in = (key & mask) == prefix;
            uint32_t bin = (key >> shift) & 255u;
            // warp-aggregate same-bin increments (scores cluster in a handful of exponent bins)
            uint32_t act = __ballot_sync(__activemask(), in);
            if (in) {
                uint32_t peers = __match_any_sync(act, bin);
                if ((int)(threadIdx.x & 31) == __ffs(peers) - 1) atomicAdd(&hist[bin], (uint32_t)__popc(peers));
            }
        }
        __syncthreads();
        if (threadIdx.x < 32) {
            // warp-parallel scan from the top bin down: lane l owns bins [8l, 8l+8)
            const int ln = threadIdx.x;
            uint32_t c[8], local = 0;
#pragma unroll
            for (int j = 0; j < 8; j++) { c[j] = hist[ln * 8 + j]; local += c[j]; }
            uint32_t incl = local;   // inclusive suffix sum over lanes
#pragma unroll
            for (int o = 1; o < 32; o <<= 1) {
                uint32_t t = __shfl_down_sync(0xffffffffu, incl, o);
                if (ln + o < 32) incl += t;
            }
            const uint32_t above = incl - local;   // elements in the bins of higher lanes
            const bool crossing = above < remaining && remaining <= above + local;
            const bool underflow = ln == 0 && remaining > above + local;   // fewer than `remaining` elements: bin 0
            if (crossing || underflow) {
                uint32_t cum = above;
                int b = 7;
                for (; b > 0; b--) {
                    if (cum + c[b] >= remaining) break;
                    cum += c[b];
                }
                bc[0] = (uint32_t)(ln * 8 + b);
                bc[1] = cum;
            }
        }
        __syncthreads();
        prefix |= bc[0] << shift;
        mask |= 255u << shift;
        remaining -= bc[1];
        gt += bc[1];
        __syncthreads();
    }
    kth_key = prefix;
    cnt_gt = gt;
}

__device__ __forceinline__ uint32_t block_sum_u32(uint32_t v, uint32_t* scratch /*>=33*/) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    __syncthreads();
    if ((threadIdx.x & 31) == 0) scratch[threadIdx.x >> 5] = v;
    __syncthreads();
    if (threadIdx.x < 32) {
        uint32_t w = threadIdx.x < (blockDim.x >> 5) ? scratch[threadIdx.x] : 0;
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) w += __shfl_xor_sync(0xffffffffu, w, o);
        if (threadIdx.x == 0) scratch[32] = w;
    }
    __syncthreads();
    return scratch[32];
}

// ---------------------------------------------------------------------------------------------
// row_kth: threshold = r-th largest of a materialised score row (K3 sample pass)
// ---------------------------------------------------------------------------------------------
constexpr int kRowKthSmem = 8192;
__global__ void __launch_bounds__(512) row_kth_kernel(const float* __restrict__ s, int64_t ld, int64_t n, int r,
                                                      float* __restrict__ thr, const float* __restrict__ margin, int ks) {
    __shared__ uint32_t hist[256];
    __shared__ uint32_t bc[4];
    __shared__ uint32_t scratch[33];
    __shared__ float srow[kRowKthSmem];
    const float* row = s + (int64_t)blockIdx.x * ld;
    if (n <= kRowKthSmem) {   // short rows (slot maxima of a small batch): one coalesced read, passes from smem
        for (int64_t i = threadIdx.x; i < n; i += blockDim.x) srow[i] = row[i];
        __syncthreads();
        row = srow;
    }
    uint32_t fin = 0;
    for (int64_t i = threadIdx.x; i < n; i += blockDim.x) fin += row[i] > -INFINITY ? 1u : 0u;
    fin = block_sum_u32(fin, scratch);
    if (fin < (uint32_t)r) {
        if (threadIdx.x == 0) thr[blockIdx.x] = -INFINITY;
        return;
    }
    uint32_t key, gt;
    block_radix_select(row, n, (uint32_t)r, hist, bc, key, gt);
    float t = key2f(key);
    if (margin != nullptr && ks >= 1 && ks < r) {     // see launch_slot_thr: keep the threshold 2.5 eps under the estimated k-th score
        __syncthreads();
        uint32_t key2, gt2;
        block_radix_select(row, n, (uint32_t)ks, hist, bc, key2, gt2);
        t = fminf(t, key2f(key2) - 2.5f * margin[blockIdx.x]);
    }
    if (threadIdx.x == 0) thr[blockIdx.x] = t;
}

int launch_row_kth(const float* s, int64_t ld, int64_t n, int64_t nq, int r, float* thr, cudaStream_t st,
                   const float* margin, int ks) {
    if (nq <= 0) return TRX_OK;
    row_kth_kernel<<<(unsigned)nq, 512, 0, st>>>(s, ld, n, r, thr, margin, ks);
    count_launch();
    TRX_CUDA(cudaGetLastError());
    return TRX_OK;
}

// ---------------------------------------------------------------------------------------------
// exact top-k over materialised scores (certified fallback / forced exact path)
// ---------------------------------------------------------------------------------------------
constexpr int kExactMaxK = 2048;

__global__ void __launch_bounds__(1024) exact_topk_kernel(const float* __restrict__ s, int64_t ld, int64_t n, int k,
                                                          bool negate_out, int64_t id_offset,
                                                          const int32_t* __restrict__ qmap, float* __restrict__ D,
                                                          int64_t* __restrict__ I) {
    __shared__ uint32_t hist[256];
    __shared__ uint32_t bc[4];
    __shared__ uint32_t scratch[33];
    __shared__ uint32_t warp_eq[32];
    __shared__ uint32_t out_ctr;
    __shared__ uint64_t okeys[kExactMaxK];

    const float* row = s + (int64_t)blockIdx.x * ld;
    const int64_t orow = qmap ? (int64_t)qmap[blockIdx.x] : (int64_t)blockIdx.x;
    float* Dq = D + orow * k;
    int64_t* Iq = I + orow * k;
    const float fill = negate_out ? FLT_MAX : -FLT_MAX;

    uint32_t elig = 0;
    for (int64_t i = threadIdx.x; i < n; i += blockDim.x) elig += row[i] > -INFINITY ? 1u : 0u;
    elig = block_sum_u32(elig, scratch);
    const uint32_t r = elig < (uint32_t)k ? elig : (uint32_t)k;
    const int P = next_pow2(k);
    for (int i = threadIdx.x; i < P; i += blockDim.x) okeys[i] = KEY_SENTINEL;
    if (threadIdx.x == 0) out_ctr = 0;
    __syncthreads();
    if (r > 0) {
        uint32_t kth, gt;
        block_radix_select(row, n, r, hist, bc, kth, gt);
        const uint32_t need_eq = r - gt;
        uint32_t eq_base = 0;
        const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
        for (int64_t t0 = 0; t0 < n; t0 += blockDim.x) {
            int64_t i = t0 + threadIdx.x;
            float v = i < n ? row[i] : -INFINITY;
            uint32_t key = f2key(v);
            bool is_gt = i < n && key > kth;
            bool is_eq = i < n && key == kth && v > -INFINITY;
            if (is_gt) {
                uint32_t pos = atomicAdd(&out_ctr, 1u);
                okeys[pos] = pack_key(v, (uint32_t)i);
            }
            if (eq_base < need_eq) {  // uniform across the block
                uint32_t b = __ballot_sync(0xffffffffu, is_eq);
                if (lane == 0) warp_eq[wid] = __popc(b);
                __syncthreads();
                uint32_t before = 0, total = 0;
                for (int w = 0; w < (int)(blockDim.x >> 5); w++) {
                    uint32_t c = warp_eq[w];
                    if (w < wid) before += c;
                    total += c;
                }
                uint32_t rank = eq_base + before + __popc(b & ((1u << lane) - 1u));
                if (is_eq && rank < need_eq) {
                    uint32_t pos = atomicAdd(&out_ctr, 1u);
                    okeys[pos] = pack_key(v, (uint32_t)i);
                }
                eq_base += total;
                __syncthreads();
            }
        }
    }
    bitonic_sort_u64(okeys, P);
    for (int j = threadIdx.x; j < k; j += blockDim.x) {
        uint64_t key = okeys[j];
        if (key == KEY_SENTINEL || j >= (int)r) { Dq[j] = fill; Iq[j] = -1; }
        else {
            float sc = key_score(key);
            Dq[j] = negate_out ? -sc : sc;
            Iq[j] = (int64_t)key_id(key) + id_offset;
        }
    }
}

int launch_exact_topk(const float* s, int64_t ld, int64_t n, int64_t nq, int k, bool negate_out,
                      int64_t id_offset, const int32_t* qmap, float* D, int64_t* I, cudaStream_t st) {
    if (nq <= 0) return TRX_OK;
    if (k > kExactMaxK) { set_error("k=%d exceeds the supported maximum %d", k, kExactMaxK); return TRX_EINVAL; }
    exact_topk_kernel<<<(unsigned)nq, 1024, 0, st>>>(s, ld, n, k, negate_out, id_offset, qmap, D, I);
    count_launch();
    TRX_CUDA(cudaGetLastError());
    return TRX_OK;
}

// ---------------------------------------------------------------------------------------------
// K4: candidates -> exact fp32 rescore -> certificate -> final top-k
// ---------------------------------------------------------------------------------------------
// One CTA per query (256 threads; 1024 when the batch is too small to fill the chip with 256-thread
// CTAs, so that a lone query's ~256 row gathers are spread over 32 warps).  Candidates come from a bf16 prefilter whose invariant is:
// every row NOT in the list has prefilter score <= thr[q].  Let eps bound |prefilter - exact|.
// After rescoring the best m candidates (by prefilter score) exactly, the exact top-k of those m
// is the exact top-k of the whole corpus if  exact_k > max(next prefilter score, thr) + eps.
// Otherwise rescore more; if the list is exhausted the query goes to the exact scan.
template <int METRIC>
__device__ __forceinline__ float exact_score_warp(const float* __restrict__ xr, const float* __restrict__ sq, int d,
                                                  int lane) {
    float acc = 0.f;
    if ((d & 3) == 0) {
        const float4* x4 = reinterpret_cast<const float4*>(xr);
        const float4* q4 = reinterpret_cast<const float4*>(sq);
        for (int c = lane; c < (d >> 2); c += 32) {
            float4 x = __ldg(x4 + c);
            float4 q = q4[c];
            if (METRIC == TRX_METRIC_INNER_PRODUCT) {
                acc = fmaf(x.x, q.x, acc); acc = fmaf(x.y, q.y, acc);
                acc = fmaf(x.z, q.z, acc); acc = fmaf(x.w, q.w, acc);
            } else {
                float t0 = q.x - x.x, t1 = q.y - x.y, t2 = q.z - x.z, t3 = q.w - x.w;
                acc = fmaf(t0, t0, acc); acc = fmaf(t1, t1, acc);
                acc = fmaf(t2, t2, acc); acc = fmaf(t3, t3, acc);
            }
        }
    } else {
        for (int c = lane; c < d; c += 32) {
            float x = __ldg(xr + c), q = sq[c];
            if (METRIC == TRX_METRIC_INNER_PRODUCT) acc = fmaf(x, q, acc);
            else { float t = q - x; acc = fmaf(t, t, acc); }
        }
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) acc += __shfl_xor_sync(0xffffffffu, acc, o);
    return METRIC == TRX_METRIC_INNER_PRODUCT ? acc : -acc;
}

// RW rows at once: RW times the gather bytes in flight per warp (the rescore is bound by the latency of ~3 KB row
// gathers); every row keeps exactly the summation order of exact_score_warp, so scores do not depend on the grouping.
template <int METRIC, int RW>
__device__ __forceinline__ void exact_score_warp_n(const float* const (&xr)[RW], const float* __restrict__ sq, int d,
                                                   int lane, float (&e)[RW]) {
    if ((d & 3) == 0) {
        float acc[RW];
#pragma unroll
        for (int r = 0; r < RW; r++) acc[r] = 0.f;
        const float4* q4 = reinterpret_cast<const float4*>(sq);
        constexpr int kGatherUnroll = K4_UNROLL;
#pragma unroll kGatherUnroll
        for (int c = lane; c < (d >> 2); c += 32) {
            float4 x[RW];
#pragma unroll
            for (int r = 0; r < RW; r++) x[r] = __ldg(reinterpret_cast<const float4*>(xr[r]) + c);
            const float4 q = q4[c];
#pragma unroll
            for (int r = 0; r < RW; r++) {
                if (METRIC == TRX_METRIC_INNER_PRODUCT) {
                    acc[r] = fmaf(x[r].x, q.x, acc[r]); acc[r] = fmaf(x[r].y, q.y, acc[r]);
                    acc[r] = fmaf(x[r].z, q.z, acc[r]); acc[r] = fmaf(x[r].w, q.w, acc[r]);
                } else {
                    float t0 = q.x - x[r].x, t1 = q.y - x[r].y, t2 = q.z - x[r].z, t3 = q.w - x[r].w;
                    acc[r] = fmaf(t0, t0, acc[r]); acc[r] = fmaf(t1, t1, acc[r]);
                    acc[r] = fmaf(t2, t2, acc[r]); acc[r] = fmaf(t3, t3, acc[r]);
                }
            }
        }
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) {
#pragma unroll
            for (int r = 0; r < RW; r++) acc[r] += __shfl_xor_sync(0xffffffffu, acc[r], o);
        }
#pragma unroll
        for (int r = 0; r < RW; r++) e[r] = METRIC == TRX_METRIC_INNER_PRODUCT ? acc[r] : -acc[r];
    } else {
#pragma unroll
        for (int r = 0; r < RW; r++) e[r] = exact_score_warp<METRIC>(xr[r], sq, d, lane);
    }
}

template <int METRIC, int NT>
__global__ void __launch_bounds__(NT, NT == 256 ? K4_MINBLOCKS : 1) k4_rescore_kernel(RescoreArgs a) {
    extern __shared__ __align__(16) unsigned char smem_raw[];
    // layout: keys[cap] u64 | ekeys[cap] u64 | q[dpad] f32 | dedup only: sgrp[cap] i32 | oidx[k] i32 | slead[cap] u8
    uint64_t* keys = reinterpret_cast<uint64_t*>(smem_raw);
    uint64_t* ekeys = keys + a.cap;
    float* sq = reinterpret_cast<float*>(ekeys + a.cap);
    int32_t* sgrp = reinterpret_cast<int32_t*>(sq + ((a.d + 3) & ~3));
    int32_t* oidx = sgrp + a.cap;
    uint8_t* slead = reinterpret_cast<uint8_t*>(oidx + a.k);
    __shared__ uint32_t s_valid;
    __shared__ float s_qn2;
    __shared__ int s_cnt[1024];
    __shared__ int s_kpos, s_nlead;

    const int64_t q = blockIdx.x;
    const int64_t oq = a.qmap ? (int64_t)a.qmap[q] : q;   // caller-visible query index
    const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5, nwarp = blockDim.x >> 5;
    const int k = a.k;
    float* Dq = a.D + oq * k;
    int64_t* Iq = a.I + oq * k;
    const float fill = METRIC == TRX_METRIC_L2 ? FLT_MAX : -FLT_MAX;

    const uint32_t cnt_raw = a.cand_cnt[q];
    if (cnt_raw > (uint32_t)a.cap) {  // list overflowed: invariant lost
        if (threadIdx.x == 0) {
            uint32_t slot = atomicAdd(a.fb_count, 1u);
            a.fb_list[slot] = (int32_t)oq;
            a.fb_thr[slot] = -INFINITY;
            atomicAdd((unsigned long long*)&a.counters[2], 1ull);
            atomicAdd((unsigned long long*)&a.counters[3], (unsigned long long)cnt_raw);
        }
        return;
    }
    const int n_c = (int)cnt_raw;
    const int32_t ex = (a.excl != nullptr && a.groups != nullptr) ? a.excl[q] : -1;
    if (threadIdx.x == 0) { s_valid = 0; s_qn2 = 0.f; }
    float qn2_part = 0.f;
    for (int c = threadIdx.x; c < a.d; c += blockDim.x) {
        float v = a.q32[q * (int64_t)a.d + c];
        sq[c] = v;
        qn2_part = fmaf(v, v, qn2_part);
    }
    const int P = next_pow2(n_c);
    __syncthreads();
    uint32_t my_valid = 0;
    if (a.presorted) {   // launch_bounds already applied the masks and sorted the list
        for (int i = threadIdx.x; i < n_c; i += blockDim.x) {
            Cand c = a.cand[q * (int64_t)a.cap + i];
            keys[i] = pack_key(c.score, (uint32_t)c.row);
        }
        if (threadIdx.x == 0) s_valid = (uint32_t)n_c;
    } else {
        for (int i = threadIdx.x; i < P; i += blockDim.x) {
            uint64_t key = KEY_SENTINEL;
            if (i < n_c) {
                Cand c = a.cand[q * (int64_t)a.cap + i];
                bool ok = !(ex >= 0 && __ldg(a.groups + c.row) == ex);
                if (a.attr != nullptr && __ldg(a.attr + c.row) >= a.attr_below) ok = false;
                if (ok) { key = pack_key(c.score, (uint32_t)c.row); my_valid++; }
            }
            keys[i] = key;
        }
        if (my_valid) atomicAdd(&s_valid, my_valid);
    }
    if (METRIC == TRX_METRIC_L2) {
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) qn2_part += __shfl_xor_sync(0xffffffffu, qn2_part, o);
        if (lane == 0) atomicAdd(&s_qn2, qn2_part);
    }
    if (a.presorted) __syncthreads();
    else bitonic_sort_u64(keys, P);  // (prefilter score desc, row asc); masked / padding last
    int n_valid = (int)s_valid;
    const float qn2 = s_qn2;
    const float thr = a.thr[q];
    const float eps = a.eps[q];
    bool complete = !(thr > -INFINITY);  // every eligible row is in the list

    // keys[0 .. n_valid) are sorted by prefilter score, best first: number of them scoring at least `cut`
    auto count_at_least = [&](float cut) {
        int lo = 0, hi = n_valid;
        while (lo < hi) {
            const int mid = (lo + hi) >> 1;
            if (key_score(keys[mid]) >= cut) lo = mid + 1; else hi = mid;
        }
        return lo;
    };
    int m_done = 0;
    int s_have = 0, s_pos = 0;       // results available / position of the k-th one after the last round
    // First round: the certificate needs  s_k > ps[m] + eps  (ps = prefilter scores in order, s_k = exact k-th score).
    // s_k is not known yet but sits within the ACTUAL bf16 error (a small fraction of the rigorous eps) of ps[k-1], so
    // rescoring every candidate down to  ps[k-1] - 1.15 eps  certifies almost always -- ~150 rows on Gaussian data,
    // ~450 on clustered unit-norm data (eps ~ one in-cluster sigma), instead of a fixed count that is too many for
    // the one and too few for the other.  Distinct-groups mode: the k-th LEADER is what counts; start from 2k + 56.
    int m;
    if (a.dedup || n_valid <= k) m = (2 * k + 56 + 7) & ~7;
    else m = (count_at_least(key_score(keys[k - 1]) - 1.15f * eps) + 7) & ~7;
    if (m < k + 8) m = k + 8;
    // Row-sharded search: no row scoring under floor[q] can be in the GLOBAL top-k (the shards together hold k rows
    // that beat it, rigorously), and rows the prefilter did not list score <= thr.  If the floor is above thr the rows
    // down to the floor are all this shard can contribute: rescore exactly those -- usually far fewer than the local
    // top-k would need -- and the list is complete.
    if (a.floor != nullptr && !a.dedup) {
        const float fl = a.floor[q];
        if (fl > thr) {
            const int n_eff = count_at_least(fl);
            if (n_eff <= m) { n_valid = n_eff; complete = true; }
        }
    }
    if (m > n_valid || complete) m = n_valid;
    bool certified = false;
    for (;;) {
        constexpr int RW = K4_RW;    // rows a warp gathers at once (measured: 2 -> 710 us, 3 -> 678 us, 4 -> 690 us at C2)
        for (int i = m_done + RW * wid; i < m; i += RW * nwarp) {
            uint32_t rows[RW];
            const float* xr[RW];
            float e[RW];
#pragma unroll
            for (int r = 0; r < RW; r++) {
                rows[r] = key_id(keys[i + r < m ? i + r : i]);
                xr[r] = a.x32 + (int64_t)rows[r] * a.d;
            }
            exact_score_warp_n<METRIC, RW>(xr, sq, a.d, lane, e);
            if (lane == 0) {
#pragma unroll
                for (int r = 0; r < RW; r++)
                    if (i + r < m) ekeys[i + r] = pack_key(e[r], rows[r]);
            }
        }
        const int P2 = next_pow2(m);
        for (int i = m + threadIdx.x; i < P2; i += blockDim.x) ekeys[i] = KEY_SENTINEL;
        bitonic_sort_u64(ekeys, P2);  // includes the leading barrier
        // distinct-groups mode: only the best row of each group counts ("leaders", in exact-score order)
        int n_have = m, kpos = k - 1;
        if (a.dedup) {
            for (int i = threadIdx.x; i < m; i += blockDim.x) sgrp[i] = __ldg(a.groups + key_id(ekeys[i]));
            __syncthreads();
            const int C = (m + (int)blockDim.x - 1) / (int)blockDim.x;
            const int lo = min(m, (int)threadIdx.x * C), hi = min(m, lo + C);
            int cnt = 0;
            for (int i = lo; i < hi; i++) {
                const int g = sgrp[i];
                bool lead = true;
                for (int j = 0; j < i; j++)
                    if (sgrp[j] == g) { lead = false; break; }
                slead[i] = lead ? 1 : 0;
                cnt += lead ? 1 : 0;
            }
            s_cnt[threadIdx.x] = cnt;
            if (threadIdx.x == 0) s_kpos = -1;
            __syncthreads();
            for (int off = 1; off < (int)blockDim.x; off <<= 1) {   // inclusive scan of the per-thread leader counts
                int v = threadIdx.x >= (unsigned)off ? s_cnt[threadIdx.x - off] : 0;
                __syncthreads();
                s_cnt[threadIdx.x] += v;
                __syncthreads();
            }
            int pos = s_cnt[threadIdx.x] - cnt;     // leaders before this thread's chunk
            for (int i = lo; i < hi; i++) {
                if (slead[i]) {
                    if (pos < k) oidx[pos] = i;
                    if (pos == k - 1) s_kpos = i;
                    pos++;
                }
            }
            if (threadIdx.x == blockDim.x - 1) s_nlead = s_cnt[threadIdx.x];
            __syncthreads();
            n_have = s_nlead; kpos = s_kpos;
        }
        float sk = 0.f;                  // k-th best exact score so far, in the prefilter's domain
        if (m == n_valid && complete) certified = true;
        else if (n_have >= k) {
            sk = key_score(ekeys[kpos]);
            if (METRIC == TRX_METRIC_L2) sk += qn2;  // prefilter domain: |q|^2 - dist
            float bound = m < n_valid ? key_score(keys[m]) : thr;
            certified = sk > bound + eps;
        }
        s_have = n_have; s_pos = kpos;   // (thread-uniform values kept for the epilogue)
        if (certified || m == n_valid) break;
        m_done = m;
        // Next round.  With k results in hand the requirement is known exactly: every candidate down to  s_k - eps
        // (s_k can only grow, so this round certifies unless the list ends first).  Without: twice as many.
        int m_next = 2 * m;
        if (n_have >= k) m_next = max((count_at_least(sk - eps) + 7) & ~7, m + 8);
        m = m_next < n_valid ? m_next : n_valid;
        __syncthreads();
    }
    if (threadIdx.x == 0) {
        atomicAdd((unsigned long long*)&a.counters[0], (unsigned long long)m);
        atomicAdd((unsigned long long*)&a.counters[3], (unsigned long long)n_c);
    }
    if (!certified) {
        if (threadIdx.x == 0) {
            uint32_t slot = atomicAdd(a.fb_count, 1u);
            a.fb_list[slot] = (int32_t)oq;
            // every true top-k row scores at least the k-th best exact score seen so far
            a.fb_thr[slot] = s_have >= k ? key_score(ekeys[s_pos]) - a.eps_acc[q] : -INFINITY;
            atomicAdd((unsigned long long*)&a.counters[1], 1ull);
        }
        return;
    }
    for (int j = threadIdx.x; j < k; j += blockDim.x) {
        if (j < s_have) {
            uint64_t key = ekeys[a.dedup ? oidx[j] : j];
            float sc = key_score(key);
            Dq[j] = METRIC == TRX_METRIC_L2 ? -sc : sc;
            Iq[j] = (int64_t)key_id(key) + a.id_offset;
        } else { Dq[j] = fill; Iq[j] = -1; }
    }
}

// ---------------------------------------------------------------------------------------------
// first half of a two-phase (row-sharded) search: filter + sort the candidate lists, export the best scores
// ---------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256) k4_bounds_kernel(BoundsArgs a) {
    extern __shared__ __align__(16) unsigned char smem_raw[];
    uint64_t* keys = reinterpret_cast<uint64_t*>(smem_raw);
    __shared__ uint32_t s_valid;
    const int64_t q = blockIdx.x;
    float* out = a.payload + q * (a.nb + 1);
    const uint32_t cnt_raw = a.cand_cnt[q];
    if (threadIdx.x == 0) out[a.nb] = a.eps[q];
    if (cnt_raw > (uint32_t)a.cap) {   // overflowed list: no bound from this shard; the second half sends it to the scan
        for (int j = threadIdx.x; j < a.nb; j += blockDim.x) out[j] = -INFINITY;
        return;
    }
    const int n_c = (int)cnt_raw;
    const int32_t ex = (a.excl != nullptr && a.groups != nullptr) ? a.excl[q] : -1;
    if (threadIdx.x == 0) s_valid = 0;
    const int P = next_pow2(n_c);
    __syncthreads();
    uint32_t my_valid = 0;
    for (int i = threadIdx.x; i < P; i += blockDim.x) {
        uint64_t key = KEY_SENTINEL;
        if (i < n_c) {
            Cand c = a.cand[q * (int64_t)a.cap + i];
            bool ok = !(ex >= 0 && __ldg(a.groups + c.row) == ex);
            if (a.attr != nullptr && __ldg(a.attr + c.row) >= a.attr_below) ok = false;
            if (ok) { key = pack_key(c.score, (uint32_t)c.row); my_valid++; }
        }
        keys[i] = key;
    }
    if (my_valid) atomicAdd(&s_valid, my_valid);
    bitonic_sort_u64(keys, P);
    const int n_valid = (int)s_valid;
    for (int i = threadIdx.x; i < n_valid; i += blockDim.x) {
        Cand c; c.score = key_score(keys[i]); c.row = (int32_t)key_id(keys[i]);
        a.cand[q * (int64_t)a.cap + i] = c;
    }
    for (int j = threadIdx.x; j < a.nb; j += blockDim.x) out[j] = j < n_valid ? key_score(keys[j]) : -INFINITY;
    if (threadIdx.x == 0) a.cand_cnt[q] = (uint32_t)n_valid;
}

int launch_bounds(const BoundsArgs& a, cudaStream_t st) {
    if (a.nq <= 0) return TRX_OK;
    const size_t smem = (size_t)a.cap * 8;
    if (smem > kK4MaxSmem) { set_error("k4 bounds: cap=%d exceeds shared memory", a.cap); return TRX_EINVAL; }
    TRX_CUDA(cudaFuncSetAttribute(k4_bounds_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    k4_bounds_kernel<<<(unsigned)a.nq, 256, smem, st>>>(a);
    count_launch();
    TRX_CUDA(cudaGetLastError());
    return TRX_OK;
}

int launch_rescore(const RescoreArgs& a, cudaStream_t st) {
    if (a.nq <= 0) return TRX_OK;
    if (a.dedup && a.groups == nullptr) { set_error("k4: distinct-groups mode needs groups"); return TRX_EINVAL; }
    const size_t smem = k4_smem_bytes(a.cap, a.d, a.k, a.dedup != 0);
    if (smem > kK4MaxSmem) { set_error("k4: cap=%d d=%d exceed shared memory", a.cap, a.d); return TRX_EINVAL; }
    const bool wide = a.nq <= 296;   // few queries: 1024-thread CTAs so that one query's gathers use 32 warps
#define TRX_K4(M, NT)                                                                                        \
    do {                                                                                                     \
        auto kern = k4_rescore_kernel<M, NT>;                                                                \
        TRX_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));        \
        kern<<<(unsigned)a.nq, NT, smem, st>>>(a);                                                           \
    } while (0)
    if (a.metric == TRX_METRIC_L2) { if (wide) TRX_K4(TRX_METRIC_L2, 1024); else TRX_K4(TRX_METRIC_L2, 256); }
    else { if (wide) TRX_K4(TRX_METRIC_INNER_PRODUCT, 1024); else TRX_K4(TRX_METRIC_INNER_PRODUCT, 256); }
#undef TRX_K4
    count_launch();
    TRX_CUDA(cudaGetLastError());
    return TRX_OK;
}

// ---------------------------------------------------------------------------------------------
// distinct-groups filter for the exact path: rows of kx sorted results -> the first k group leaders
// ---------------------------------------------------------------------------------------------
// Runs in rounds when k * (largest group) exceeds the widest exact scan (kx): a round appends the leaders found in
// its kx best rows behind the nfound[q] leaders of earlier rounds; mask_seen_groups_kernel then removes every row
// of the groups the round saw from the score rows, so the next round's rows all belong to new groups and its
// leaders continue the (score desc, id asc) order.  nfound[q] == k marks a finished query.
__global__ void __launch_bounds__(128) dedup_rows_kernel(const float* __restrict__ Dx, const int64_t* __restrict__ Ix,
                                                         int kx, const int32_t* __restrict__ groups, int64_t id_offset,
                                                         int k, float fill, const int32_t* __restrict__ qmap,
                                                         float* __restrict__ D, int64_t* __restrict__ I,
                                                         int32_t* __restrict__ nfound, uint32_t* __restrict__ unfinished) {
    extern __shared__ int32_t sg[];   // [kx] group of each result (INT32_MIN for padding)
    __shared__ int s_out, s_valid;
    const int64_t q = blockIdx.x;
    const int64_t oq = qmap ? (int64_t)qmap[q] : q;
    const int have = nfound ? nfound[q] : 0;
    if (have >= k) return;            // finished in an earlier round
    const float* Dr = Dx + q * kx;
    const int64_t* Ir = Ix + q * kx;
    if (threadIdx.x == 0) { s_out = 0; s_valid = 0; }
    __syncthreads();
    int nv = 0;
    for (int i = threadIdx.x; i < kx; i += blockDim.x) {
        const int64_t id = Ir[i];
        sg[i] = id >= 0 ? __ldg(groups + (id - id_offset)) : INT32_MIN;
        nv += id >= 0 ? 1 : 0;
    }
    if (nv) atomicAdd(&s_valid, nv);
    __syncthreads();
    // leaders are taken in order by one warp: ballot over 32 results at a time keeps the order stable
    if (threadIdx.x < 32) {
        int out = have;
        for (int i0 = 0; i0 < kx && out < k; i0 += 32) {
            const int i = i0 + threadIdx.x;
            bool lead = false;
            if (i < kx && Ir[i] >= 0) {
                lead = true;
                const int g = sg[i];
                for (int j = 0; j < i; j++)
                    if (sg[j] == g && Ir[j] >= 0) { lead = false; break; }
            }
            const uint32_t b = __ballot_sync(0xffffffffu, lead);
            const int pos = out + __popc(b & ((1u << threadIdx.x) - 1u));
            if (lead && pos < k) { D[oq * k + pos] = Dr[i]; I[oq * k + pos] = Ir[i]; }
            out += __popc(b);
        }
        if (threadIdx.x == 0) s_out = out < k ? out : k;
    }
    __syncthreads();
    const int out = s_out;
    // a full round list may hide further groups behind it: another round, unless k leaders are already found
    const bool more = out < k && s_valid == kx && nfound != nullptr;
    if (more) {
        if (threadIdx.x == 0) { nfound[q] = out; atomicAdd(unfinished, 1u); }
        return;
    }
    for (int j = out + threadIdx.x; j < k; j += blockDim.x) { D[oq * k + j] = fill; I[oq * k + j] = -1; }
    if (threadIdx.x == 0 && nfound) nfound[q] = k;
}

// Unfinished queries: score rows of every group present in the round list Ix[q] become -inf (ineligible).
__global__ void __launch_bounds__(256) mask_seen_groups_kernel(float* __restrict__ scores, int64_t ld, int64_t n,
                                                               const int64_t* __restrict__ Ix, int kx,
                                                               const int32_t* __restrict__ groups, int64_t id_offset,
                                                               const int32_t* __restrict__ nfound, int k) {
    extern __shared__ int32_t sgm[];  // [P] sorted groups of the round list
    const int64_t q = blockIdx.y;
    if (nfound[q] >= k) return;
    int P = 2;
    while (P < kx) P <<= 1;
    for (int i = threadIdx.x; i < P; i += blockDim.x) {
        const int64_t id = i < kx ? Ix[q * kx + i] : -1;
        sgm[i] = id >= 0 ? __ldg(groups + (id - id_offset)) : INT32_MAX;
    }
    for (int size = 2; size <= P; size <<= 1) {
        for (int stride = size >> 1; stride > 0; stride >>= 1) {
            __syncthreads();
            for (int t = threadIdx.x; t < (P >> 1); t += blockDim.x) {
                const int i = 2 * t - (t & (stride - 1)), j = i + stride;
                const int32_t a = sgm[i], b = sgm[j];
                if ((a > b) == ((i & size) == 0)) { sgm[i] = b; sgm[j] = a; }
            }
        }
    }
    __syncthreads();
    float* row = scores + q * ld;
    for (int64_t r = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; r < n; r += (int64_t)gridDim.x * blockDim.x) {
        if (!(row[r] > -INFINITY)) continue;
        const int32_t g = __ldg(groups + r);
        int lo = 0, hi = kx;          // an unfinished query's round list is full: the first kx sorted entries are
        while (lo < hi) {             // exactly its groups, the padding sorts behind them
            const int mid = (lo + hi) >> 1;
            if (sgm[mid] < g) lo = mid + 1; else hi = mid;
        }
        if (lo < kx && sgm[lo] == g) row[r] = -INFINITY;
    }
}

int launch_dedup_rows(const float* Dx, const int64_t* Ix, int kx, const int32_t* groups, int64_t id_offset, int k,
                      bool l2, const int32_t* qmap, int64_t nq, float* D, int64_t* I, int32_t* nfound,
                      uint32_t* unfinished, cudaStream_t st) {
    if (nq <= 0) return TRX_OK;
    dedup_rows_kernel<<<(unsigned)nq, 128, (size_t)kx * 4, st>>>(Dx, Ix, kx, groups, id_offset, k, l2 ? FLT_MAX : -FLT_MAX,
                                                                qmap, D, I, nfound, unfinished);
    count_launch();
    TRX_CUDA(cudaGetLastError());
    return TRX_OK;
}

int launch_mask_seen_groups(float* scores, int64_t ld, int64_t n, const int64_t* Ix, int kx, const int32_t* groups,
                            int64_t id_offset, const int32_t* nfound, int k, int64_t nq, cudaStream_t st) {
    if (nq <= 0) return TRX_OK;
    int P = 2;
    while (P < kx) P <<= 1;
    const unsigned bx = (unsigned)std::min<int64_t>(296, (n + 255) / 256);
    mask_seen_groups_kernel<<<dim3(bx, (unsigned)nq), 256, (size_t)P * 4, st>>>(scores, ld, n, Ix, kx, groups, id_offset,
                                                                                  nfound, k);
    count_launch();
    TRX_CUDA(cudaGetLastError());
    return TRX_OK;
}

// ---------------------------------------------------------------------------------------------
// K5: k-way merge of G per-shard sorted lists (after the NCCL all-gather)
// ---------------------------------------------------------------------------------------------
// Position index g*k+j is the tie-break: shards hold ascending, disjoint id ranges and every
// list is already (score, id)-ordered, so position order == id order among equal scores.
__global__ void __launch_bounds__(256) k5_merge_kernel(int metric, const float* __restrict__ Dg,
                                                       const int64_t* __restrict__ Ig, int G, int64_t nq, int k,
                                                       float* __restrict__ D, int64_t* __restrict__ I) {
    extern __shared__ __align__(16) unsigned char smem_raw[];
    uint64_t* keys = reinterpret_cast<uint64_t*>(smem_raw);
    const int64_t q = blockIdx.x;
    const int total = G * k;
    const int P = next_pow2(total);
    for (int i = threadIdx.x; i < P; i += blockDim.x) {
        uint64_t key = KEY_SENTINEL;
        if (i < total) {
            int g = i / k, j = i - g * k;
            int64_t src = ((int64_t)g * nq + q) * k + j;
            if (Ig[src] >= 0) {
                float v = Dg[src];
                key = pack_key(metric == TRX_METRIC_L2 ? -v : v, (uint32_t)i);
            }
        }
        keys[i] = key;
    }
    bitonic_sort_u64(keys, P);
    const float fill = metric == TRX_METRIC_L2 ? FLT_MAX : -FLT_MAX;
    for (int j = threadIdx.x; j < k; j += blockDim.x) {
        uint64_t key = keys[j];
        if (key == KEY_SENTINEL) { D[q * k + j] = fill; I[q * k + j] = -1; }
        else {
            int pos = (int)key_id(key);
            int g = pos / k, jj = pos - g * k;
            int64_t src = ((int64_t)g * nq + q) * k + jj;
            D[q * k + j] = Dg[src];
            I[q * k + j] = Ig[src];
        }
    }
}

int launch_merge(int metric, const float* Dg, const int64_t* Ig, int G, int64_t nq, int k, float* D, int64_t* I,
                 cudaStream_t st) {
    if (nq <= 0) return TRX_OK;
    int64_t total = (int64_t)G * k;
    int P = 2;
    while (P < total) P <<= 1;
    size_t smem = (size_t)P * 8;
    if (smem > 200 * 1024) { set_error("merge: G*k=%lld too large", (long long)total); return TRX_EINVAL; }
    TRX_CUDA(cudaFuncSetAttribute(k5_merge_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    k5_merge_kernel<<<(unsigned)nq, 256, smem, st>>>(metric, Dg, Ig, G, nq, k, D, I);
    count_launch();
    TRX_CUDA(cudaGetLastError());
    return TRX_OK;
}

// ---------------------------------------------------------------------------------------------
// slot_thr: threshold from the K2 SLOTMAX epilogue (r-th largest of the S*32 slot maxima)
// ---------------------------------------------------------------------------------------------
// The slot maxima are a subset of the sample scores, so their r-th largest is <= the sample's
// r-th largest: a slightly permissive threshold, never a wrong one (the certificate in K4 is
// what guarantees exactness; the threshold only sizes the candidate list).
__global__ void __launch_bounds__(256) slot_thr_kernel(const float* __restrict__ slots, int64_t nq, int S, int r,
                                                       float* __restrict__ thr, const float* __restrict__ margin, int ks) {
    const int lane = threadIdx.x & 31;
    const int64_t q = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    if (q >= nq) return;
    float v[8];
#pragma unroll
    for (int i = 0; i < 8; i++) v[i] = i < S ? slots[(q * S + i) * 32 + lane] : -INFINITY;
    float best = -INFINITY, best_ks = INFINITY;
    for (int it = 0; it < r; it++) {
        float m = v[0];
#pragma unroll
        for (int i = 1; i < 8; i++) m = fmaxf(m, v[i]);
        float wm = m;
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) wm = fmaxf(wm, __shfl_xor_sync(0xffffffffu, wm, o));
        best = wm;
        if (it == ks - 1) best_ks = wm;
        uint32_t owners = __ballot_sync(0xffffffffu, m == wm);
        if (lane == __ffs(owners) - 1) {  // remove one instance
            bool done = false;
#pragma unroll
            for (int i = 0; i < 8; i++)
                if (!done && v[i] == wm) { v[i] = -INFINITY; done = true; }
        }
    }
    if (lane == 0) thr[q] = (margin != nullptr && ks >= 1 && ks < r) ? fminf(best, best_ks - 2.5f * margin[q]) : best;
}

int launch_slot_thr(const float* slots, int64_t nq, int S, int r, float* thr, cudaStream_t st, const float* margin,
                    int ks) {
    if (nq <= 0) return TRX_OK;
    if (r > 32 * S) { set_error("slot_thr: S=%d r=%d unsupported", S, r); return TRX_EINVAL; }
    if (S > 8) return launch_row_kth(slots, (int64_t)S * 32, (int64_t)S * 32, nq, r, thr, st, margin, ks);  // small batches: many slices
    slot_thr_kernel<<<(unsigned)((nq + 7) / 8), 256, 0, st>>>(slots, nq, S, r, thr, margin, ks);
    count_launch();
    TRX_CUDA(cudaGetLastError());
    return TRX_OK;
}

}  // namespace trx
