// k3_stream.cu -- CUDA-core streaming scorer (bandwidth-bound regime).
//
// Replaces the arithmetic of faiss `index.search` (retrieve/retrieve_faiss.py:71) for
//   * the exact path: fp32 corpus, fp32 FMA, direct sum (q-x)^2 for L2 -- this is also the
//     certified fallback every bf16-prefiltered query can fall back to;
//   * tiny batches: one pass over the bf16 corpus with 128-bit loads, candidates above the
//     sampled threshold appended on the fly (no score matrix in HBM).
//
// Mapping.  A warp owns R = 4 consecutive corpus rows at a time; lane l reads the 16-byte chunks
// l, l+32, ... of each of the four rows (ld.global.nc.L1::no_allocate.v4: a warp instruction covers
// 512 contiguous bytes of one row), so every query chunk fetched from shared memory (the QB <= 4
// queries of the pass live there as fp32) is used for four rows -- the kernel stays HBM-bound instead
// of LDS-bound -- and up to 8 independent 16-byte loads per lane are in flight.  Arithmetic is packed
// fp32 (fma.rn.f32x2, two lanes of the dot product per instruction).  The R*QB partial sums of a
// warp are reduced with a transposing butterfly (V/2 + V/4 + ... shuffles instead of 5 V), after
// which lane (idx << s) holds the total of (row idx / QB, query idx % QB) and does its own
// threshold test / store.
#include "common.cuh"

namespace trx {

namespace {

constexpr int R = 4;   // corpus rows a warp scores together

__device__ __forceinline__ uint4 ldg_nc_16(const void* p) {
    uint4 r;
    asm volatile("ld.global.nc.L1::no_allocate.v4.u32 {%0,%1,%2,%3}, [%4];"
                 : "=r"(r.x), "=r"(r.y), "=r"(r.z), "=r"(r.w) : "l"(p));
    return r;
}

// d = a * b + c on two packed fp32 lanes
__device__ __forceinline__ float2 ffma2(float2 a, float2 b, float2 c) {
    unsigned long long ra, rb, rc, rd;
    asm("mov.b64 %0, {%1, %2};" : "=l"(ra) : "f"(a.x), "f"(a.y));
    asm("mov.b64 %0, {%1, %2};" : "=l"(rb) : "f"(b.x), "f"(b.y));
    asm("mov.b64 %0, {%1, %2};" : "=l"(rc) : "f"(c.x), "f"(c.y));
    asm("fma.rn.f32x2 %0, %1, %2, %3;" : "=l"(rd) : "l"(ra), "l"(rb), "l"(rc));
    float2 d;
    asm("mov.b64 {%0, %1}, %2;" : "=f"(d.x), "=f"(d.y) : "l"(rd));
    return d;
}

__device__ __forceinline__ float2 bf16x2_to_f32x2(uint32_t u) {
    return make_float2(__uint_as_float(u << 16), __uint_as_float(u & 0xffff0000u));
}

// Sum each of V (power of two, <= 32) per-lane values over the warp.  On return lane l holds the
// total of value index l >> (5 - log2 V) (lanes sharing an index hold the same total).
template <int V>
__device__ __forceinline__ float warp_multi_reduce(float (&v)[V], int lane) {
    int n = V;
#pragma unroll
    for (int off = 16; off > 0; off >>= 1) {
        if (n > 1) {
            const int half = n >> 1;
            const bool upper = (lane & off) != 0;
#pragma unroll
            for (int i = 0; i < V / 2; i++) {
                if (i < half) {
                    const float send = upper ? v[i] : v[i + half];
                    const float keep = upper ? v[i + half] : v[i];
                    v[i] = keep + __shfl_xor_sync(0xffffffffu, send, off);
                }
            }
            n = half;
        } else {
            v[0] += __shfl_xor_sync(0xffffffffu, v[0], off);
        }
    }
    return v[0];
}

template <int V> struct Log2;
template <> struct Log2<1> { static constexpr int v = 0; };
template <> struct Log2<2> { static constexpr int v = 1; };
template <> struct Log2<4> { static constexpr int v = 2; };
template <> struct Log2<8> { static constexpr int v = 3; };
template <> struct Log2<16> { static constexpr int v = 4; };
template <> struct Log2<32> { static constexpr int v = 5; };

}  // namespace

// BF16: x rows are bf16 with `pitch` elements, reduce over d (multiple of 8) columns, IP formula.
// !BF16: x rows fp32, reduce over d columns, METRIC selects dot or -(sum sq diff).
template <bool BF16, int METRIC, int QB, bool APPEND>
__global__ void __launch_bounds__(256) k3_stream_kernel(StreamArgs a, int64_t q0) {
    extern __shared__ __align__(16) float sq[];  // [QB][d] fp32 queries
    __shared__ int32_t s_ex[QB];
    __shared__ float s_thr[QB];
    const int d = a.d;
    const int nqb = (int)min((int64_t)QB, a.nq - q0);
    for (int i = threadIdx.x; i < QB * d; i += blockDim.x) {
        int q = i / d, c = i - q * d;
        float v = 0.f;
        if (q < nqb) {
            v = BF16 ? __bfloat162float(a.q16[(q0 + q) * a.q_pitch + c]) : a.q32[(q0 + q) * a.q_pitch + c];
        }
        sq[i] = v;
    }
    if (threadIdx.x < QB) {
        const int q = threadIdx.x;
        s_ex[q] = (a.excl != nullptr && a.groups != nullptr && q < nqb) ? a.excl[q0 + q] : -1;
        s_thr[q] = (APPEND && q < nqb) ? a.thr[q0 + q] : 0.f;
    }
    __syncthreads();

    const int lane = threadIdx.x & 31;
    const int64_t warp0 = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    const int64_t nwarps = ((int64_t)gridDim.x * blockDim.x) >> 5;
    const bool vec_ok = BF16 || (d & 3) == 0;
    const int nvec = BF16 ? (d >> 3) : (d >> 2);       // 16-byte chunks per row
    const size_t row_bytes = (size_t)a.pitch * (BF16 ? 2 : 4);
    const float4* sq4 = reinterpret_cast<const float4*>(sq);
    constexpr int V = R * QB;
    constexpr int SH = 5 - Log2<V>::v;

    for (int64_t r0 = warp0 * R; r0 < a.n; r0 += nwarps * R) {
        float2 acc[R][QB];
#pragma unroll
        for (int i = 0; i < R; i++)
#pragma unroll
            for (int q = 0; q < QB; q++) acc[i][q] = make_float2(0.f, 0.f);

        if (vec_ok) {
            const char* xr[R];
#pragma unroll
            for (int i = 0; i < R; i++) {
                const int64_t row = r0 + i < a.n ? r0 + i : a.n - 1;   // tail rows re-read the last row, never stored
                xr[i] = reinterpret_cast<const char*>(a.x) + (size_t)row * row_bytes;
            }
#pragma unroll 2
            for (int c = lane; c < nvec; c += 32) {
                uint4 v[R];
#pragma unroll
                for (int i = 0; i < R; i++) v[i] = ldg_nc_16(xr[i] + (size_t)c * 16);
                if (BF16) {
                    float2 s[QB][4];
#pragma unroll
                    for (int q = 0; q < QB; q++) {
                        const float4 s0 = sq4[(q * d >> 2) + 2 * c], s1 = sq4[(q * d >> 2) + 2 * c + 1];
                        s[q][0] = make_float2(s0.x, s0.y); s[q][1] = make_float2(s0.z, s0.w);
                        s[q][2] = make_float2(s1.x, s1.y); s[q][3] = make_float2(s1.z, s1.w);
                    }
#pragma unroll
                    for (int i = 0; i < R; i++) {
                        const float2 x0 = bf16x2_to_f32x2(v[i].x), x1 = bf16x2_to_f32x2(v[i].y),
                                     x2 = bf16x2_to_f32x2(v[i].z), x3 = bf16x2_to_f32x2(v[i].w);
#pragma unroll
                        for (int q = 0; q < QB; q++) {
                            acc[i][q] = ffma2(x0, s[q][0], acc[i][q]);
                            acc[i][q] = ffma2(x1, s[q][1], acc[i][q]);
                            acc[i][q] = ffma2(x2, s[q][2], acc[i][q]);
                            acc[i][q] = ffma2(x3, s[q][3], acc[i][q]);
                        }
                    }
                } else {
                    float2 s[QB][2];
#pragma unroll
                    for (int q = 0; q < QB; q++) {
                        const float4 s0 = sq4[(q * d >> 2) + c];
                        s[q][0] = make_float2(s0.x, s0.y); s[q][1] = make_float2(s0.z, s0.w);
                    }
#pragma unroll
                    for (int i = 0; i < R; i++) {
                        const float2 x0 = make_float2(__uint_as_float(v[i].x), __uint_as_float(v[i].y));
                        const float2 x1 = make_float2(__uint_as_float(v[i].z), __uint_as_float(v[i].w));
#pragma unroll
                        for (int q = 0; q < QB; q++) {
                            if (METRIC == TRX_METRIC_INNER_PRODUCT) {
                                acc[i][q] = ffma2(x0, s[q][0], acc[i][q]);
                                acc[i][q] = ffma2(x1, s[q][1], acc[i][q]);
                            } else {
                                const float2 m1 = make_float2(-1.f, -1.f);
                                const float2 t0 = ffma2(x0, m1, s[q][0]), t1 = ffma2(x1, m1, s[q][1]);   // q - x
                                acc[i][q] = ffma2(t0, t0, acc[i][q]);
                                acc[i][q] = ffma2(t1, t1, acc[i][q]);
                            }
                        }
                    }
                }
            }
        } else {  // fp32 rows whose length is not a multiple of 4: not 16-byte aligned, scalar loads
#pragma unroll
            for (int i = 0; i < R; i++) {
                const int64_t row = r0 + i < a.n ? r0 + i : a.n - 1;
                const float* xr = reinterpret_cast<const float*>(a.x) + row * a.pitch;
                for (int c = lane; c < d; c += 32) {
                    const float x0 = __ldg(xr + c);
#pragma unroll
                    for (int q = 0; q < QB; q++) {
                        const float s = sq[q * d + c];
                        if (METRIC == TRX_METRIC_INNER_PRODUCT) acc[i][q].x = fmaf(x0, s, acc[i][q].x);
                        else { const float t = s - x0; acc[i][q].x = fmaf(t, t, acc[i][q].x); }
                    }
                }
            }
        }

        float v[V];
#pragma unroll
        for (int i = 0; i < R; i++)
#pragma unroll
            for (int q = 0; q < QB; q++) v[i * QB + q] = acc[i][q].x + acc[i][q].y;
        const float tot = warp_multi_reduce<V>(v, lane);
        if ((lane & ((1 << SH) - 1)) == 0) {
            const int idx = lane >> SH;
            const int i = idx / QB, q = idx - i * QB;
            const int64_t row = r0 + i;
            if (row < a.n && q < nqb) {
                float s = (!BF16 && METRIC == TRX_METRIC_L2) ? -tot : tot;
                if (APPEND) {
                    // group masking is applied by K4; here only the threshold test
                    if (s > s_thr[q]) {
                        uint32_t pos = atomicAdd(a.cand_cnt + (q0 + q), 1u);
                        if (pos < (uint32_t)a.cap) {
                            Cand c; c.score = s; c.row = (int32_t)row;
                            a.cand[(q0 + q) * (int64_t)a.cap + pos] = c;
                        }
                    }
                } else {
                    const int32_t ex = s_ex[q];
                    if (ex >= 0 && __ldg(a.groups + row) == ex) s = -INFINITY;
                    if (a.attr != nullptr && __ldg(a.attr + row) >= a.attr_below) s = -INFINITY;
                    a.out[(q0 + q) * a.out_ld + row] = s;
                }
            }
        }
    }
}

template <bool BF16, int METRIC, int QB>
static int launch_qb(const StreamArgs& a, int64_t q0, int grid, size_t smem, cudaStream_t st) {
    if (a.append) {
        auto kern = k3_stream_kernel<BF16, METRIC, QB, true>;
        if (smem > 48 * 1024) TRX_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        kern<<<grid, 256, smem, st>>>(a, q0);
    } else {
        auto kern = k3_stream_kernel<BF16, METRIC, QB, false>;
        if (smem > 48 * 1024) TRX_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        kern<<<grid, 256, smem, st>>>(a, q0);
    }
    count_launch();
    TRX_CUDA(cudaGetLastError());
    return TRX_OK;
}

int launch_stream(const StreamArgs& a, int sm_count, cudaStream_t st) {
    if (a.nq <= 0 || a.n <= 0) return TRX_OK;
    if (a.bf16 && (a.d & 7)) { set_error("k3: bf16 mode needs d %% 8 == 0"); return TRX_EINVAL; }
    // The corpus is re-read once per QB queries: QB = 4 quarters the traffic of a many-query exact scan
    // and still leaves the pass HBM-bound.  Persistent grid: 8 CTAs of 8 warps per SM at most.
    int64_t want = (a.n + 8 * R - 1) / (8 * R);
    int grid = (int)(want < (int64_t)sm_count * 8 ? want : (int64_t)sm_count * 8);
    for (int64_t q0 = 0; q0 < a.nq;) {
        int64_t left = a.nq - q0;
        int qb = left >= 4 ? 4 : (left >= 2 ? 2 : 1);
        while (qb > 1 && (size_t)qb * a.d * sizeof(float) > 200 * 1024) qb >>= 1;   // very wide rows: fewer queries per pass
        size_t smem = (size_t)qb * a.d * sizeof(float);
        if (smem > 200 * 1024) { set_error("k3: d=%d too large for the shared-memory query tile", a.d); return TRX_EINVAL; }
        int rc;
#define TRX_K3(BF, M)                                                       \
        (qb == 4 ? launch_qb<BF, M, 4>(a, q0, grid, smem, st)               \
                 : qb == 2 ? launch_qb<BF, M, 2>(a, q0, grid, smem, st)     \
                           : launch_qb<BF, M, 1>(a, q0, grid, smem, st))
        if (a.bf16) rc = TRX_K3(true, TRX_METRIC_INNER_PRODUCT);
        else if (a.metric == TRX_METRIC_L2) rc = TRX_K3(false, TRX_METRIC_L2);
        else rc = TRX_K3(false, TRX_METRIC_INNER_PRODUCT);
#undef TRX_K3
        TRX_TRY(rc);
        q0 += qb;
    }
    return TRX_OK;
}

}  // namespace trx
