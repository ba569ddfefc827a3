// k3_stream.cu -- CUDA-core streaming scorer (bandwidth-bound regime).
//
// Replaces the arithmetic of faiss `index.search` (retrieve/retrieve_faiss.py:71) for
//   * the exact path: fp32 corpus, fp32 FMA, direct sum (q-x)^2 for L2 -- this is also the
//     certified fallback every bf16-prefiltered query can fall back to;
//   * tiny batches: one pass over the bf16 corpus with 128-bit loads, candidates above the
//     sampled threshold appended on the fly (no score matrix in HBM).
// One warp per corpus row, QB queries per pass held in shared memory as fp32; every lane
// issues 16-byte loads (ld.global.nc, no L1 allocation) so a warp reads 512 contiguous bytes.
#include "common.cuh"

namespace trx {

__device__ __forceinline__ uint4 ldg_nc_16(const void* p) {
    uint4 r;
    asm volatile("ld.global.nc.L1::no_allocate.v4.u32 {%0,%1,%2,%3}, [%4];"
                 : "=r"(r.x), "=r"(r.y), "=r"(r.z), "=r"(r.w) : "l"(p));
    return r;
}

__device__ __forceinline__ float bf16lo(uint32_t u) { return __uint_as_float(u << 16); }
__device__ __forceinline__ float bf16hi(uint32_t u) { return __uint_as_float(u & 0xffff0000u); }

template <int QB>
__device__ __forceinline__ void warp_reduce_all(float (&acc)[QB]) {
#pragma unroll
    for (int q = 0; q < QB; q++) {
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) acc[q] += __shfl_xor_sync(0xffffffffu, acc[q], o);
    }
}

// BF16: x rows are bf16 with `pitch` elements, reduce over d (multiple of 8) columns, IP formula.
// !BF16: x rows fp32, reduce over d columns, METRIC selects dot or -(sum sq diff).
template <bool BF16, int METRIC, int QB, bool APPEND>
__global__ void __launch_bounds__(256) k3_stream_kernel(StreamArgs a, int64_t q0) {
    extern __shared__ float sq[];  // [QB][d] fp32 queries
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
    __syncthreads();

    const int lane = threadIdx.x & 31;
    const int64_t warp0 = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    const int64_t nwarps = ((int64_t)gridDim.x * blockDim.x) >> 5;
    int32_t ex[QB];
    float thr[QB];
#pragma unroll
    for (int q = 0; q < QB; q++) {
        ex[q] = (a.excl != nullptr && a.groups != nullptr && q < nqb) ? a.excl[q0 + q] : -1;
        thr[q] = (APPEND && q < nqb) ? a.thr[q0 + q] : 0.f;
    }

    for (int64_t r = warp0; r < a.n; r += nwarps) {
        float acc[QB];
#pragma unroll
        for (int q = 0; q < QB; q++) acc[q] = 0.f;
        if (BF16) {
            const char* xr = reinterpret_cast<const char*>(a.x) + r * a.pitch * 2;
            const int nv = d >> 3;
#pragma unroll 4
            for (int c = lane; c < nv; c += 32) {
                uint4 v = ldg_nc_16(xr + (size_t)c * 16);
                float xv[8] = {bf16lo(v.x), bf16hi(v.x), bf16lo(v.y), bf16hi(v.y),
                               bf16lo(v.z), bf16hi(v.z), bf16lo(v.w), bf16hi(v.w)};
#pragma unroll
                for (int q = 0; q < QB; q++) {
                    const float4* s4 = reinterpret_cast<const float4*>(sq + q * d + c * 8);
                    float4 s0 = s4[0], s1 = s4[1];
                    acc[q] = fmaf(xv[0], s0.x, acc[q]); acc[q] = fmaf(xv[1], s0.y, acc[q]);
                    acc[q] = fmaf(xv[2], s0.z, acc[q]); acc[q] = fmaf(xv[3], s0.w, acc[q]);
                    acc[q] = fmaf(xv[4], s1.x, acc[q]); acc[q] = fmaf(xv[5], s1.y, acc[q]);
                    acc[q] = fmaf(xv[6], s1.z, acc[q]); acc[q] = fmaf(xv[7], s1.w, acc[q]);
                }
            }
        } else if ((d & 3) == 0) {
            const char* xr = reinterpret_cast<const char*>(a.x) + r * a.pitch * 4;
            const int nv = d >> 2;
#pragma unroll 4
            for (int c = lane; c < nv; c += 32) {
                uint4 v = ldg_nc_16(xr + (size_t)c * 16);
                float x0 = __uint_as_float(v.x), x1 = __uint_as_float(v.y), x2 = __uint_as_float(v.z),
                      x3 = __uint_as_float(v.w);
#pragma unroll
                for (int q = 0; q < QB; q++) {
                    float4 s = *reinterpret_cast<const float4*>(sq + q * d + c * 4);
                    if (METRIC == TRX_METRIC_INNER_PRODUCT) {
                        acc[q] = fmaf(x0, s.x, acc[q]); acc[q] = fmaf(x1, s.y, acc[q]);
                        acc[q] = fmaf(x2, s.z, acc[q]); acc[q] = fmaf(x3, s.w, acc[q]);
                    } else {
                        float t0 = s.x - x0, t1 = s.y - x1, t2 = s.z - x2, t3 = s.w - x3;
                        acc[q] = fmaf(t0, t0, acc[q]); acc[q] = fmaf(t1, t1, acc[q]);
                        acc[q] = fmaf(t2, t2, acc[q]); acc[q] = fmaf(t3, t3, acc[q]);
                    }
                }
            }
        } else {  // d not a multiple of 4: rows are not 16-byte aligned, scalar loads
            const float* xr = reinterpret_cast<const float*>(a.x) + r * a.pitch;
            for (int c = lane; c < d; c += 32) {
                float x0 = __ldg(xr + c);
#pragma unroll
                for (int q = 0; q < QB; q++) {
                    float s = sq[q * d + c];
                    if (METRIC == TRX_METRIC_INNER_PRODUCT) acc[q] = fmaf(x0, s, acc[q]);
                    else { float t = s - x0; acc[q] = fmaf(t, t, acc[q]); }
                }
            }
        }
        warp_reduce_all<QB>(acc);
        if (lane == 0) {
            int32_t g = (a.groups != nullptr && a.excl != nullptr) ? __ldg(a.groups + r) : -2;
#pragma unroll
            for (int q = 0; q < QB; q++) {
                if (q >= nqb) break;
                float s = (!BF16 && METRIC == TRX_METRIC_L2) ? -acc[q] : acc[q];
                if (APPEND) {
                    // group masking is applied by K4; here only the threshold test
                    if (s > thr[q]) {
                        uint32_t pos = atomicAdd(a.cand_cnt + (q0 + q), 1u);
                        if (pos < (uint32_t)a.cap) {
                            Cand c; c.score = s; c.row = (int32_t)r;
                            a.cand[(q0 + q) * (int64_t)a.cap + pos] = c;
                        }
                    }
                } else {
                    if (ex[q] >= 0 && g == ex[q]) s = -INFINITY;
                    a.out[(q0 + q) * a.out_ld + r] = s;
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
    // rows per pass are re-read once per QB queries: QB=4 keeps the kernel HBM-bound
    // (4 LDS.128 per 16-byte global load) while quartering corpus traffic.
    int64_t want = ((a.n + 7) / 8);
    int grid = (int)(want < (int64_t)sm_count * 8 ? want : (int64_t)sm_count * 8);
    for (int64_t q0 = 0; q0 < a.nq;) {
        int64_t left = a.nq - q0;
        int qb = left >= 4 ? 4 : (left >= 2 ? 2 : 1);
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
