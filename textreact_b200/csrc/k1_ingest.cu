// k1_ingest.cu -- corpus / query ingestion kernels (HBM-bound, one warp per row).
//
// Replaces the storage half of faiss `index.add(train_fps)` (retrieve/retrieve_faiss.py:66):
// FAISS keeps one fp32 row-major copy; we keep that copy (for the exact rescore) plus a bf16
// copy laid out for TMA (row pitch Kp, a multiple of 64 elements = one 128-byte swizzle row).
#include <cuda_fp16.h>

#include "common.cuh"

namespace trx {

__device__ __forceinline__ float warp_sum(float v) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    return v;
}

// Split a non-negative fp32 value into three bf16 whose sum reproduces it to ~2^-24.
__device__ __forceinline__ void split3(float v, __nv_bfloat16& a, __nv_bfloat16& b, __nv_bfloat16& c) {
    a = __float2bfloat16_rn(v);
    float r1 = v - __bfloat162float(a);
    b = __float2bfloat16_rn(r1);
    float r2 = r1 - __bfloat162float(b);
    c = __float2bfloat16_rn(r2);
}

// grid-stride over rows, one warp per row.
__global__ void __launch_bounds__(256) k1_ingest_kernel(const float* __restrict__ x, int64_t n, int d, int Kp,
                                                        int metric, __nv_bfloat16* __restrict__ x16,
                                                        float* __restrict__ xnorm2,
                                                        uint32_t* __restrict__ norm2_max_bits /*[2]: max |x|^2, max |x - bf16(x)|^2*/) {
    const int lane = threadIdx.x & 31;
    const int64_t warp0 = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    const int64_t nwarps = ((int64_t)gridDim.x * blockDim.x) >> 5;
    float wmax = 0.f, emax = 0.f;
    for (int64_t r = warp0; r < n; r += nwarps) {
        const float* xr = x + r * (int64_t)d;
        __nv_bfloat16* yr = x16 + r * (int64_t)Kp;
        float acc = 0.f, err = 0.f;   // |x|^2 and |x - bf16(x)|^2 (each difference is exact in fp32)
        if ((d & 3) == 0) {
            const float4* x4 = reinterpret_cast<const float4*>(xr);
            for (int c = lane; c < (d >> 2); c += 32) {
                float4 v = __ldg(x4 + c);
                acc = fmaf(v.x, v.x, acc); acc = fmaf(v.y, v.y, acc);
                acc = fmaf(v.z, v.z, acc); acc = fmaf(v.w, v.w, acc);
                __nv_bfloat162 lo = __floats2bfloat162_rn(v.x, v.y);
                __nv_bfloat162 hi = __floats2bfloat162_rn(v.z, v.w);
                const float2 lf = __bfloat1622float2(lo), hf = __bfloat1622float2(hi);
                const float e0 = v.x - lf.x, e1 = v.y - lf.y, e2 = v.z - hf.x, e3 = v.w - hf.y;
                err = fmaf(e0, e0, err); err = fmaf(e1, e1, err); err = fmaf(e2, e2, err); err = fmaf(e3, e3, err);
                uint2 pk;
                pk.x = *reinterpret_cast<uint32_t*>(&lo);
                pk.y = *reinterpret_cast<uint32_t*>(&hi);
                *reinterpret_cast<uint2*>(yr + 4 * c) = pk;
            }
        } else {
            for (int c = lane; c < d; c += 32) {
                float v = __ldg(xr + c);
                acc = fmaf(v, v, acc);
                const __nv_bfloat16 b = __float2bfloat16_rn(v);
                const float e = v - __bfloat162float(b);
                err = fmaf(e, e, err);
                yr[c] = b;
            }
        }
        acc = warp_sum(acc);
        err = warp_sum(err);
        // padding columns (and the norm split for L2)
        for (int c = d + lane; c < Kp; c += 32) yr[c] = __float2bfloat16_rn(0.f);
        __syncwarp();
        if (lane == 0) {
            xnorm2[r] = acc;
            if (metric == TRX_METRIC_L2) {
                __nv_bfloat16 a, b, c3;
                split3(acc, a, b, c3);
                yr[d] = a; yr[d + 1] = b; yr[d + 2] = c3;
            }
            wmax = fmaxf(wmax, acc);
            emax = fmaxf(emax, err);
        }
    }
    if (lane == 0 && wmax > 0.f) atomicMax(norm2_max_bits, __float_as_uint(wmax));  // >= 0: bit order == value order
    if (lane == 0 && emax > 0.f) atomicMax(norm2_max_bits + 1, __float_as_uint(emax));
}

// Typed ingestion: the reference hands FAISS int8 Morgan bits and int64 difference counts
// (retrieve/retrieve_faiss.py:26, :40) and FAISS's wrapper converts them to fp32 on the host; here the raw
// array crosses PCIe as it is and is widened on the device.
template <typename T>
__global__ void __launch_bounds__(256) k1_widen_kernel(const T* __restrict__ src, int64_t count, float* __restrict__ dst) {
    const int64_t stride = (int64_t)gridDim.x * blockDim.x;
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < count; i += stride) dst[i] = (float)src[i];
}
template <>
__global__ void __launch_bounds__(256) k1_widen_kernel<__half>(const __half* __restrict__ src, int64_t count,
                                                                float* __restrict__ dst) {
    const int64_t stride = (int64_t)gridDim.x * blockDim.x;
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < count; i += stride) dst[i] = __half2float(src[i]);
}

int dtype_size(int dtype) {
    switch (dtype) {
        case TRX_DTYPE_F32: case TRX_DTYPE_I32: return 4;
        case TRX_DTYPE_F64: case TRX_DTYPE_I64: return 8;
        case TRX_DTYPE_F16: case TRX_DTYPE_I16: return 2;
        case TRX_DTYPE_I8: case TRX_DTYPE_U8: return 1;
    }
    return 0;
}

int launch_widen(const void* src, int dtype, int64_t count, float* dst, cudaStream_t st) {
    if (count <= 0) return TRX_OK;
    int64_t blocks = (count + 255) / 256;
    if (blocks > 148 * 32) blocks = 148 * 32;
    const unsigned g = (unsigned)blocks;
    switch (dtype) {
        case TRX_DTYPE_F64: k1_widen_kernel<double><<<g, 256, 0, st>>>((const double*)src, count, dst); break;
        case TRX_DTYPE_F16: k1_widen_kernel<__half><<<g, 256, 0, st>>>((const __half*)src, count, dst); break;
        case TRX_DTYPE_I8: k1_widen_kernel<int8_t><<<g, 256, 0, st>>>((const int8_t*)src, count, dst); break;
        case TRX_DTYPE_U8: k1_widen_kernel<uint8_t><<<g, 256, 0, st>>>((const uint8_t*)src, count, dst); break;
        case TRX_DTYPE_I16: k1_widen_kernel<int16_t><<<g, 256, 0, st>>>((const int16_t*)src, count, dst); break;
        case TRX_DTYPE_I32: k1_widen_kernel<int32_t><<<g, 256, 0, st>>>((const int32_t*)src, count, dst); break;
        case TRX_DTYPE_I64: k1_widen_kernel<long long><<<g, 256, 0, st>>>((const long long*)src, count, dst); break;
        default: set_error("widen: bad dtype %d", dtype); return TRX_EINVAL;
    }
    count_launch();
    TRX_CUDA(cudaGetLastError());
    return TRX_OK;
}

int launch_ingest(const float* x, int64_t n, int d, int Kp, int metric, __nv_bfloat16* x16, float* xnorm2,
                  uint32_t* norm2_max_bits, cudaStream_t st) {
    if (n <= 0) return TRX_OK;
    int64_t blocks = (n + 7) / 8;
    if (blocks > 148 * 16) blocks = 148 * 16;
    k1_ingest_kernel<<<(unsigned)blocks, 256, 0, st>>>(x, n, d, Kp, metric, x16, xnorm2, norm2_max_bits);
    count_launch();
    TRX_CUDA(cudaGetLastError());
    return TRX_OK;
}

// Sample row j of group g = rows [g*rate, (g+1)*rate): a hash picks the member, so that periodic
// structure in the add order cannot alias with the sampling stride.
__host__ __device__ __forceinline__ uint32_t mix32(uint32_t h) {
    h ^= h >> 16; h *= 0x7feb352du; h ^= h >> 15; h *= 0x846ca68bu; h ^= h >> 16;
    return h;
}

__global__ void __launch_bounds__(256) k1_sample_gather_kernel(const __nv_bfloat16* __restrict__ x16, int64_t n,
                                                               int Kp, int rate,
                                                               __nv_bfloat16* __restrict__ xs16, int64_t ns) {
    const int lane = threadIdx.x & 31;
    const int64_t warp0 = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    const int64_t nwarps = ((int64_t)gridDim.x * blockDim.x) >> 5;
    const int vec = Kp >> 3;  // 16-byte chunks per row (Kp % 64 == 0)
    for (int64_t g = warp0; g < ns; g += nwarps) {
        int64_t base = g * rate;
        int64_t span = n - base < rate ? n - base : rate;
        int64_t src = base + (int64_t)(mix32((uint32_t)g * 2654435761u + 12345u) % (uint32_t)span);
        const uint4* s4 = reinterpret_cast<const uint4*>(x16 + src * (int64_t)Kp);
        uint4* d4 = reinterpret_cast<uint4*>(xs16 + g * (int64_t)Kp);
        for (int c = lane; c < vec; c += 32) d4[c] = __ldg(s4 + c);
    }
}

int launch_sample_gather(const __nv_bfloat16* x16, int64_t n, int Kp, int rate, __nv_bfloat16* xs16, int64_t ns,
                         cudaStream_t st) {
    if (ns <= 0) return TRX_OK;
    int64_t blocks = (ns + 7) / 8;
    if (blocks > 148 * 16) blocks = 148 * 16;
    k1_sample_gather_kernel<<<(unsigned)blocks, 256, 0, st>>>(x16, n, Kp, rate, xs16, ns);
    count_launch();
    TRX_CUDA(cudaGetLastError());
    return TRX_OK;
}

// Certificate slack (see DESIGN.md "exactness certificate"): a rigorous bound on |prefilter score - fp32 score|.
//   rounding of the operands to bf16 (q^ = bf16(q), x^ = bf16(x)):
//       q^.x^ - q.x = q^.(x^ - x) + x.(q^ - q)   =>   |.| <= |q^| |x^ - x| + |x| |q^ - q|       (Cauchy-Schwarz)
//     |x^ - x| is MEASURED per row at ingest (its maximum over the corpus is kept next to max |x|), |q^ - q| and
//     |q^| per query here: no worst-case rounding model (which would give (2u + u^2)|q||x| with u = 2^-8, the bf16
//     unit roundoff) -- typically 3x tighter, and exactly 0 for integer-valued fingerprints.
//   fp32 accumulation (tensor core prefilter) and fp32 rescore: each <= d * 2^-23 |q||x| (loose)
//   IP : eps = |q^| max|x^-x| + max|x| |q^-q| + 2 d 2^-23 |q| max|x|
//   L2 : prefilter score is 2 q.x - |x|^2  ->  2x the IP slack, + 2^-22 max|x|^2 for the 3-way norm
//        split and + 2^-21 (|q|^2 + max|x|^2) for the fp32 evaluation of |q|^2 - sum (q-x)^2.
__device__ __forceinline__ void certificate_slack(float qn2, float qhat2, float qerr2, float xm2, float xerr2m, int d,
                                                  int metric, float& eps, float& eps_acc) {
    float qn = sqrtf(qn2) * 1.000001f, xn = sqrtf(xm2) * 1.000001f;
    float cacc = 2.f * (float)(d + 3) * 1.1920929e-7f;
    // the squared error norms are themselves fp32 sums of d terms: inflate by (1 + d 2^-23) and a fixed 0.1 %
    float infl = 1.001f + (float)d * 1.1920929e-7f;
    float rnd = (sqrtf(qhat2) * sqrtf(xerr2m) + xn * sqrtf(qerr2)) * infl;
    float e = rnd + cacc * qn * xn, ea = cacc * qn * xn;
    if (metric == TRX_METRIC_L2) {
        float extra = 2.4e-7f * xm2 + 4.8e-7f * (qn2 + xm2);
        e = 2.f * e + extra;
        ea = cacc * (qn2 + xm2) * 2.f + extra;  // sum (q-x)^2 <= 2(|q|^2+|x|^2)
    }
    eps = e * 1.0001f + 1e-37f;
    eps_acc = ea * 1.0001f + 1e-37f;   // two fp32 evaluations of the same score differ by at most this
}

// Start of a batch, one launch: queries fp32 -> bf16 (+|q|^2), certificate slack, and the per-batch
// counters (candidate counts, fallback count) zeroed.
__global__ void __launch_bounds__(256) k1_query_prep_kernel(const float* __restrict__ q, int64_t B, int d, int Kp,
                                                            int metric, __nv_bfloat16* __restrict__ q16,
                                                            float* __restrict__ qnorm2,
                                                            const uint32_t* __restrict__ norm2_max_bits,
                                                            float* __restrict__ eps, float* __restrict__ eps_acc,
                                                            uint32_t* __restrict__ cand_cnt,
                                                            uint32_t* __restrict__ fb_count) {
    const int lane = threadIdx.x & 31;
    const int64_t warp0 = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    const int64_t nwarps = ((int64_t)gridDim.x * blockDim.x) >> 5;
    const float scale = metric == TRX_METRIC_L2 ? 2.f : 1.f;  // exact in bf16
    if (fb_count != nullptr && blockIdx.x == 0 && threadIdx.x == 0) *fb_count = 0u;
    for (int64_t r = warp0; r < B; r += nwarps) {
        const float* qr = q + r * (int64_t)d;
        __nv_bfloat16* yr = q16 + r * (int64_t)Kp;
        float acc = 0.f, hat = 0.f, err = 0.f;   // |q|^2, |bf16(q)|^2, |q - bf16(q)|^2
        for (int c = lane; c < d; c += 32) {
            float v = __ldg(qr + c);
            acc = fmaf(v, v, acc);
            const float vb = __bfloat162float(__float2bfloat16_rn(v));
            hat = fmaf(vb, vb, hat);
            err = fmaf(v - vb, v - vb, err);
            yr[c] = __float2bfloat16_rn(scale * vb);
        }
        for (int c = d + lane; c < Kp; c += 32)
            yr[c] = __float2bfloat16_rn((metric == TRX_METRIC_L2 && c < d + 3) ? -1.f : 0.f);
        acc = warp_sum(acc);
        hat = warp_sum(hat);
        err = warp_sum(err);
        if (lane == 0) {
            qnorm2[r] = acc;
            if (eps != nullptr) {
                float e, ea;
                certificate_slack(acc, hat, err, __uint_as_float(norm2_max_bits[0]), __uint_as_float(norm2_max_bits[1]),
                                  d, metric, e, ea);
                eps[r] = e; eps_acc[r] = ea;
            }
            if (cand_cnt != nullptr) cand_cnt[r] = 0u;
        }
    }
}

int launch_query_prep(const float* q, int64_t B, int d, int Kp, int metric, __nv_bfloat16* q16, float* qnorm2,
                      const uint32_t* norm2_max_bits, float* eps, float* eps_acc, uint32_t* cand_cnt,
                      uint32_t* fb_count, cudaStream_t st) {
    if (B <= 0) return TRX_OK;
    int64_t blocks = (B + 7) / 8;
    if (blocks > 148 * 8) blocks = 148 * 8;
    k1_query_prep_kernel<<<(unsigned)blocks, 256, 0, st>>>(q, B, d, Kp, metric, q16, qnorm2, norm2_max_bits, eps,
                                                           eps_acc, cand_cnt, fb_count);
    count_launch();
    TRX_CUDA(cudaGetLastError());
    return TRX_OK;
}

}  // namespace trx
