// k2_score_umma.cu -- the batched scoring kernel: TMA-fed tcgen05.mma tiles over the bf16 corpus,
// fp32 accumulators in TMEM, per-query selection fused into the epilogue so the score matrix
// never reaches HBM.
//
// Replaces the `sgemm` + heap/reservoir inner loop of faiss `index.search`
// (retrieve/retrieve_faiss.py:71) for batched queries.
//
// Three tilings of the same kernel (template parameters PAIR, AR):
//   PAIR = false  one CTA per tile: D[128 q x 256 rows] += A[128x64] . B[256x64]^T,
//                 tcgen05.mma.cta_group::1.
//                 AR = 128: 4-stage 48 KB smem ring (33..128 queries per tile)
//                 AR = 32 : the small-batch variant (<= 32 queries): the A slot of a stage is 4 KB, so the
//                           ring holds 6 stages = 192 KB of corpus in flight per SM -- this regime is
//                           HBM-bound and bytes in flight are what buys bandwidth.  The MMA still has
//                           M = 128; lanes >= 32 multiply whatever follows in the stage and are ignored.
//                 In both, TMA only moves the query rows that exist (box of round8(nq) rows) when the
//                 batch is a single tile: no out-of-bounds zero fill, less L2->smem traffic.
//   PAIR = true   a CTA pair (cluster of 2, one TPC) per 256 q x 256 rows tile,
//                 tcgen05.mma.cta_group::2 (M=256): each CTA stages its own 128 query rows and HALF of
//                 the corpus tile (32 KB / stage -> 6 stages), the leader CTA issues the MMAs for both,
//                 each CTA's TMEM receives its own 128 query rows x 256 columns.  Halves the corpus
//                 bytes every SM pulls from L2 and deepens the pipeline.
//
// Work order (STORE/THRESH): corpus-tile major per worker -- worker w owns corpus tiles w, w+W, ... and runs
// every query tile against each before moving on, so a corpus tile is fetched from HBM once (by one SM /
// SM pair) and re-read from L2 for the other query tiles: DRAM traffic == one pass over the bf16 corpus.
// Roles (256 threads per CTA):
//   warp 0    TMA producer: cp.async.bulk.tensor into the 128B-swizzled smem ring
//   warp 1    MMA issuer (leader CTA only in PAIR mode): one elected lane issues tcgen05.mma (K=16 steps)
//   warp 2    TMEM allocator (512 columns = two 128x256 fp32 accumulators, double buffered)
//   warps 4-7 epilogue: tcgen05.ld 32x32b.x32 (software pipelined), one query row per thread
// Epilogue modes:
//   STORE    fp32 scores to HBM (tests / profiling only)
//   THRESH   compare against the per-query threshold; hits go to a per-thread private log in HBM with
//            a register cursor (no atomics, fire-and-forget 16-byte stores), and a scatter kernel
//            turns the logs into per-query candidate lists afterwards
//   SLOTMAX  running max of every (column mod 32) slot -> threshold estimation on the sample
#include <cudaTypedefs.h>

#include "common.cuh"

namespace trx {

namespace {

constexpr int BM = 128;      // query rows per CTA (TMEM lanes)
constexpr int BN = 256;      // corpus rows per tile (UMMA N, TMEM columns per accumulator)
constexpr int BK = 64;       // bf16 elements per k-block = one 128-byte swizzle row
constexpr int UMMA_K = 16;
constexpr int TMEM_COLS = 512;
constexpr int NUM_THREADS = 256;

template <bool PAIR, int AR>
struct Cfg {
    static_assert(AR == 128 || (AR == 32 && !PAIR), "A slot: 128 rows, or 32 for the small-batch single-CTA variant");
    static constexpr int STAGES = (PAIR || AR == 32) ? 6 : 4;
    static constexpr int A_BYTES = AR * BK * 2;                      // 16 KB (4 KB when AR == 32)
    static constexpr int B_ROWS = PAIR ? BN / 2 : BN;                // corpus rows THIS CTA stages
    static constexpr int B_BYTES = B_ROWS * BK * 2;                  // 16 / 32 KB
    static constexpr int STAGE_BYTES = A_BYTES + B_BYTES;
    static constexpr int TX_SCALE = PAIR ? 2 : 1;                    // both CTAs' bytes land on the leader's full barrier
    static constexpr int SMEM_BYTES = STAGES * STAGE_BYTES + 1024 /*align*/ + 256 /*barriers*/;
    static constexpr int TILE_M = PAIR ? 2 * BM : BM;                // query rows per work tile
    static constexpr int UMMA_M = TILE_M;
    // kind::f16 instruction descriptor: D=f32, A=B=bf16, both K-major, N=256, M=128/256
    static constexpr uint32_t IDESC = (1u << 4) | (1u << 7) | (1u << 10) | ((uint32_t)(BN >> 3) << 17) |
                                      ((uint32_t)(UMMA_M >> 4) << 24);
};

enum { MODE_STORE = 0, MODE_THRESH = 1, MODE_SLOTMAX = 2 };

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }

__device__ __forceinline__ void mbar_init(uint32_t bar, uint32_t count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(bar), "r"(count));
}
__device__ __forceinline__ void mbar_expect_tx(uint32_t bar, uint32_t bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"(bytes) : "memory");
}
// arrive on the barrier at the same smem offset in CTA `cta` of the cluster
__device__ __forceinline__ void mbar_arrive_cluster(uint32_t bar, uint32_t cta) {
    asm volatile(
        "{\n\t"
        ".reg .b32 ra;\n\t"
        "mapa.shared::cluster.u32 ra, %0, %1;\n\t"
        "mbarrier.arrive.shared::cluster.b64 _, [ra];\n\t"
        "}" ::"r"(bar), "r"(cta) : "memory");
}
__device__ __forceinline__ void mbar_arrive(uint32_t bar) {
    asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(bar) : "memory");
}
__device__ __forceinline__ uint64_t global_timer_ns() {
    uint64_t t;
    asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t));
    return t;
}
// Blocks until the barrier phase completes.  A wait that lasts longer than kWaitTimeoutNs (a lost TMA transaction,
// a peer CTA that died) traps: the launch fails with a CUDA error instead of hanging the GPU.
constexpr uint64_t kWaitTimeoutNs = 20ull * 1000 * 1000 * 1000;
__device__ __forceinline__ void mbar_wait(uint32_t bar, uint32_t parity) {
    uint32_t done;
    asm volatile(
        "{\n\t"
        ".reg .pred p;\n\t"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
        "selp.u32 %0, 1, 0, p;\n\t"
        "}" : "=r"(done) : "r"(bar), "r"(parity) : "memory");
    if (done) return;
    const uint64_t t0 = global_timer_ns();
    uint32_t polls = 0;
    for (;;) {
        asm volatile(
            "{\n\t"
            ".reg .pred p;\n\t"
            "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
            "selp.u32 %0, 1, 0, p;\n\t"
            "}" : "=r"(done) : "r"(bar), "r"(parity) : "memory");
        if (done) return;
        if ((++polls & 0xfffu) == 0 && global_timer_ns() - t0 > kWaitTimeoutNs) __trap();
    }
}
__device__ __forceinline__ bool elect_one() {
    uint32_t pred;
    asm volatile(
        "{\n\t"
        ".reg .pred p;\n\t"
        "elect.sync _|p, 0xffffffff;\n\t"
        "selp.u32 %0, 1, 0, p;\n\t"
        "}" : "=r"(pred));
    return pred != 0;
}
__device__ __forceinline__ uint32_t cluster_ctarank() {
    uint32_t r;
    asm volatile("mov.u32 %0, %%cluster_ctarank;" : "=r"(r));
    return r;
}
__device__ __forceinline__ void cluster_sync() {
    asm volatile("barrier.cluster.arrive.release.aligned;" ::: "memory");
    asm volatile("barrier.cluster.wait.acquire.aligned;" ::: "memory");
}
template <bool PAIR>
__device__ __forceinline__ void tma_load_2d(uint32_t dst, const CUtensorMap* map, uint32_t bar, int c0, int c1) {
    if (PAIR) {
        // executed by both CTAs of the pair; the peer bit of the barrier address is cleared so the
        // transaction bytes of both land on the LEADER's full barrier
        asm volatile(
            "cp.async.bulk.tensor.2d.cta_group::2.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4}], [%2];"
            ::"r"(dst), "l"(reinterpret_cast<uint64_t>(map)), "r"(bar & 0xFEFFFFFFu), "r"(c0), "r"(c1) : "memory");
    } else {
        asm volatile(
            "cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4}], [%2];"
            ::"r"(dst), "l"(reinterpret_cast<uint64_t>(map)), "r"(bar), "r"(c0), "r"(c1) : "memory");
    }
}
__device__ __forceinline__ void tc_fence_after() { asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tc_fence_before() { asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory"); }
template <bool PAIR>
__device__ __forceinline__ void tc_commit(uint32_t bar) {
    if (PAIR) {  // arrive on the barrier at this offset in BOTH CTAs of the pair
        asm volatile("tcgen05.commit.cta_group::2.mbarrier::arrive::one.shared::cluster.multicast::cluster.b64 [%0], %1;"
                     ::"r"(bar), "h"((uint16_t)3) : "memory");
    } else {
        asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(bar) : "memory");
    }
}
template <bool PAIR>
__device__ __forceinline__ void tc_mma(uint32_t tmem_d, uint64_t adesc, uint64_t bdesc, uint32_t idesc, uint32_t acc) {
    if (PAIR) {
        asm volatile(
            "{\n\t"
            ".reg .pred p;\n\t"
            "setp.ne.b32 p, %4, 0;\n\t"
            "tcgen05.mma.cta_group::2.kind::f16 [%0], %1, %2, %3, p;\n\t"
            "}" ::"r"(tmem_d), "l"(adesc), "l"(bdesc), "r"(idesc), "r"(acc) : "memory");
    } else {
        asm volatile(
            "{\n\t"
            ".reg .pred p;\n\t"
            "setp.ne.b32 p, %4, 0;\n\t"
            "tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n\t"
            "}" ::"r"(tmem_d), "l"(adesc), "l"(bdesc), "r"(idesc), "r"(acc) : "memory");
    }
}
__device__ __forceinline__ void tc_ld32(uint32_t taddr, uint32_t (&v)[32]) {
    asm volatile(
        "tcgen05.ld.sync.aligned.32x32b.x32.b32 "
        "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, "
        "%16, %17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31}, [%32];"
        : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]), "=r"(v[7]),
          "=r"(v[8]), "=r"(v[9]), "=r"(v[10]), "=r"(v[11]), "=r"(v[12]), "=r"(v[13]), "=r"(v[14]), "=r"(v[15]),
          "=r"(v[16]), "=r"(v[17]), "=r"(v[18]), "=r"(v[19]), "=r"(v[20]), "=r"(v[21]), "=r"(v[22]), "=r"(v[23]),
          "=r"(v[24]), "=r"(v[25]), "=r"(v[26]), "=r"(v[27]), "=r"(v[28]), "=r"(v[29]), "=r"(v[30]), "=r"(v[31])
        : "r"(taddr));
}
__device__ __forceinline__ void tc_wait_ld() { asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory"); }

// UMMA shared-memory descriptor, K-major, SWIZZLE_128B: 8-row x 128-byte atoms, SBO = 1024 B.
__device__ __forceinline__ uint64_t make_smem_desc(uint32_t saddr) {
    uint64_t d = 0;
    d |= (uint64_t)((saddr >> 4) & 0x3fffu);  // start address
    d |= (uint64_t)1 << 16;                   // LBO (unused for swizzled K-major; CUTLASS writes 1)
    d |= (uint64_t)(1024 >> 4) << 32;         // SBO
    d |= (uint64_t)1 << 46;                   // descriptor version (Blackwell)
    d |= (uint64_t)2 << 61;                   // SWIZZLE_128B
    return d;
}

struct KParams {
    int64_t nq, n;
    int KB;          // k-blocks
    int MT, NT, S;   // query tiles (of TILE_M rows), row tiles, slices (SLOTMAX only)
    int a_box;       // query rows one A load moves (TMA box rows), <= AR
    float* out; int64_t out_ld;
    const float* thr; uint32_t* cand_cnt;
    HitRec* log; uint32_t* log_cnt; int log_cap;   // THRESH: [grid*128][log_cap] private hit logs
    float* slots;    // SLOTMAX: [nq][S][32]
};

// Work units of one worker (CTA or CTA pair), identical for the three roles.
//   SLOTMAX      unit = (query tile, slice of consecutive corpus tiles), round-robin over workers
//   STORE/THRESH unit = (corpus tile, query tile), corpus-tile major per worker (see header)
struct Sched {
    int MT, NT, S, W, w, R, rem;
    bool slotmax;
    __device__ __forceinline__ Sched(const KParams& p, int worker, int nworkers, bool sm)
        : MT(p.MT), NT(p.NT), S(p.S), W(nworkers), w(worker), R(p.NT / nworkers), rem(p.NT % nworkers), slotmax(sm) {}
    __device__ __forceinline__ int count() const {
        if (slotmax) { const int u = MT * S; return u > w ? (u - w + W - 1) / W : 0; }
        const int tail = rem * MT;
        return R * MT + (tail > w ? (tail - w + W - 1) / W : 0);
    }
    __device__ __forceinline__ void get(int i, int& mt, int& nt0, int& nt1, int& sl) const {
        if (slotmax) {
            const int u = w + i * W;
            mt = u % MT; sl = u / MT;
            nt0 = (int)((int64_t)sl * NT / S); nt1 = (int)((int64_t)(sl + 1) * NT / S);
        } else {
            if (i < R * MT) { const int r = i / MT; mt = i - r * MT; nt0 = r * W + w; }
            else { const int j = w + (i - R * MT) * W; const int t = j / MT; mt = j - t * MT; nt0 = R * W + t; }
            nt1 = nt0 + 1; sl = nt0;
        }
    }
};

}  // namespace

template <int MODE, bool PAIR, int AR>
__global__ void __launch_bounds__(NUM_THREADS, 1)
k2_umma_kernel(const __grid_constant__ CUtensorMap tmap_q, const __grid_constant__ CUtensorMap tmap_x, KParams p) {
    using C = Cfg<PAIR, AR>;
    constexpr int STAGES = C::STAGES;
    extern __shared__ unsigned char smem_dyn[];
    const uint32_t smem_base = (smem_u32(smem_dyn) + 1023u) & ~1023u;
    const uint32_t bar_base = smem_base + STAGES * C::STAGE_BYTES;
    // barrier layout (8 bytes each): full[STAGES] empty[STAGES] tfull[2] tempty[2] ; then tmem ptr
    auto full_bar = [&](int s) { return bar_base + 8u * s; };
    auto empty_bar = [&](int s) { return bar_base + 8u * (STAGES + s); };
    auto tfull_bar = [&](int s) { return bar_base + 8u * (2 * STAGES + s); };
    auto tempty_bar = [&](int s) { return bar_base + 8u * (2 * STAGES + 2 + s); };
    const uint32_t tmem_slot = bar_base + 8u * (2 * STAGES + 4);
    volatile uint32_t* tmem_slot_ptr =
        reinterpret_cast<volatile uint32_t*>(smem_dyn + (tmem_slot - smem_u32(smem_dyn)));

    const int warp = threadIdx.x >> 5;
    const int lane = threadIdx.x & 31;
    const uint32_t rank = PAIR ? cluster_ctarank() : 0u;      // position in the CTA pair
    const bool leader = rank == 0;
    const int worker = PAIR ? (int)(blockIdx.x >> 1) : (int)blockIdx.x;    // tile-processing unit id
    const int nworkers = PAIR ? (int)(gridDim.x >> 1) : (int)gridDim.x;

    if (warp == 0 && lane == 0) {
        asm volatile("prefetch.tensormap [%0];" ::"l"(reinterpret_cast<uint64_t>(&tmap_q)) : "memory");
        asm volatile("prefetch.tensormap [%0];" ::"l"(reinterpret_cast<uint64_t>(&tmap_x)) : "memory");
    }
    if (warp == 1 && lane == 0) {
        for (int s = 0; s < STAGES; s++) { mbar_init(full_bar(s), 1); mbar_init(empty_bar(s), 1); }
        // tempty: one arrival per epilogue warp of every CTA feeding this accumulator
        for (int s = 0; s < 2; s++) { mbar_init(tfull_bar(s), 1); mbar_init(tempty_bar(s), PAIR ? 8 : 4); }
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    if (warp == 2) {
        if (PAIR) {
            asm volatile("tcgen05.alloc.cta_group::2.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(tmem_slot), "r"(TMEM_COLS) : "memory");
            asm volatile("tcgen05.relinquish_alloc_permit.cta_group::2.sync.aligned;" ::: "memory");
        } else {
            asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(tmem_slot), "r"(TMEM_COLS) : "memory");
            asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
        }
    }
    tc_fence_before();
    if (PAIR) cluster_sync(); else __syncthreads();
    tc_fence_after();
    const uint32_t tmem_base = *tmem_slot_ptr;

    const Sched sched(p, worker, nworkers, MODE == MODE_SLOTMAX);
    const int my_units = sched.count();
    const uint32_t tx_bytes = (uint32_t)(C::TX_SCALE * (p.a_box * BK * 2 + C::B_BYTES));

    if (warp == 0) {
        // ===================== TMA producer (every CTA stages its own operands) =====================
        if (elect_one()) {
            int stage = 0; uint32_t phase = 0;
            for (int ui = 0; ui < my_units; ui++) {
                int mt, nt0, nt1, sl;
                sched.get(ui, mt, nt0, nt1, sl);
                for (int nt = nt0; nt < nt1; nt++) {
                    for (int kb = 0; kb < p.KB; kb++) {
                        mbar_wait(empty_bar(stage), phase ^ 1u);
                        if (leader) mbar_expect_tx(full_bar(stage), tx_bytes);
                        const uint32_t sa = smem_base + stage * C::STAGE_BYTES;
                        tma_load_2d<PAIR>(sa, &tmap_q, full_bar(stage), kb * BK, mt * C::TILE_M + (int)rank * BM);
                        tma_load_2d<PAIR>(sa + C::A_BYTES, &tmap_x, full_bar(stage), kb * BK,
                                          nt * BN + (int)rank * C::B_ROWS);
                        if (++stage == STAGES) { stage = 0; phase ^= 1u; }
                    }
                }
            }
        }
    } else if (warp == 1) {
        // ===================== MMA issuer (leader CTA drives both tensor cores) =====================
        if (leader) {
            int stage = 0; uint32_t phase = 0;
            int as = 0; uint32_t aphase = 0;
            for (int ui = 0; ui < my_units; ui++) {
                int mt, nt0, nt1, sl;
                sched.get(ui, mt, nt0, nt1, sl);
                for (int nt = nt0; nt < nt1; nt++) {
                    mbar_wait(tempty_bar(as), aphase ^ 1u);
                    tc_fence_after();
                    const uint32_t tmem_d = tmem_base + (uint32_t)(as * BN);
                    for (int kb = 0; kb < p.KB; kb++) {
                        mbar_wait(full_bar(stage), phase);
                        tc_fence_after();
                        if (elect_one()) {
                            const uint32_t sa = smem_base + stage * C::STAGE_BYTES;
                            const uint64_t adesc = make_smem_desc(sa);
                            const uint64_t bdesc = make_smem_desc(sa + C::A_BYTES);
#pragma unroll
                            for (int k = 0; k < BK / UMMA_K; k++) {
                                // advance 16 elements = 32 bytes inside the 128-byte swizzle row
                                tc_mma<PAIR>(tmem_d, adesc + (uint64_t)(2 * k), bdesc + (uint64_t)(2 * k), C::IDESC,
                                             (kb | k) != 0 ? 1u : 0u);
                            }
                            tc_commit<PAIR>(empty_bar(stage));                    // frees the smem slot(s) when the MMAs retire
                            if (kb == p.KB - 1) tc_commit<PAIR>(tfull_bar(as));   // accumulator(s) ready
                        }
                        __syncwarp();
                        if (++stage == STAGES) { stage = 0; phase ^= 1u; }
                    }
                    if (++as == 2) { as = 0; aphase ^= 1u; }
                }
            }
        }
    } else if (warp >= 4) {
        // ===================== epilogue =====================
        const int wq = warp & 3;  // TMEM lane quarter this warp may access
        int as = 0; uint32_t aphase = 0;
        const int64_t log_id = (int64_t)blockIdx.x * 128 + (threadIdx.x - 128);
        HitRec* my_log = MODE == MODE_THRESH ? p.log + log_id * p.log_cap : nullptr;
        uint32_t nlog = 0;
        // threshold of the NEXT unit's query row is fetched one unit ahead (the query tile changes every unit)
        auto row_of = [&](int mt) { return (int64_t)mt * C::TILE_M + rank * BM + wq * 32 + lane; };
        float thr_next = INFINITY;
        if (MODE == MODE_THRESH && my_units > 0) {
            int mt, a, b, c;
            sched.get(0, mt, a, b, c);
            if (row_of(mt) < p.nq) thr_next = p.thr[row_of(mt)];
        }
        for (int ui = 0; ui < my_units; ui++) {
            int mt, nt0, nt1, sl;
            sched.get(ui, mt, nt0, nt1, sl);
            const int64_t row = row_of(mt);  // query of this thread
            const bool row_ok = row < p.nq;
            const float thr = thr_next;
            if (MODE == MODE_THRESH) {
                thr_next = INFINITY;
                if (ui + 1 < my_units) {
                    int mt2, a, b, c;
                    sched.get(ui + 1, mt2, a, b, c);
                    if (row_of(mt2) < p.nq) thr_next = p.thr[row_of(mt2)];
                }
            }
            // a warp whose 32 query rows do not exist only hands the accumulator back
            const bool warp_idle = (int64_t)mt * C::TILE_M + rank * BM + wq * 32 >= p.nq;
            float slot[32];
            if (MODE == MODE_SLOTMAX) {
#pragma unroll
                for (int j = 0; j < 32; j++) slot[j] = -INFINITY;
            }
            for (int nt = nt0; nt < nt1; nt++) {
                mbar_wait(tfull_bar(as), aphase);
                tc_fence_after();
                if (warp_idle) {
                    tc_fence_before();
                    __syncwarp();
                    if (lane == 0) {
                        if (PAIR) mbar_arrive_cluster(tempty_bar(as), 0);
                        else mbar_arrive(tempty_bar(as));
                    }
                    if (++as == 2) { as = 0; aphase ^= 1u; }
                    continue;
                }
                const int64_t col0 = (int64_t)nt * BN;
                const bool full_tile = col0 + BN <= p.n;
                const uint32_t tbase = tmem_base + ((uint32_t)(wq * 32) << 16) + (uint32_t)(as * BN);
                // One 32-column chunk of this thread's query row.
                auto process = [&](const uint32_t (&v)[32], int ch) {
                    const int64_t cbase = col0 + ch * 32;
                    if (MODE == MODE_STORE) {
                        if (row_ok) {
#pragma unroll
                            for (int j = 0; j < 32; j++)
                                if (cbase + j < p.n) p.out[row * p.out_ld + cbase + j] = __uint_as_float(v[j]);
                        }
                    } else if (MODE == MODE_THRESH) {
                        // common case: nothing in the chunk beats the threshold -> 18 FMNMX + 1 compare
                        float g[4];
#pragma unroll
                        for (int gi = 0; gi < 4; gi++) {
                            float m = __uint_as_float(v[8 * gi]);
#pragma unroll
                            for (int j = 1; j < 8; j++) m = fmaxf(m, __uint_as_float(v[8 * gi + j]));
                            g[gi] = m;
                        }
                        const float mx = fmaxf(fmaxf(g[0], g[1]), fmaxf(g[2], g[3]));
                        if (mx > thr) {
                            // rare: scan only the 8-column groups that contain a hit
#pragma unroll
                            for (int gi = 0; gi < 4; gi++) {
                                if (g[gi] > thr) {
#pragma unroll
                                    for (int j = 0; j < 8; j++) {
                                        const float sc = __uint_as_float(v[8 * gi + j]);
                                        const int col = (int)cbase + 8 * gi + j;
                                        if (sc > thr && (full_tile || col < (int)p.n)) {
                                            if (nlog < (uint32_t)p.log_cap) {
                                                int4 rec = make_int4((int)row, col, __float_as_int(sc), 0);
                                                *reinterpret_cast<int4*>(my_log + nlog) = rec;
                                            } else {
                                                atomicOr(p.cand_cnt + row, 0x80000000u);  // log full: query overflowed
                                            }
                                            nlog++;
                                        }
                                    }
                                }
                            }
                        }
                    } else {
                        if (full_tile) {
#pragma unroll
                            for (int j = 0; j < 32; j++) slot[j] = fmaxf(slot[j], __uint_as_float(v[j]));
                        } else {
#pragma unroll
                            for (int j = 0; j < 32; j++)
                                if (cbase + j < p.n) slot[j] = fmaxf(slot[j], __uint_as_float(v[j]));
                        }
                    }
                };
                // Software pipeline: the TMEM load of chunk c+1 is in flight while chunk c is scanned.
                uint32_t va[32], vb[32];
                tc_ld32(tbase, va);
#pragma unroll 1
                for (int ch = 0; ch < BN / 32; ch += 2) {
                    tc_wait_ld();
                    tc_ld32(tbase + (uint32_t)((ch + 1) * 32), vb);
                    process(va, ch);
                    tc_wait_ld();
                    if (ch + 2 < BN / 32) tc_ld32(tbase + (uint32_t)((ch + 2) * 32), va);
                    else {  // everything of this accumulator is in registers: hand TMEM back to the MMA warp
                        tc_fence_before();
                        __syncwarp();
                        if (lane == 0) {
                            if (PAIR) mbar_arrive_cluster(tempty_bar(as), 0);   // the leader's barrier
                            else mbar_arrive(tempty_bar(as));
                        }
                    }
                    process(vb, ch + 1);
                }
                if (++as == 2) { as = 0; aphase ^= 1u; }
            }
            if (MODE == MODE_SLOTMAX && row_ok) {
                float* dst = p.slots + (row * p.S + sl) * 32;
#pragma unroll
                for (int j = 0; j < 32; j += 4)
                    *reinterpret_cast<float4*>(dst + j) = make_float4(slot[j], slot[j + 1], slot[j + 2], slot[j + 3]);
            }
        }
        if (MODE == MODE_THRESH) p.log_cnt[log_id] = nlog < (uint32_t)p.log_cap ? nlog : (uint32_t)p.log_cap;
    }

    tc_fence_before();
    if (PAIR) cluster_sync(); else __syncthreads();   // PAIR: the peer may still target our barriers / TMEM
    if (warp == 2) {
        tc_fence_after();
        if (PAIR) asm volatile("tcgen05.dealloc.cta_group::2.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(TMEM_COLS) : "memory");
        else asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(TMEM_COLS) : "memory");
    }
}

// One warp per private log: entries -> per-query candidate lists.  The entries of one log come from one
// epilogue thread, i.e. from few distinct queries (one, for a single-tile batch), so the 32 entries a warp
// handles per step are aggregated by query first: one global atomic per distinct query per step.
__global__ void __launch_bounds__(256) k2_scatter_kernel(const HitRec* __restrict__ log,
                                                         const uint32_t* __restrict__ log_cnt, int nlogs, int log_cap,
                                                         Cand* __restrict__ cand, uint32_t* __restrict__ cand_cnt,
                                                         int cap) {
    const int lane = threadIdx.x & 31;
    const int w = (blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    if (w >= nlogs) return;
    const uint32_t n = log_cnt[w];
    const HitRec* L = log + (int64_t)w * log_cap;
    for (uint32_t i0 = 0; i0 < n; i0 += 32) {
        const uint32_t i = i0 + lane;
        const bool have = i < n;
        const uint32_t active = __ballot_sync(0xffffffffu, have);
        if (!have) continue;
        int4 raw = __ldg(reinterpret_cast<const int4*>(L + i));
        HitRec h = *reinterpret_cast<HitRec*>(&raw);
        const uint32_t peers = __match_any_sync(active, h.q);
        const int lead = __ffs(peers) - 1;
        uint32_t base = 0;
        if (lane == lead) base = atomicAdd(cand_cnt + h.q, (uint32_t)__popc(peers));
        base = __shfl_sync(peers, base, lead);
        const uint32_t pos = base + (uint32_t)__popc(peers & ((1u << lane) - 1u));
        if (pos < (uint32_t)cap) {
            Cand c; c.score = h.score; c.row = h.row;
            cand[(int64_t)h.q * cap + pos] = c;
        }
    }
}

// ---------------------------------------------------------------------------------------------
// host side
// ---------------------------------------------------------------------------------------------
namespace {

PFN_cuTensorMapEncodeTiled_v12000 g_encode = nullptr;

int make_map(CUtensorMap* map, const void* base, int64_t rows, int Kp, int box_rows) {
    cuuint64_t dims[2] = {(cuuint64_t)Kp, (cuuint64_t)rows};
    cuuint64_t strides[1] = {(cuuint64_t)Kp * 2};
    cuuint32_t box[2] = {(cuuint32_t)BK, (cuuint32_t)box_rows};
    cuuint32_t estr[2] = {1, 1};
    CUresult r = g_encode(map, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2, const_cast<void*>(base), dims, strides, box, estr,
                          CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                          CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (r != CUDA_SUCCESS) { set_error("cuTensorMapEncodeTiled failed: %d (rows=%lld Kp=%d)", (int)r, (long long)rows, Kp); return TRX_ECUDA; }
    return TRX_OK;
}

template <int MODE, bool PAIR, int AR>
int launch_mode(const CUtensorMap& mq, const CUtensorMap& mx, const KParams& p, int grid, cudaStream_t st) {
    auto kern = k2_umma_kernel<MODE, PAIR, AR>;
    using C = Cfg<PAIR, AR>;
    // per device (an attribute set on one device does not carry to another one used by the same process)
    static bool attr_done[64] = {false};
    int dev = 0;
    TRX_CUDA(cudaGetDevice(&dev));
    if (dev < 0 || dev >= 64 || !attr_done[dev]) {
        TRX_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, C::SMEM_BYTES));
        if (dev >= 0 && dev < 64) attr_done[dev] = true;
    }
    cudaLaunchConfig_t cfg{};
    cfg.gridDim = dim3((unsigned)grid);
    cfg.blockDim = dim3(NUM_THREADS);
    cfg.dynamicSmemBytes = C::SMEM_BYTES;
    cfg.stream = st;
    cudaLaunchAttribute attr[1];
    attr[0].id = cudaLaunchAttributeClusterDimension;
    attr[0].val.clusterDim.x = PAIR ? 2 : 1;
    attr[0].val.clusterDim.y = 1;
    attr[0].val.clusterDim.z = 1;
    cfg.attrs = attr;
    cfg.numAttrs = 1;
    TRX_CUDA(cudaLaunchKernelEx(&cfg, kern, mq, mx, p));
    count_launch();
    return TRX_OK;
}

template <bool PAIR, int AR>
int launch_tiling(const UmmaArgs& a, int sm_count, cudaStream_t st) {
    using C = Cfg<PAIR, AR>;
    KParams p;
    p.nq = a.nq; p.n = a.n; p.KB = a.Kp / BK;
    p.MT = (int)((a.nq + C::TILE_M - 1) / C::TILE_M);
    p.NT = (int)((a.n + BN - 1) / BN);
    p.S = a.mode == MODE_SLOTMAX ? umma_num_slices(a.n, a.nq, sm_count, PAIR) : p.NT;
    // a batch that is a single query tile moves only the rows that exist (the caller allocates q16 for
    // round8(nq) rows): no TMA out-of-bounds fill.  Several tiles: full boxes, the tail tile zero-filled.
    const int64_t nq8 = (a.nq + 7) / 8 * 8;
    p.a_box = (!PAIR && p.MT == 1) ? (int)(nq8 < AR ? nq8 : AR) : BM;
    CUtensorMap mq, mx;
    TRX_TRY(make_map(&mq, a.q16, (!PAIR && p.MT == 1) ? nq8 : a.nq, a.Kp, p.a_box));
    TRX_TRY(make_map(&mx, a.x16, a.n, a.Kp, C::B_ROWS));
    p.out = a.out; p.out_ld = a.out_ld;
    p.thr = a.thr; p.cand_cnt = a.cand_cnt;
    p.log = a.log; p.log_cnt = a.log_cnt; p.log_cap = a.log_cap;
    p.slots = a.out;  // SLOTMAX reuses `out` as the [nq][S][32] slot buffer
    const int grid = umma_grid(a.nq, a.n, sm_count, PAIR, a.mode == MODE_SLOTMAX);
    switch (a.mode) {
        case MODE_STORE: return launch_mode<MODE_STORE, PAIR, AR>(mq, mx, p, grid, st);
        case MODE_THRESH: {
            TRX_TRY((launch_mode<MODE_THRESH, PAIR, AR>(mq, mx, p, grid, st)));
            const int nlogs = grid * 128;
            k2_scatter_kernel<<<(nlogs * 32 + 255) / 256, 256, 0, st>>>(a.log, a.log_cnt, nlogs, a.log_cap, a.cand,
                                                                         a.cand_cnt, a.cap);
            count_launch();
            TRX_CUDA(cudaGetLastError());
            return TRX_OK;
        }
        case MODE_SLOTMAX: return launch_mode<MODE_SLOTMAX, PAIR, AR>(mq, mx, p, grid, st);
    }
    set_error("k2: bad mode %d", a.mode);
    return TRX_EINVAL;
}

}  // namespace

int umma_init() {
    if (g_encode) return TRX_OK;
    void* fn = nullptr;
    cudaDriverEntryPointQueryResult qres;
    TRX_CUDA(cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &fn, cudaEnableDefault, &qres));
    if (qres != cudaDriverEntryPointSuccess || fn == nullptr) { set_error("cuTensorMapEncodeTiled not available"); return TRX_ECUDA; }
    g_encode = reinterpret_cast<PFN_cuTensorMapEncodeTiled_v12000>(fn);
    return TRX_OK;
}

// SLOTMAX slices per query tile: enough (query tile, slice) units to occupy every worker, at least 8.
int umma_num_slices(int64_t n, int64_t nq, int sm_count, bool pair) {
    const int tile_m = pair ? 2 * BM : BM;
    const int64_t MT = (nq + tile_m - 1) / tile_m;
    const int64_t NT = (n + BN - 1) / BN;
    const int64_t workers = pair ? sm_count / 2 : sm_count;
    int64_t S = (workers + MT - 1) / MT;
    if (S < 8) S = 8;
    if (S > 256) S = 256;
    return (int)(NT < S ? NT : S);
}

// CTAs a launch uses: persistent, one CTA (or CTA pair) per SM, never more workers than work units.
int umma_grid(int64_t nq, int64_t n, int sm_count, bool pair, bool slotmax) {
    const int tile_m = pair ? 2 * BM : BM;
    int64_t MT = (nq + tile_m - 1) / tile_m;
    int64_t NT = (n + BN - 1) / BN;
    int64_t units = MT * (slotmax ? umma_num_slices(n, nq, sm_count, pair) : NT);
    int64_t workers = pair ? sm_count / 2 : sm_count;
    if (units < workers) workers = units;
    return (int)(pair ? 2 * workers : workers);
}

int launch_umma(const UmmaArgs& a, int sm_count, cudaStream_t st) {
    if (a.nq <= 0 || a.n <= 0) return TRX_OK;
    TRX_TRY(umma_init());
    if (a.Kp % BK) { set_error("k2: Kp=%d not a multiple of %d", a.Kp, BK); return TRX_EINVAL; }
    if (a.pair) return launch_tiling<true, 128>(a, sm_count, st);
    return a.nq <= 32 ? launch_tiling<false, 32>(a, sm_count, st) : launch_tiling<false, 128>(a, sm_count, st);
}

}  // namespace trx
