// First Linear of the SOCIAL grid embedding, second-generation tcgen05 kernel ("TS" form: the
// A operand lives in tensor memory, optionally executed by a CTA pair).
//
//   hidden1[p, :] = relu(b1 + sum_{winning (cell, j) of p} W1[:, cell-slab] . lat_j)
//   (reference: GridBasedPooling.social + the first Linear of two_layer,
//    trajnetbaselines/lstm/gridbased_pooling.py:145-170,227-305,316-323)
//
// Orientation (the transpose of sparse_layer1_tc_kernel): D[p, col] with the pedestrians on the M
// side -- 128 TMEM lanes per CTA, one pedestrian per lane -- and the output columns on the N side.
// Per grid cell c one K = 16 product
//     D[p, col] += L_c[p, 0:16] . W_c[col, 0:16]
// where L_c[p, :] is the latent vector of p's winning neighbour in cell c, or zero.
//   * L_c is never materialised in shared memory: four builder warps (one per TMEM lane quarter, one
//     thread per pedestrian) look the winner up in a per-tile cell map and write the (hi, lo) bf16 rows
//     straight into a small TMEM ring with tcgen05.st; the MMA reads A from TMEM.  Shared-memory
//     bandwidth is spent on the weights only (the SS kernel re-read A and B for every pass and was
//     bound by it).
//   * The weight slabs W_c are the B operand.  They are stored pre-swizzled (SWIZZLE_32B image) in
//     global memory, so one plain bulk copy (cp.async.bulk) of any multiple of 16 rows lands the
//     tile the UMMA descriptor expects.
//   * kPair: two CTAs (one TPC) run tcgen05.mma.cta_group::2 with M = 256: each CTA owns 128
//     pedestrians and HALF of the weight columns of the tile, so every weight byte fetched from L2
//     serves 256 pedestrians and each SM's shared memory only sees half of B.
//   * Work decomposition: the (pedestrian tile, 32-column block) space is linearised and cut into
//     equal contiguous ranges, one per CTA (pair); a range is processed in rounds of at most two
//     contexts (tile, column range), so a range that straddles a tile boundary still keeps all SMs
//     equally loaded (5120 pedestrians x 1024 columns on 74 pairs: 8.65 blocks each).
//   * Precision: 3-pass bf16 (hi, lo) split, fp32 accumulation in TMEM, like the other kernels.
//
//   * Pipeline item = 4 grid cells (K = 64; 2 cells when a round spans two pedestrian tiles): one
//     barrier round trip builder -> MMA issuer -> tcgen05.commit per item, 3 items in flight.  With one
//     cell per item the loop was bound by that round trip (measured 1200 cycles per cell against a
//     415-cycle tensor floor).
//
// Warp roles (512 threads): 0 = weight producer, 1 = TMEM alloc + MMA issuer (leader CTA only),
// 4..7 = builders of context 0, 8..11 = builders of context 1; warps 4..15 run the epilogue (3 per
// TMEM lane quarter); all 16 warps build the cell maps at the start of a round.
#include <cuda.h>
#include <cuda_bf16.h>
#include <cstdio>
#include <cstdlib>
#include <vector>

#include "common.cuh"

namespace tb2 {

constexpr int kTsThreads = 512;
constexpr int kTsStages = 3;             // pipeline items in flight (TMEM ring of A tiles + smem ring of weight slabs)
constexpr int kTsKC = 4;                 // grid cells per pipeline item (2 when a round spans two pedestrian tiles)
constexpr int kTsMaxBlocks = 9;          // 32-column blocks per round (D = 288 TMEM columns)
constexpr int kTsDCols = kTsMaxBlocks * 32;
constexpr int kTsACol0 = kTsDCols;       // A ring: 192 columns = tiles x stages x cells-per-item x 16 (hi 8 | lo 8)
constexpr int kTsMaxCells = 256;
constexpr int kTsLatRows = 256;          // local latent table (rows of the scenes a 128-row tile touches) + 1

template <bool kPair> struct TsCfg {
    static constexpr int nC = kPair ? 2 : 1;
    static constexpr uint32_t cell_bytes = (uint32_t)(kTsDCols / nC) * 32u * 2u;    // (hi, lo) rows of 32 bytes
    static constexpr uint32_t stage_bytes = cell_bytes * (kPair ? kTsKC : 2);       // one CTA: 2 cells per item
};

__device__ __forceinline__ uint32_t ts_smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void ts_mbar_init(uint32_t bar, uint32_t count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(bar), "r"(count));
}
__device__ __forceinline__ void ts_mbar_expect_tx(uint32_t bar, uint32_t bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"(bytes) : "memory");
}
__device__ __forceinline__ void ts_mbar_arrive_cluster(uint32_t cluster_addr) {
    asm volatile("mbarrier.arrive.release.cluster.shared::cluster.b64 _, [%0];" ::"r"(cluster_addr) : "memory");
}
__device__ __forceinline__ void ts_mbar_wait(uint32_t bar, uint32_t parity) {
    asm volatile(
        "{\n"
        ".reg .pred p;\n"
        "TS_WAIT_LOOP:\n"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n"
        "@p bra TS_WAIT_DONE;\n"
        "bra TS_WAIT_LOOP;\n"
        "TS_WAIT_DONE:\n"
        "}\n" ::"r"(bar), "r"(parity) : "memory");
}
// the leader's "A and B of this cell are ready" barrier is also arrived on from the peer CTA: acquire at cluster scope
__device__ __forceinline__ void ts_mbar_wait_cluster(uint32_t bar, uint32_t parity) {
    asm volatile(
        "{\n"
        ".reg .pred p;\n"
        "TSC_WAIT_LOOP:\n"
        "mbarrier.try_wait.parity.acquire.cluster.shared::cta.b64 p, [%0], %1;\n"
        "@p bra TSC_WAIT_DONE;\n"
        "bra TSC_WAIT_LOOP;\n"
        "TSC_WAIT_DONE:\n"
        "}\n" ::"r"(bar), "r"(parity) : "memory");
}
__device__ __forceinline__ void ts_bulk_load(uint32_t dst, const void* src, uint32_t bytes, uint32_t bar) {
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
                 ::"r"(dst), "l"(src), "r"(bytes), "r"(bar) : "memory");
}
// K-major operand with 32-byte rows, SWIZZLE_32B: 8-row atoms of 256 bytes, SBO = 256
__device__ __forceinline__ uint64_t ts_umma_desc(uint32_t smem_addr) {
    uint64_t d = 0;
    d |= (uint64_t)((smem_addr & 0x3FFFF) >> 4);
    d |= (uint64_t)(256 >> 4) << 32;
    d |= (uint64_t)1 << 46;
    d |= (uint64_t)6 << 61;          // SWIZZLE_32B
    return d;
}
template <bool kPair>
__device__ __forceinline__ void ts_umma(uint32_t tmem_d, uint32_t tmem_a, uint64_t b, uint32_t idesc, uint32_t acc) {
    if (kPair)
        asm volatile(
            "{\n"
            ".reg .pred p;\n"
            "setp.ne.b32 p, %4, 0;\n"
            "tcgen05.mma.cta_group::2.kind::f16 [%0], [%1], %2, %3, p;\n"
            "}\n" ::"r"(tmem_d), "r"(tmem_a), "l"(b), "r"(idesc), "r"(acc) : "memory");
    else
        asm volatile(
            "{\n"
            ".reg .pred p;\n"
            "setp.ne.b32 p, %4, 0;\n"
            "tcgen05.mma.cta_group::1.kind::f16 [%0], [%1], %2, %3, p;\n"
            "}\n" ::"r"(tmem_d), "r"(tmem_a), "l"(b), "r"(idesc), "r"(acc) : "memory");
}
// completion of all earlier tcgen05.mma of this thread -> arrive on `bar` (pair: in both CTAs)
template <bool kPair>
__device__ __forceinline__ void ts_commit(uint32_t bar) {
    if (kPair) {
        const uint16_t mask = 3;
        asm volatile("tcgen05.commit.cta_group::2.mbarrier::arrive::one.shared::cluster.multicast::cluster.b64 [%0], %1;"
                     ::"r"(bar), "h"(mask) : "memory");
    } else {
        asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(bar) : "memory");
    }
}
__device__ __forceinline__ void ts_tmem_st8(uint32_t taddr, const uint4& a, const uint4& b) {
    asm volatile("tcgen05.st.sync.aligned.32x32b.x8.b32 [%0], {%1, %2, %3, %4, %5, %6, %7, %8};"
                 ::"r"(taddr), "r"(a.x), "r"(a.y), "r"(a.z), "r"(a.w), "r"(b.x), "r"(b.y), "r"(b.z), "r"(b.w) : "memory");
}
__device__ __forceinline__ void ts_tmem_ld32(uint32_t taddr, uint32_t (&r)[32]) {
    asm volatile(
        "tcgen05.ld.sync.aligned.32x32b.x32.b32 "
        "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, "
        "%16, %17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31}, [%32];"
        : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]),
          "=r"(r[8]), "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15]),
          "=r"(r[16]), "=r"(r[17]), "=r"(r[18]), "=r"(r[19]), "=r"(r[20]), "=r"(r[21]), "=r"(r[22]), "=r"(r[23]),
          "=r"(r[24]), "=r"(r[25]), "=r"(r[26]), "=r"(r[27]), "=r"(r[28]), "=r"(r[29]), "=r"(r[30]), "=r"(r[31])
        : "r"(taddr));
}

struct TsParams {
    const int* scene_off;
    const int* row_scene;
    const int* win_count;
    const uint32_t* win_ent;
    const float* lat;            // [M, 16] fp32
    const float* benc;           // [16]
    const float* base;           // [OUT]
    const unsigned char* w_hi;   // bf16 [cells][OUT][16], 16-byte chunks swizzled by ((col >> 2) & 1)
    const unsigned char* w_lo;
    float* out;                  // fp32 [M, OUT] or null
    __nv_bfloat16* out_hi;       // bf16 split [M, OUT] or null
    __nv_bfloat16* out_lo;
    int M, OUT, cells, nm1, units;
    float constant;
    long long* dbg;              // optional [units * nC, 8] cycle counters (TB2_L1_DEBUG=1)
};

struct TsCtx { int tile, col0, ncols; };

// next round of the contiguous block range [cur, b1): at most two contexts, <= 8 blocks each, <= 9 together
__device__ __forceinline__ bool ts_next_round(long long& cur, long long b1, int nblk_tile, TsCtx (&c)[2]) {
    if (cur >= b1) return false;
    const int tile0 = (int)(cur / nblk_tile);
    long long e0 = (long long)(tile0 + 1) * nblk_tile;
    if (e0 > b1) e0 = b1;
    if (e0 > cur + 8) e0 = cur + 8;
    const int n0 = (int)(e0 - cur);
    c[0].tile = tile0; c[0].col0 = (int)(cur - (long long)tile0 * nblk_tile) * 32; c[0].ncols = n0 * 32;
    c[1].tile = tile0; c[1].col0 = 0; c[1].ncols = 0;
    cur = e0;
    if (cur < b1) {
        const int tile1 = (int)(cur / nblk_tile);
        const int room = kTsMaxBlocks - n0 < 8 ? kTsMaxBlocks - n0 : 8;
        long long e1 = (long long)(tile1 + 1) * nblk_tile;
        if (e1 > b1) e1 = b1;
        if (e1 > cur + room) e1 = cur + room;
        const int n1 = (int)(e1 - cur);
        c[1].tile = tile1; c[1].col0 = (int)(cur - (long long)tile1 * nblk_tile) * 32; c[1].ncols = n1 * 32;
        cur = e1;
    }
    return true;
}

template <bool kPair>
__global__ void __launch_bounds__(kTsThreads, 1) sparse_layer1_ts_kernel(TsParams p) {
    using Cfg = TsCfg<kPair>;
    constexpr int nC = Cfg::nC;
    constexpr int S = kTsStages;
    extern __shared__ __align__(1024) unsigned char smem_ts[];
    __shared__ __align__(8) uint64_t full_bar[S];        // leader: A (both CTAs) + B (both CTAs) of an item ready
    __shared__ __align__(8) uint64_t full_b[S];          // local weight slabs landed
    __shared__ __align__(8) uint64_t empty_bar[S];       // MMAs that read the stage (TMEM A tiles + smem slabs) done
    __shared__ __align__(8) uint64_t acc_full_bar;
    __shared__ uint32_t tmem_base_slot;

    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    uint32_t rank = 0;
    if (kPair) asm volatile("mov.u32 %0, %%cluster_ctarank;" : "=r"(rank));
    const int unit = kPair ? (int)(blockIdx.x >> 1) : (int)blockIdx.x;
    long long* dbg = p.dbg ? p.dbg + (size_t)blockIdx.x * 8 : nullptr;
    const long long t_begin = clock64();

    const uint32_t ring = (ts_smem_u32(smem_ts) + 1023u) & ~1023u;
    unsigned char* ring_ptr = smem_ts + (ring - ts_smem_u32(smem_ts));
    unsigned char* tail = ring_ptr + (size_t)S * Cfg::stage_bytes;
    unsigned char* cellmap[2];
    cellmap[0] = tail;
    cellmap[1] = tail + (size_t)kTsMaxCells * 128;
    uint4* latH[2];
    uint4* latL[2];
    latH[0] = reinterpret_cast<uint4*>(tail + (size_t)2 * kTsMaxCells * 128);
    latL[0] = latH[0] + kTsLatRows * 2;
    latH[1] = latL[0] + kTsLatRows * 2;
    latL[1] = latH[1] + kTsLatRows * 2;

    // ---- prologue: barriers, TMEM ---------------------------------------------------------------
    if (warp == 0 && lane == 0) {
        for (int s = 0; s < S; ++s) {
            ts_mbar_init(ts_smem_u32(&full_bar[s]), 8 * nC);      // 8 builder warps per CTA
            ts_mbar_init(ts_smem_u32(&full_b[s]), 1);
            ts_mbar_init(ts_smem_u32(&empty_bar[s]), 1);
        }
        ts_mbar_init(ts_smem_u32(&acc_full_bar), 1);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    if (warp == 1) {
        uint32_t ncols = 512;
        if (kPair) {
            asm volatile("tcgen05.alloc.cta_group::2.sync.aligned.shared::cta.b32 [%0], %1;"
                         ::"r"(ts_smem_u32(&tmem_base_slot)), "r"(ncols) : "memory");
            asm volatile("tcgen05.relinquish_alloc_permit.cta_group::2.sync.aligned;" ::: "memory");
        } else {
            asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;"
                         ::"r"(ts_smem_u32(&tmem_base_slot)), "r"(ncols) : "memory");
            asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
        }
    }
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    if (kPair) {      // the peer's barriers must be initialised before the first remote arrive / multicast commit
        asm volatile("barrier.cluster.arrive.release.aligned;" ::: "memory");
        asm volatile("barrier.cluster.wait.acquire.aligned;" ::: "memory");
    }
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    const uint32_t tmem_base = tmem_base_slot;
    grid_dep_wait();          // winners / latent vectors come from pool_prepare
    grid_dep_launch();

    // leader's full barrier as seen from this CTA (cluster address space)
    uint32_t full_remote[S];
#pragma unroll
    for (int s = 0; s < S; ++s) {
        uint32_t local = ts_smem_u32(&full_bar[s]);
        if (kPair) asm volatile("mapa.shared::cluster.u32 %0, %1, %2;" : "=r"(full_remote[s]) : "r"(local), "r"(0));
        else full_remote[s] = local;
    }

    const int R = 128 * nC;                               // pedestrian rows per tile
    const int tiles = (p.M + R - 1) / R;
    const int nblk_tile = p.OUT / 32;
    const long long total_blk = (long long)tiles * nblk_tile;
    long long cur = (long long)unit * total_blk / p.units;
    const long long b1 = (long long)(unit + 1) * total_blk / p.units;

    uint32_t git = 0;                // global item counter (pipeline phases run on across rounds)
    uint32_t round_idx = 0;
    long long t_setup_sum = 0, t_loop_sum = 0, t_epi_sum = 0, wait_full = 0;
    TsCtx ctx[2];
    while (ts_next_round(cur, b1, nblk_tile, ctx)) {
        const long long t_r0 = clock64();
        // ---- round setup (all warps): cell maps + split latent tables of the one / two tiles ------
        const bool two = ctx[1].ncols > 0;
        const bool alias = two && ctx[1].tile == ctx[0].tile;       // same tile: context 1 reuses context 0's A tiles
        const bool distinct2 = two && !alias;
        const int KC = (distinct2 || !kPair) ? 2 : kTsKC;           // cells per pipeline item
        int rbase[2], nrows[2], lbase[2], nlat[2];
#pragma unroll
        for (int x = 0; x < 2; ++x) {
            rbase[x] = ctx[x].tile * R + (int)rank * 128;
            int nr = p.M - rbase[x];
            nr = nr < 0 ? 0 : (nr > 128 ? 128 : nr);
            if (ctx[x].ncols == 0 || (x == 1 && alias)) nr = 0;
            nrows[x] = nr;
            lbase[x] = 0; nlat[x] = 0;
            if (nr > 0) {
                const int s_lo = p.row_scene[rbase[x]], s_hi = p.row_scene[rbase[x] + nr - 1];
                lbase[x] = p.scene_off[s_lo];
                nlat[x] = p.scene_off[s_hi + 1] - lbase[x];
            }
        }
#pragma unroll
        for (int x = 0; x < 2; ++x) {
            if (ctx[x].ncols == 0 || (x == 1 && alias)) continue;
            uint4* cm = reinterpret_cast<uint4*>(cellmap[x]);
            for (int i = tid; i < p.cells * 128 / 16; i += kTsThreads) cm[i] = make_uint4(~0u, ~0u, ~0u, ~0u);
            for (int idx = tid; idx < (nlat[x] + 1) * 8; idx += kTsThreads) {       // 2 values per thread
                const int r = idx >> 3, k = (idx & 7) * 2;
                const float* src = r < nlat[x] ? p.lat + (size_t)(lbase[x] + r) * 16 : p.benc;
                const float v0 = src[k] - p.constant, v1 = src[k + 1] - p.constant;
                const __nv_bfloat16 h0 = __float2bfloat16_rn(v0), h1 = __float2bfloat16_rn(v1);
                const __nv_bfloat16 l0 = __float2bfloat16_rn(v0 - __bfloat162float(h0));
                const __nv_bfloat16 l1 = __float2bfloat16_rn(v1 - __bfloat162float(h1));
                reinterpret_cast<uint32_t*>(latH[x])[idx] = (uint32_t)__bfloat16_as_ushort(h0) | ((uint32_t)__bfloat16_as_ushort(h1) << 16);
                reinterpret_cast<uint32_t*>(latL[x])[idx] = (uint32_t)__bfloat16_as_ushort(l0) | ((uint32_t)__bfloat16_as_ushort(l1) << 16);
            }
        }
        __syncthreads();
#pragma unroll
        for (int x = 0; x < 2; ++x) {
            const int total = nrows[x] * p.nm1;
            for (int idx = tid; idx < total; idx += kTsThreads) {
                const int r = idx / p.nm1, k = idx - r * p.nm1;
                const int row = rbase[x] + r;
                if (k < p.win_count[row]) {
                    const uint32_t e = p.win_ent[(size_t)row * p.nm1 + k];
                    const int j = (int)(e & 0xffffu);
                    const int li = j == 0xffff ? nlat[x] : p.scene_off[p.row_scene[row]] + j - lbase[x];
                    cellmap[x][(e >> 16) * 128 + r] = (unsigned char)li;
                }
            }
        }
        __syncthreads();
        const long long t_r1 = clock64();
        t_setup_sum += t_r1 - t_r0;

        const int n_items = (p.cells + KC - 1) / KC;
        // rows of the weight slab each CTA holds per context, byte offsets inside one cell of a stage
        const uint32_t nr0 = (uint32_t)ctx[0].ncols / nC, nr1 = (uint32_t)ctx[1].ncols / nC;
        const uint32_t cell_bytes = 2u * 32u * (nr0 + nr1);
        const uint32_t off_hi[2] = {0u, 2u * nr0 * 32u};
        const uint32_t off_lo[2] = {nr0 * 32u, 2u * nr0 * 32u + nr1 * 32u};
        // TMEM A tiles: [tile slot][stage][cell of the item][hi 8 | lo 8 columns]
        const uint32_t a_slot1 = distinct2 ? (uint32_t)(S * KC * 16) : 0u;

        if (warp == 0) {
            // ===== weight producer: one bulk copy per (cell, context, hi / lo) =====
            if (lane == 0) {
                for (int it = 0; it < n_items; ++it) {
                    const uint32_t g = git + (uint32_t)it;
                    const uint32_t s = g % S, ph = (g / S) & 1u;
                    const int c0 = it * KC;
                    const int kcn = p.cells - c0 < KC ? p.cells - c0 : KC;
                    ts_mbar_wait(ts_smem_u32(&empty_bar[s]), ph ^ 1u);
                    const uint32_t bar = ts_smem_u32(&full_b[s]);
                    const uint32_t st = ring + s * Cfg::stage_bytes;
                    ts_mbar_expect_tx(bar, cell_bytes * (uint32_t)kcn);
                    for (int kc = 0; kc < kcn; ++kc) {
#pragma unroll
                        for (int x = 0; x < 2; ++x) {
                            const uint32_t nr = x == 0 ? nr0 : nr1;
                            if (nr == 0) continue;
                            const size_t src = ((size_t)(c0 + kc) * p.OUT + ctx[x].col0 + rank * nr) * 32;
                            ts_bulk_load(st + kc * cell_bytes + off_hi[x], p.w_hi + src, nr * 32u, bar);
                            ts_bulk_load(st + kc * cell_bytes + off_lo[x], p.w_lo + src, nr * 32u, bar);
                        }
                    }
                }
            }
            __syncwarp();
        } else if (warp == 1) {
            // ===== MMA issuer (leader CTA of a pair) =====
            if (lane == 0 && rank == 0) {
                uint32_t idesc[2];
#pragma unroll
                for (int x = 0; x < 2; ++x)
                    idesc[x] = (1u << 4) | (1u << 7) | (1u << 10) | ((uint32_t)(ctx[x].ncols >> 3) << 17) | ((uint32_t)(R >> 4) << 24);
                const uint32_t d_col[2] = {0u, (uint32_t)ctx[0].ncols};
                for (int it = 0; it < n_items; ++it) {
                    const uint32_t g = git + (uint32_t)it;
                    const uint32_t s = g % S, ph = (g / S) & 1u;
                    const int c0 = it * KC;
                    const int kcn = p.cells - c0 < KC ? p.cells - c0 : KC;
                    const long long t0 = clock64();
                    ts_mbar_wait_cluster(ts_smem_u32(&full_bar[s]), ph);
                    wait_full += clock64() - t0;
                    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
                    const uint32_t st = ring + s * Cfg::stage_bytes;
                    for (int kc = 0; kc < kcn; ++kc) {
#pragma unroll
                        for (int x = 0; x < 2; ++x) {
                            if (ctx[x].ncols == 0) continue;
                            const uint32_t a_hi = tmem_base + (uint32_t)kTsACol0 + (x == 1 ? a_slot1 : 0u) +
                                                  (uint32_t)((s * KC + kc) * 16);
                            const uint32_t a_lo = a_hi + 8u;
                            const uint64_t b_hi = ts_umma_desc(st + kc * cell_bytes + off_hi[x]);
                            const uint64_t b_lo = ts_umma_desc(st + kc * cell_bytes + off_lo[x]);
                            const uint32_t d = tmem_base + d_col[x];
                            ts_umma<kPair>(d, a_hi, b_hi, idesc[x], (it > 0 || kc > 0) ? 1u : 0u);
                            ts_umma<kPair>(d, a_lo, b_hi, idesc[x], 1u);
                            ts_umma<kPair>(d, a_hi, b_lo, idesc[x], 1u);
                        }
                    }
                    ts_commit<kPair>(ts_smem_u32(&empty_bar[s]));
                }
                ts_commit<kPair>(ts_smem_u32(&acc_full_bar));
            }
            __syncwarp();
        } else if (warp >= 4 && warp < 12) {
            // ===== A builders: thread = pedestrian row (TMEM lane), one tcgen05.st pair per cell =====
            const int x = (warp - 4) >> 2;                 // context
            const int q = warp & 3;                        // TMEM lane quarter of this warp
            const int r = q * 32 + lane;
            const bool active = ctx[x].ncols > 0 && !(x == 1 && alias);
            const unsigned char* cm = cellmap[x] + r;
            const uint4* lh = latH[x];
            const uint4* ll = latL[x];
            const uint32_t a_base = tmem_base + ((uint32_t)(q * 32) << 16) + (uint32_t)kTsACol0 + (x == 1 ? a_slot1 : 0u);
            for (int it = 0; it < n_items; ++it) {
                const uint32_t g = git + (uint32_t)it;
                const uint32_t s = g % S, ph = (g / S) & 1u;
                const int c0 = it * KC;
                const int kcn = p.cells - c0 < KC ? p.cells - c0 : KC;
                ts_mbar_wait(ts_smem_u32(&empty_bar[s]), ph ^ 1u);
                if (active) {
                    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
                    for (int kc = 0; kc < kcn; ++kc) {
                        const uint32_t li = cm[(size_t)(c0 + kc) * 128];
                        uint4 h0 = make_uint4(0u, 0u, 0u, 0u), h1 = h0, l0 = h0, l1 = h0;
                        if (li != 0xffu) {
                            h0 = lh[li * 2]; h1 = lh[li * 2 + 1];
                            l0 = ll[li * 2]; l1 = ll[li * 2 + 1];
                        }
                        const uint32_t a = a_base + (uint32_t)((s * KC + kc) * 16);
                        ts_tmem_st8(a, h0, h1);
                        ts_tmem_st8(a + 8u, l0, l1);
                    }
                    asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory");
                }
                if (x == 0)       // the weights of this item have landed in THIS CTA (chained into the arrive below)
                    ts_mbar_wait(ts_smem_u32(&full_b[s]), ph);
                asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
                __syncwarp();
                if (lane == 0) ts_mbar_arrive_cluster(full_remote[s]);
            }
        }
        if (warp >= 4) {
            // ===== epilogue (12 warps, 3 per TMEM lane quarter): thread = pedestrian row, 32 columns per tcgen05.ld =====
            const int q = warp & 3, sub = (warp - 4) >> 2;
            ts_mbar_wait(ts_smem_u32(&acc_full_bar), round_idx & 1u);
            asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
            const long long t_e0 = clock64();
            if (warp == 12 && lane == 0) t_loop_sum += t_e0 - t_r1;
            const uint32_t trow = tmem_base + ((uint32_t)(q * 32) << 16);
            const int nchunk0 = ctx[0].ncols >> 5, nchunk = (ctx[0].ncols + ctx[1].ncols) >> 5;
            for (int ch = sub; ch < nchunk; ch += 3) {
                const int x = ch >= nchunk0 ? 1 : 0;
                const int c0 = (x ? ch - nchunk0 : ch) * 32;
                const int tile = x ? ctx[1].tile : ctx[0].tile;
                const int col = (x ? ctx[1].col0 : ctx[0].col0) + c0;
                const int row = tile * R + (int)rank * 128 + q * 32 + lane;
                uint32_t v[32];
                ts_tmem_ld32(trow + (uint32_t)(ch * 32), v);
                asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
                if (row < p.M) {
                    const size_t o = (size_t)row * p.OUT + col;
                    if (p.out_hi) {
                        uint32_t ph[16], pl[16];
#pragma unroll
                        for (int i = 0; i < 16; ++i) {
                            const float x0 = fmaxf(__uint_as_float(v[2 * i]) + __ldg(p.base + col + 2 * i), 0.f);
                            const float x1 = fmaxf(__uint_as_float(v[2 * i + 1]) + __ldg(p.base + col + 2 * i + 1), 0.f);
                            const __nv_bfloat16 h0 = __float2bfloat16_rn(x0), h1 = __float2bfloat16_rn(x1);
                            const __nv_bfloat16 l0 = __float2bfloat16_rn(x0 - __bfloat162float(h0));
                            const __nv_bfloat16 l1 = __float2bfloat16_rn(x1 - __bfloat162float(h1));
                            ph[i] = (uint32_t)__bfloat16_as_ushort(h0) | ((uint32_t)__bfloat16_as_ushort(h1) << 16);
                            pl[i] = (uint32_t)__bfloat16_as_ushort(l0) | ((uint32_t)__bfloat16_as_ushort(l1) << 16);
                        }
                        uint4* dh = reinterpret_cast<uint4*>(p.out_hi + o);
                        uint4* dl = reinterpret_cast<uint4*>(p.out_lo + o);
#pragma unroll
                        for (int i = 0; i < 4; ++i) {
                            dh[i] = make_uint4(ph[4 * i], ph[4 * i + 1], ph[4 * i + 2], ph[4 * i + 3]);
                            dl[i] = make_uint4(pl[4 * i], pl[4 * i + 1], pl[4 * i + 2], pl[4 * i + 3]);
                        }
                    } else {
                        float4* d4 = reinterpret_cast<float4*>(p.out + o);
#pragma unroll
                        for (int i = 0; i < 8; ++i) {
                            float4 y;
                            y.x = fmaxf(__uint_as_float(v[4 * i]) + __ldg(p.base + col + 4 * i), 0.f);
                            y.y = fmaxf(__uint_as_float(v[4 * i + 1]) + __ldg(p.base + col + 4 * i + 1), 0.f);
                            y.z = fmaxf(__uint_as_float(v[4 * i + 2]) + __ldg(p.base + col + 4 * i + 2), 0.f);
                            y.w = fmaxf(__uint_as_float(v[4 * i + 3]) + __ldg(p.base + col + 4 * i + 3), 0.f);
                            d4[i] = y;
                        }
                    }
                }
            }
            if (warp == 12 && lane == 0) t_epi_sum += clock64() - t_e0;
        }
        git += (uint32_t)n_items;
        ++round_idx;
        // the next round overwrites cell maps, the TMEM accumulators and (pair) the peer's accumulators
        asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
        __syncthreads();
        if (kPair) {
            asm volatile("barrier.cluster.arrive.release.aligned;" ::: "memory");
            asm volatile("barrier.cluster.wait.acquire.aligned;" ::: "memory");
        }
        asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    }
    if (dbg) {
        if (tid == 0) {
            dbg[0] = t_setup_sum; dbg[4] = round_idx; dbg[6] = clock64() - t_begin;
            dbg[5] = (long long)ctx[0].ncols | ((long long)ctx[1].ncols << 12) | ((long long)(ctx[1].ncols > 0 && ctx[1].tile != ctx[0].tile) << 24);
        }
        if (warp == 1 && lane == 0) dbg[2] = wait_full;
        if (warp == 12 && lane == 0) { dbg[1] = t_loop_sum; dbg[3] = t_epi_sum; }
    }
    if (warp == 1) {
        uint32_t ncols = 512;
        if (kPair) asm volatile("tcgen05.dealloc.cta_group::2.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(ncols) : "memory");
        else asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(ncols) : "memory");
    }
}

// ------------------------------------------------------------------------------------------
// host side
// ------------------------------------------------------------------------------------------
template <bool kPair> static size_t ts_smem_bytes() {
    return 1024 + (size_t)kTsStages * TsCfg<kPair>::stage_bytes + (size_t)2 * kTsMaxCells * 128 +
           (size_t)4 * kTsLatRows * 32 + 64;
}

// mode: 1 = single CTA (cta_group::1), 2 = CTA pair (cta_group::2)
bool sparse_ts_supported(const tb2_lstm* m, const tb2_layout* l) {
    if (m->cfg.pool_type != TB2_POOL_SOCIAL || m->C != 16 || m->Wt1_sw_hi == nullptr) return false;
    if (m->mlp_dims[1] % 32 != 0 || m->cells > kTsMaxCells) return false;
    // local latent table of a 128-row tile: the scenes it touches, <= 128 + 2 (n_max - 1) rows, + 1
    return 128 + 2 * (l->n_max - 1) + 1 <= kTsLatRows - 1;
}

template <bool kPair>
static int launch_sparse_ts_t(const tb2_lstm* m, const tb2_layout* l, Workspace* ws, float* out, void* out_hi,
                              void* out_lo, cudaStream_t st) {
    constexpr int nC = kPair ? 2 : 1;
    static int sm_count = 0;
    if (sm_count == 0) {
        int dev = 0;
        TB2_CHECK_CUDA(cudaGetDevice(&dev));
        TB2_CHECK_CUDA(cudaDeviceGetAttribute(&sm_count, cudaDevAttrMultiProcessorCount, dev));
    }
    const int nm1 = l->n_max > 1 ? l->n_max - 1 : 1;
    const int d1 = m->mlp_dims[1];
    const int R = 128 * nC;
    const long long tiles = (l->M + R - 1) / R;
    const long long total_blk = tiles * (d1 / 32);
    int units = sm_count / nC;
    {
        const char* e = getenv("TB2_TS_UNITS");      // debug knob
        if (e && atoi(e) > 0) units = atoi(e);
    }
    if ((long long)units > total_blk) units = (int)total_blk;
    TsParams p;
    p.scene_off = l->scene_off;
    p.row_scene = l->row_scene;
    p.win_count = ws->win_count;
    p.win_ent = ws->win_ent;
    p.lat = ws->lat;
    p.benc = m->benc;
    p.base = m->base1;
    p.w_hi = (const unsigned char*)m->Wt1_sw_hi;
    p.w_lo = (const unsigned char*)m->Wt1_sw_lo;
    p.out = out;
    p.out_hi = (__nv_bfloat16*)out_hi;
    p.out_lo = (__nv_bfloat16*)out_lo;
    p.M = l->M;
    p.OUT = d1;
    p.cells = m->cells;
    p.nm1 = nm1;
    p.units = units;
    p.constant = m->cfg.constant;
    p.dbg = nullptr;
    static long long* dbg_buf = nullptr;
    static int dbg_calls = 0;
    const int n_cta = units * nC;
    {
        const char* e = getenv("TB2_L1_DEBUG");
        if (e && e[0] == '1') {
            if (!dbg_buf) { cudaMalloc(&dbg_buf, (size_t)1024 * 8 * sizeof(long long)); cudaMemset(dbg_buf, 0, (size_t)1024 * 8 * sizeof(long long)); }
            if (n_cta <= 1024) p.dbg = dbg_buf;
        }
    }
    const size_t smem = ts_smem_bytes<kPair>();
    static DynSmemConfig configured;
    TB2_CHECK_CUDA(configured.ensure(sparse_layer1_ts_kernel<kPair>, smem));
    {
        KernelTimer kt(kPair ? "sparse_layer1_ts2" : "sparse_layer1_ts1", st);
        cudaLaunchConfig_t cfg = {};
        cfg.gridDim = dim3(n_cta);
        cfg.blockDim = dim3(kTsThreads);
        cfg.dynamicSmemBytes = smem;
        cfg.stream = st;
        cudaLaunchAttribute attr[2];
        int na = 0;
        if (pdl_enabled()) {
            attr[na].id = cudaLaunchAttributeProgrammaticStreamSerialization;
            attr[na].val.programmaticStreamSerializationAllowed = 1;
            ++na;
        }
        if (kPair) {
            attr[na].id = cudaLaunchAttributeClusterDimension;
            attr[na].val.clusterDim.x = 2;
            attr[na].val.clusterDim.y = 1;
            attr[na].val.clusterDim.z = 1;
            ++na;
        }
        cfg.attrs = attr;
        cfg.numAttrs = na;
        TB2_CHECK_CUDA(cudaLaunchKernelEx(&cfg, sparse_layer1_ts_kernel<kPair>, p));
    }
    TB2_LAUNCH_CHECK();
    if (p.dbg && ++dbg_calls == 60) {
        std::vector<long long> h((size_t)n_cta * 8);
        cudaStreamSynchronize(st);
        cudaMemcpy(h.data(), dbg_buf, h.size() * sizeof(long long), cudaMemcpyDeviceToHost);
        double a[7] = {0, 0, 0, 0, 0, 0, 0};
        double mx = 0;
        for (int c = 0; c < n_cta; ++c) {
            for (int k = 0; k < 7; ++k) a[k] += (double)h[(size_t)c * 8 + k] / n_cta;
            if ((double)h[(size_t)c * 8 + 6] > mx) mx = (double)h[(size_t)c * 8 + 6];
        }
        if (getenv("TB2_L1_DEBUG_ALL")) {
            for (int c = 0; c < n_cta; c += nC) {
                const long long* d = &h[(size_t)c * 8];
                fprintf(stderr, "  unit %3d  N0 %3d N1 %3d two-tiles %d | setup %6lld loop %7lld wait %7lld epi %6lld total %7lld\n",
                        c / nC, (int)(d[5] & 0xfff), (int)((d[5] >> 12) & 0xfff), (int)(d[5] >> 24), d[0], d[1], d[2], d[3], d[6]);
            }
        }
        fprintf(stderr, "[tb2 sparse_ts%d debug] per-CTA cycles: setup %.0f | cells loop until accumulators ready %.0f | "
                        "MMA thread waiting for A/B (leader CTAs, averaged over all) %.0f | epilogue %.0f | rounds %.1f | total %.0f (max %.0f)\n",
                nC, a[0], a[1], a[2], a[3], a[4], a[6], mx);
    }
    return TB2_OK;
}

int launch_sparse_ts(const tb2_lstm* m, const tb2_layout* l, int mode, Workspace* ws, float* out, void* out_hi,
                     void* out_lo, cudaStream_t st) {
    if (mode == 2) return launch_sparse_ts_t<true>(m, l, ws, out, out_hi, out_lo, st);
    return launch_sparse_ts_t<false>(m, l, ws, out, out_hi, out_lo, st);
}

// weight repack: W1[o][c * cells + cell] -> (hi, lo)[cell][o][16] bf16 with the two 16-byte halves of a
// row exchanged where ((o >> 2) & 1): the SWIZZLE_32B shared-memory image of any slab whose first
// row is a multiple of 8, so a plain bulk copy produces the tile the UMMA descriptor expects
__global__ void repack_layer1_sw_kernel(const float* __restrict__ W1, __nv_bfloat16* __restrict__ hi,
                                        __nv_bfloat16* __restrict__ lo, int OUT, int cells) {
    size_t total = (size_t)cells * OUT * 16;
    for (size_t idx = (size_t)blockIdx.x * blockDim.x + threadIdx.x; idx < total;
         idx += (size_t)gridDim.x * blockDim.x) {
        const int c = (int)(idx & 15);
        const size_t co = idx >> 4;
        const int o = (int)(co % OUT), cell = (int)(co / OUT);
        const float v = W1[(size_t)o * 16 * cells + (size_t)c * cells + cell];
        const __nv_bfloat16 h = __float2bfloat16_rn(v);
        const int chunk = (c >> 3) ^ ((o >> 2) & 1);
        const size_t dst = (co << 4) + (size_t)(chunk * 8 + (c & 7));
        hi[dst] = h;
        lo[dst] = __float2bfloat16_rn(v - __bfloat162float(h));
    }
}

int launch_repack_layer1_sw(const float* W1, void* hi, void* lo, int OUT, int cells, cudaStream_t st) {
    repack_layer1_sw_kernel<<<1024, 256, 0, st>>>(W1, (__nv_bfloat16*)hi, (__nv_bfloat16*)lo, OUT, cells);
    TB2_LAUNCH_CHECK();
    return TB2_OK;
}

}  // namespace tb2
