// conv_tc4: Conv2DReflect 3x3 as an implicit GEMM with ALIGNED on-chip tap reuse.
//
// v2 (conv_tc.cu) re-fetches the 128x64 activation tile from L2 for each of the 9 filter taps; on the
// 64-channel layers it is L2->SM operand-bandwidth bound (442 KB per 128x64 output tile at ~11.6 TB/s
// aggregate = the measured 249 TFLOP/s).  v3 (conv_tc3.cu) staged ONE patch and read all 9 taps through
// descriptors shifted by whole 128-byte rows -- functionally right, but UMMA operand fetch from a start
// address that is not on an 8-row (1024 B) swizzle atom runs ~3x slower.  v4 keeps every descriptor aligned:
//
//   tile  = 8 rows x 16 columns of output pixels (M = 128);
//   A     : per 64-channel slice THREE patches, one per horizontal tap kx: box [10 rows][16 px][64 ch] of the
//           reflect-padded plane starting at column x0+kx (5-D TMA, 20 KB per plane).  The three vertical taps
//           read the same patch at start + ky*16 rows = ky*2048 B: a multiple of the swizzle atom.
//           A traffic per tile: 3 x 40 KB = 120 KB instead of 9 x 32 KB = 288 KB;
//   B     : as v3 -- optional cluster of 2 CTAs on two M tiles of the same cout tile, each CTA loads half of every
//           weight tile and TMA-multicasts it to both (.multicast::cluster), stages released by a multicast commit.
//
// Everything else is v2: persistent CTAs, 4-k-iteration accumulation chunks in a TMEM ring drained into
// registers with round-to-nearest adds, split-fp16 x3 products, coalesced epilogue through shared memory.
#include "common.cuh"

namespace wctb {

struct Conv4Params {
    int N, H, W, Cin, Cout, Hp, Wp;
    long long P;
    int tiles_x, tiles_y, m_tiles, n_tiles;
    int na_stages;       // A ring depth (one stage = one kx patch, both planes)
    int nb_stages;       // B ring depth
    int cluster;         // 1 or 2
    long long* trace;    // optional [4][256] clock64 samples of CTA 0 (wctb200_debug_conv4_trace)
    int dbg;             // timing probes only (results are garbage): 1 = skip the MMAs, 2 = skip the A loads, 4 = skip the B loads
    int flags;
    const float* bias;
    __half* out;
    unsigned int* err;
};

__device__ __forceinline__ void tma_load_5d(void* smem_dst, const void* map, uint64_t* bar, int c0, int c1, int c2, int c3,
                                            int c4) {
    asm volatile(
        "cp.async.bulk.tensor.5d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5, %6, %7}], [%2];"
        ::"r"(smem_u32(smem_dst)), "l"(reinterpret_cast<uint64_t>(map)), "r"(smem_u32(bar)), "r"(c0), "r"(c1), "r"(c2),
        "r"(c3), "r"(c4)
        : "memory");
}
__device__ __forceinline__ void tma_load_3d_mc(void* smem_dst, const void* map, uint64_t* bar, int c0, int c1, int c2,
                                               uint16_t mask) {
    asm volatile(
        "cp.async.bulk.tensor.3d.shared::cluster.global.mbarrier::complete_tx::bytes.multicast::cluster"
        " [%0], [%1, {%4, %5, %6}], [%2], %3;"
        ::"r"(smem_u32(smem_dst)), "l"(reinterpret_cast<uint64_t>(map)), "r"(smem_u32(bar)), "h"(mask), "r"(c0), "r"(c1),
        "r"(c2)
        : "memory");
}
__device__ __forceinline__ void umma_commit_mc(uint64_t* bar, uint16_t mask) {
    asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.multicast::cluster.b64 [%0], %1;"
                 ::"r"(smem_u32(bar)), "h"(mask)
                 : "memory");
}
__device__ __forceinline__ uint32_t cluster_ctarank() {
    uint32_t r;
    asm volatile("mov.u32 %0, %%cluster_ctarank;" : "=r"(r));
    return r;
}
__device__ __forceinline__ void cluster_sync_all() {
    asm volatile("barrier.cluster.arrive.release.aligned;" ::: "memory");
    asm volatile("barrier.cluster.wait.acquire.aligned;" ::: "memory");
}
template <int BN, bool FUSE_>
struct Conv4Cfg {
    static constexpr int B_BYTES = BN * 64 * 2;             // one plane of one weight tile
    static constexpr int B_STAGE = 2 * B_BYTES;
    // FUSE: a_hi*[b_hi|b_lo] as ONE MMA with N = 2*BN (the planes are adjacent in a stage), a_lo*b_hi with N = BN into
    // the first column range; the epilogue adds the two ranges (same scheme as conv_tc2, see Conv2Cfg::FUSE)
    static constexpr bool FUSE = FUSE_ && BN <= 128;
    static constexpr int ACC_COLS = FUSE ? 2 * BN : BN;
    static constexpr int NBUF = 512 / ACC_COLS >= 4 ? 4 : 2;
    static constexpr int TMEM_COLS = NBUF * ACC_COLS;
    static constexpr int CH = 4;
    static constexpr int EPI_WARPS = BN == 256 ? 8 : 4;
    static constexpr int THREADS = 96 + 32 * EPI_WARPS;    // + TMA-A producer warp (last warp)
    static constexpr int NACC = BN / (EPI_WARPS / 4);
    static constexpr int AUX_BYTES = 512 + BN * 4;
    static constexpr int STG_BYTES = 4 * 8192;              // per-epilogue-warp store staging
    static constexpr int MAX_B_STAGES = 6;
    static constexpr int MAX_A_STAGES = 4;
    static constexpr int TR = 8, TC = 16;                   // output tile: 8 rows x 16 columns
    static constexpr int A_PLANE = (TR + 2) * TC * 128;     // one plane of one kx patch: 10 x 16 rows of 128 B = 20 KB
    static constexpr int A_STAGE = 2 * A_PLANE;
    static constexpr int KY_BYTES = TC * 128;               // one patch row = 2048 B = two swizzle atoms
};

struct Tile4 {
    int n0;
    int img, y0, x0;     // image and first output pixel of the tile (ragged edges shift inwards: overlap is recomputed identically)
    bool live;           // false: padding tile of an odd cluster pair (runs the pipeline, stores nothing)
};

__device__ __forceinline__ Tile4 tile4(const Conv4Params& p, int m_tile, int n_idx, int BN) {
    Tile4 t;
    t.n0 = n_idx * BN;
    t.live = m_tile < p.m_tiles;
    const int mt = t.live ? m_tile : p.m_tiles - 1;
    const int per_img = p.tiles_x * p.tiles_y;
    t.img = mt / per_img;
    const int r = mt - t.img * per_img;
    const int ty = r / p.tiles_x, tx = r - ty * p.tiles_x;
    t.y0 = min(ty * 8, p.H - 8);
    t.x0 = min(tx * 16, p.W - 16);
    return t;
}

template <int BN, bool FUSE_>
__global__ void __launch_bounds__(Conv4Cfg<BN, FUSE_>::THREADS, 1)
conv_tc4_kernel(const __grid_constant__ CUtensorMap mapA, const __grid_constant__ CUtensorMap mapB, const Conv4Params p) {
    using Cfg = Conv4Cfg<BN, FUSE_>;
    extern __shared__ uint8_t smem_raw[];
    uint8_t* smem = smem_raw + ((1024u - (smem_u32(smem_raw) & 1023u)) & 1023u);
    uint8_t* a_base = smem;
    uint8_t* b_base = smem + p.na_stages * Cfg::A_STAGE;
    uint8_t* aux = b_base + p.nb_stages * Cfg::B_STAGE;
    uint64_t* fullA = reinterpret_cast<uint64_t*>(aux);
    uint64_t* emptyA = fullA + Cfg::MAX_A_STAGES;
    uint64_t* fullB = emptyA + Cfg::MAX_A_STAGES;
    uint64_t* emptyB = fullB + Cfg::MAX_B_STAGES;
    uint64_t* tfull = emptyB + Cfg::MAX_B_STAGES;
    uint64_t* tempty = tfull + Cfg::NBUF;
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(tempty + Cfg::NBUF);
    volatile int* abort_flag = reinterpret_cast<volatile int*>(tmem_slot + 1);
    float* sbias = reinterpret_cast<float*>(aux + 512);

    const int warp = threadIdx.x >> 5;
    const int lane = threadIdx.x & 31;
    const uint32_t crank = p.cluster > 1 ? cluster_ctarank() : 0u;
    const uint16_t mc_mask = (uint16_t)((1u << p.cluster) - 1u);

    // NOTE: no early exit on a previous error here: both CTAs of a cluster must take the same path
    if (threadIdx.x == 0) {
        for (int s = 0; s < p.na_stages; ++s) {
            mbar_init(&fullA[s], 1);
            mbar_init(&emptyA[s], 1);
        }
        for (int s = 0; s < p.nb_stages; ++s) {
            mbar_init(&fullB[s], 1);
            mbar_init(&emptyB[s], p.cluster);          // released by every CTA of the cluster
        }
        for (int b = 0; b < Cfg::NBUF; ++b) {
            mbar_init(&tfull[b], 1);
            mbar_init(&tempty[b], Cfg::EPI_WARPS);
        }
        *abort_flag = 0;
        fence_barrier_init();
        tma_prefetch_desc(&mapA);
        tma_prefetch_desc(&mapB);
    }
    if (warp == 1) tmem_alloc(tmem_slot, Cfg::TMEM_COLS);
    tc_fence_before();
    __syncthreads();
    if (p.cluster > 1) cluster_sync_all();             // peers' barriers are initialised before any multicast lands
    tc_fence_after();
    const uint32_t tmem_base = *tmem_slot;

    const int ksl = p.Cin / 64;
    const int kiters = 9 * ksl;
    const int nchunks = (kiters + Cfg::CH - 1) / Cfg::CH;
    const int num_clusters = gridDim.x / p.cluster;
    const int cid = blockIdx.x / p.cluster;
    const int m_pairs = (p.m_tiles + p.cluster - 1) / p.cluster;
    const int total_q = m_pairs * p.n_tiles;

    if (warp == 0) {
        // ---- TMA producer, weights (B): one [BN x 64] tile pair per (slice, tap) ----
        if (lane == 0) {
            uint32_t ib = 0;
            for (int q = cid; q < total_q; q += num_clusters) {
                const int n0 = (q % p.n_tiles) * BN;
                for (int ks = 0; ks < ksl; ++ks) {
                    for (int r = 0; r < 9; ++r, ++ib) {
                        const int tap = (r % 3) * 3 + r / 3;           // kx outer, ky inner: the order the A patches arrive in
                        const int sb = ib % p.nb_stages;
                        mbar_wait(&emptyB[sb], ((ib / p.nb_stages) & 1) ^ 1u, abort_flag, p.err, 0x120u + sb);
                        uint8_t* sbp = b_base + sb * Cfg::B_STAGE;
                        if (p.dbg & 4) { mbar_arrive(&fullB[sb]); continue; }
                        mbar_arrive_expect_tx(&fullB[sb], Cfg::B_STAGE);
                        const int kc = tap * p.Cin + ks * 64;
                        if (p.cluster > 1) {
                            // this CTA fetches rows [crank*BN/2, +BN/2) of the tile and multicasts them to both CTAs
                            const int half = BN / 2;
                            const int ro = (int)crank * half;
                            tma_load_3d_mc(sbp + ro * 128, &mapB, &fullB[sb], kc, n0 + ro, 0, mc_mask);
                            tma_load_3d_mc(sbp + Cfg::B_BYTES + ro * 128, &mapB, &fullB[sb], kc, n0 + ro, 1, mc_mask);
                        } else {
                            tma_load_3d(sbp, &mapB, &fullB[sb], kc, n0, 0);
                            tma_load_3d(sbp + Cfg::B_BYTES, &mapB, &fullB[sb], kc, n0, 1);
                        }
                    }
                }
            }
        }
        __syncwarp();
    } else if (warp == 2 + Cfg::EPI_WARPS) {
        // ---- TMA producer, activations (A): its own warp so the next patch is requested as soon
        //      as its stage is free instead of queueing behind nine weight loads ----
        if (lane == 0) {
            uint32_t ia = 0;
            for (int q = cid; q < total_q; q += num_clusters) {
                const Tile4 t = tile4(p, (q / p.n_tiles) * p.cluster + (int)crank, q % p.n_tiles, BN);
                for (int ks = 0; ks < ksl; ++ks) {
                    for (int kx = 0; kx < 3; ++kx, ++ia) {
                        const int sa = ia % p.na_stages;
                        mbar_wait(&emptyA[sa], ((ia / p.na_stages) & 1) ^ 1u, abort_flag, p.err, 0x110u + sa);
                        uint8_t* st = a_base + sa * Cfg::A_STAGE;
                        if (p.dbg & 2) { mbar_arrive(&fullA[sa]); continue; }
                        mbar_arrive_expect_tx(&fullA[sa], (uint32_t)Cfg::A_STAGE);
                        // padded coordinates: output pixel (y, x) and tap (ky, kx) read padded (y + ky, x + kx)
                        tma_load_5d(st, &mapA, &fullA[sa], ks * 64, t.x0 + kx, t.y0, t.img, 0);
                        tma_load_5d(st + Cfg::A_PLANE, &mapA, &fullA[sa], ks * 64, t.x0 + kx, t.y0, t.img, 1);
                    }
                }
            }
        }
        __syncwarp();
    } else if (warp == 1) {
        if (lane == 0) {
            constexpr uint32_t idesc = umma_idesc_f16(128, BN);
            constexpr uint32_t idesc2 = umma_idesc_f16(128, Cfg::FUSE ? 2 * BN : BN);
            // ring positions as running counters (a timeline probe showed ~600 cycles per k-iteration of bookkeeping in
            // this single thread with runtime % and / on the stage counts -- as much as issuing the MMAs)
            const uint32_t a0 = smem_u32(a_base), b0 = smem_u32(b_base);
            int sa = 0, sb = 0, b = 0;
            uint32_t pha = 0, phb = 0, pht = 0, ib = 0;
            for (int q = cid; q < total_q; q += num_clusters) {
                int it = 0, ky = 0;
                for (int c = 0; c < nchunks; ++c) {
                    mbar_wait(&tempty[b], pht ^ 1u, abort_flag, p.err, 0x400u + b);
                    tc_fence_after();
                    const uint32_t tacc = tmem_base + (uint32_t)(b * Cfg::ACC_COLS);
                    const int it_end = min(kiters, (c + 1) * Cfg::CH);
                    bool first = true;
                    for (; it < it_end; ++it, ++ib) {
                        if (ky == 0) mbar_wait(&fullA[sa], pha, abort_flag, p.err, 0x210u + sa);
                        if (p.trace && blockIdx.x == 0 && ib < 256) p.trace[ib] = clock64();               // k-iter start
                        mbar_wait(&fullB[sb], phb, abort_flag, p.err, 0x220u + sb);
                        tc_fence_after();
                        if (p.trace && blockIdx.x == 0 && ib < 256) p.trace[256 + ib] = clock64();         // operands ready
                        const uint64_t a_hi = umma_desc_sw128(a0 + (uint32_t)(sa * Cfg::A_STAGE + ky * Cfg::KY_BYTES));
                        const uint64_t a_lo = a_hi + (uint64_t)(Cfg::A_PLANE >> 4);
                        const uint64_t b_hi = umma_desc_sw128(b0 + (uint32_t)(sb * Cfg::B_STAGE));
                        const uint64_t b_lo = b_hi + (uint64_t)(Cfg::B_BYTES >> 4);
                        if (!(p.dbg & 1)) {
#pragma unroll
                            for (int k = 0; k < 4; ++k) {
                                const uint64_t ko = (uint64_t)(k * 32 >> 4);
                                if (Cfg::FUSE) {
                                    umma_f16(tacc, a_hi + ko, b_hi + ko, idesc2, (first && k == 0) ? 0u : 1u);
                                    umma_f16(tacc, a_lo + ko, b_hi + ko, idesc, 1u);
                                } else {
                                    umma_f16(tacc, a_hi + ko, b_lo + ko, idesc, (first && k == 0) ? 0u : 1u);
                                    umma_f16(tacc, a_lo + ko, b_hi + ko, idesc, 1u);
                                    umma_f16(tacc, a_hi + ko, b_hi + ko, idesc, 1u);
                                }
                            }
                        }
                        first = false;
                        if (p.cluster > 1) umma_commit_mc(&emptyB[sb], mc_mask);
                        else umma_commit(&emptyB[sb]);
                        if (++sb == p.nb_stages) { sb = 0; phb ^= 1u; }
                        if (++ky == 3) {
                            ky = 0;
                            umma_commit(&emptyA[sa]);
                            if (++sa == p.na_stages) { sa = 0; pha ^= 1u; }
                        }
                        if (p.trace && blockIdx.x == 0 && ib < 256) p.trace[512 + ib] = clock64();         // MMAs + commits issued
                    }
                    umma_commit(&tfull[b]);
                    if (++b == Cfg::NBUF) { b = 0; pht ^= 1u; }
                }
            }
        }
        __syncwarp();
    } else {
        const int e = warp - 2;
        const int g = warp & 3;
        const int colbase = (e >> 2) * Cfg::NACC;
        const int et = threadIdx.x - 64;
        constexpr int ETHREADS = 32 * Cfg::EPI_WARPS;
        const ActGeom go(p.N, p.H, p.W, p.Cout);
        const bool relu = (p.flags & WCTB200_RELU) != 0;
        const int m = g * 32 + lane;
        uint32_t cg_ = 0;
        for (int q = cid; q < total_q; q += num_clusters) {
            const Tile4 t = tile4(p, (q / p.n_tiles) * p.cluster + (int)crank, q % p.n_tiles, BN);
            asm volatile("bar.sync 1, %0;" ::"r"(ETHREADS) : "memory");
            for (int i = et; i < BN; i += ETHREADS) sbias[i] = p.bias ? p.bias[t.n0 + i] : 0.f;
            asm volatile("bar.sync 1, %0;" ::"r"(ETHREADS) : "memory");

            float acc[Cfg::NACC];
#pragma unroll
            for (int i = 0; i < Cfg::NACC; ++i) acc[i] = 0.f;
            for (int c = 0; c < nchunks; ++c, ++cg_) {
                const int b = cg_ % Cfg::NBUF;
                mbar_wait(&tfull[b], (cg_ / Cfg::NBUF) & 1, abort_flag, p.err, 0x300u + b);
                tc_fence_after();
                if (p.trace && blockIdx.x == 0 && threadIdx.x == 64 && cg_ < 128) p.trace[768 + 2 * cg_] = clock64();   // chunk complete seen
                const uint32_t tsrc = tmem_base + ((uint32_t)(g * 32) << 16) + (uint32_t)(b * Cfg::ACC_COLS + colbase);
                if (Cfg::FUSE) {
#pragma unroll
                    for (int c0 = 0; c0 < Cfg::NACC; c0 += 32) {
                        uint32_t r0[32], r1[32];
                        tmem_ld32(tsrc + c0, r0);               // a_hi b_hi + a_lo b_hi
                        tmem_ld32(tsrc + BN + c0, r1);          // a_hi b_lo
                        tmem_ld_wait();
#pragma unroll
                        for (int j = 0; j < 32; ++j) acc[c0 + j] += __uint_as_float(r0[j]) + __uint_as_float(r1[j]);
                    }
                } else {
#pragma unroll
                    for (int c0 = 0; c0 < Cfg::NACC; c0 += 64) {
                        uint32_t r0[32], r1[32];
                        tmem_ld32(tsrc + c0, r0);
                        if (c0 + 32 < Cfg::NACC) tmem_ld32(tsrc + c0 + 32, r1);
                        tmem_ld_wait();
#pragma unroll
                        for (int j = 0; j < 32; ++j) acc[c0 + j] += __uint_as_float(r0[j]);
                        if (c0 + 32 < Cfg::NACC) {
#pragma unroll
                            for (int j = 0; j < 32; ++j) acc[c0 + 32 + j] += __uint_as_float(r1[j]);
                        }
                    }
                }
                tc_fence_before();
                __syncwarp();
                if (p.trace && blockIdx.x == 0 && threadIdx.x == 64 && cg_ < 128) p.trace[768 + 2 * cg_ + 1] = clock64(); // chunk drained
                if (lane == 0) mbar_arrive(&tempty[b]);
            }
            const bool valid = t.live;
            const int n = t.img, y = t.y0 + (m >> 4), x = t.x0 + (m & 15);
            store_tile_rows<Cfg::NACC>(acc, sbias + colbase, relu, aux + Cfg::AUX_BYTES + e * 8192, lane, valid && !*abort_flag,
                                       n, y, x, p.out, go, t.n0 + colbase);
        }
    }
    tc_fence_before();
    __syncthreads();
    if (p.cluster > 1) cluster_sync_all();      // no CTA exits while its peer may still multicast into it
    if (warp == 1) tmem_dealloc(tmem_base, Cfg::TMEM_COLS);
}

// ---------------------------------------------------------------------------
// host side
// ---------------------------------------------------------------------------
typedef CUresult (*PFN_encodeTiled4)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*,
                                     const cuuint64_t*, const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave,
                                     CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

static PFN_encodeTiled4 get_encode4() {
    static PFN_encodeTiled4 fn = nullptr;
    if (!fn) {
        void* ptr = nullptr;
        cudaDriverEntryPointQueryResult q;
        if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &ptr, cudaEnableDefault, &q) == cudaSuccess &&
            q == cudaDriverEntryPointSuccess)
            fn = reinterpret_cast<PFN_encodeTiled4>(ptr);
    }
    return fn;
}

static int make_map4(CUtensorMap* m, const void* base, int rank, const cuuint64_t* dims, const cuuint64_t* strides,
                      const cuuint32_t* box) {
    PFN_encodeTiled4 enc = get_encode4();
    if (!enc) {
        set_error("cuTensorMapEncodeTiled entry point not available");
        return WCTB200_ECUDA;
    }
    cuuint32_t estr[5] = {1, 1, 1, 1, 1};
    CUresult r = enc(m, CU_TENSOR_MAP_DATA_TYPE_FLOAT16, (cuuint32_t)rank, const_cast<void*>(base), dims, strides, box, estr,
                     CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                     CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (r != CUDA_SUCCESS) {
        set_error("cuTensorMapEncodeTiled(rank %d) failed (%d)", rank, (int)r);
        return WCTB200_ECUDA;
    }
    return 0;
}

extern int g_conv_oversub;
extern int g_conv_fuse;
int g_conv4_cluster = 2;     // tuning hook (wctb200_debug_set_conv4): 1 = no cluster, 2 = weight multicast across a CTA pair
int g_conv4_dbg = 0;
long long* g_conv4_trace = nullptr;
int g_conv4_cin_max = 0;     // impl 2 dispatches to v4 for Cin <= this (0: never -- v4 is MMA-issue bound)

template <int BN, bool FUSE_>
static int launch4_bn(const CUtensorMap& mA, const CUtensorMap& mB, const Conv4Params& p, int smem_bytes, cudaStream_t st) {
    using Cfg = Conv4Cfg<BN, FUSE_>;
    static int sms = 0;
    static int attr_bytes = 0;
    if (!sms) {
        int dev = 0;
        WCTB_CUDA(cudaGetDevice(&dev));
        WCTB_CUDA(cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev));
    }
    if (smem_bytes > attr_bytes) {
        WCTB_CUDA(cudaFuncSetAttribute(conv_tc4_kernel<BN, FUSE_>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem_bytes));
        attr_bytes = smem_bytes;
    }
    const int m_pairs = (p.m_tiles + p.cluster - 1) / p.cluster;
    const int total_q = m_pairs * p.n_tiles;
    // persistent grid, over-subscribed like v2 so that queued CTAs of other streams interleave at CTA granularity
    int clusters = (sms / p.cluster) * (g_conv_oversub > 0 ? g_conv_oversub : 1);
    if (clusters > total_q) clusters = total_q;
    cudaLaunchConfig_t cfg = {};
    cfg.gridDim = dim3((unsigned)(clusters * p.cluster), 1, 1);
    cfg.blockDim = dim3(Cfg::THREADS, 1, 1);
    cfg.dynamicSmemBytes = (size_t)smem_bytes;
    cfg.stream = st;
    cudaLaunchAttribute at[1];
    at[0].id = cudaLaunchAttributeClusterDimension;
    at[0].val.clusterDim.x = (unsigned)p.cluster;
    at[0].val.clusterDim.y = 1;
    at[0].val.clusterDim.z = 1;
    cfg.attrs = at;
    cfg.numAttrs = 1;
    WCTB_CUDA(cudaLaunchKernelEx(&cfg, conv_tc4_kernel<BN, FUSE_>, mA, mB, p));
    return 0;
}

// returns 1 if this shape is handled by v4 (and launched), 0 if the caller should use v2, <0 on error
int launch_conv3x3_tc4(const __half* in, int N, int H, int W, int Cin, const __half* w_split, const float* bias, int Cout,
                       int flags, __half* out, int bn_override, cudaStream_t st) {
    if (Cin % 64 || Cout % 64 || H < 8 || W < 16) return 0;
    ActGeom gi(N, H, W, Cin);
    if (gi.P >= (1ll << 31) - 4096) return 0;
    Conv4Params p;
    p.N = N; p.H = H; p.W = W; p.Cin = Cin; p.Cout = Cout; p.Hp = gi.Hp; p.Wp = gi.Wp; p.P = gi.P;
    p.flags = flags; p.bias = bias; p.out = out; p.err = device_error_word();
    p.cluster = g_conv4_cluster == 1 ? 1 : 2;
    p.dbg = g_conv4_dbg;
    p.trace = g_conv4_trace;
    if (p.dbg) p.cluster = 1;
    int BN = Cout % 128 == 0 ? 128 : 64;
    if (bn_override && Cout % bn_override == 0 && bn_override <= 128) BN = bn_override;
    p.n_tiles = Cout / BN;
    p.tiles_x = (W + 15) / 16;
    p.tiles_y = (H + 7) / 8;
    p.m_tiles = N * p.tiles_x * p.tiles_y;

    CUtensorMap mA, mB;
    {
        cuuint64_t dims[5] = {(cuuint64_t)Cin, (cuuint64_t)gi.Wp, (cuuint64_t)gi.Hp, (cuuint64_t)N, 2};
        cuuint64_t strides[4] = {(cuuint64_t)Cin * 2, (cuuint64_t)gi.Wp * Cin * 2, (cuuint64_t)gi.Hp * gi.Wp * Cin * 2,
                                 (cuuint64_t)gi.plane * 2};
        cuuint32_t box[5] = {64, 16, 10, 1, 1};
        int rc = make_map4(&mA, in, 5, dims, strides, box);
        if (rc) return rc;
    }
    {
        const cuuint64_t K = (cuuint64_t)9 * Cin;
        cuuint64_t dims[3] = {K, (cuuint64_t)Cout, 2};
        cuuint64_t strides[2] = {K * 2, K * Cout * 2};
        cuuint32_t box[3] = {64, (cuuint32_t)(p.cluster > 1 ? BN / 2 : BN), 1};
        int rc = make_map4(&mB, w_split, 3, dims, strides, box);
        if (rc) return rc;
    }
    // shared memory: 3 A stages (one slice = 3 kx patches) + as many B stages as fit (<= 6); a 4th A stage if room is left
    const int a_stage = 2 * 10 * 16 * 128;
    const int b_stage = 2 * BN * 128;
    const int aux = 512 + BN * 4 + 4 * 8192 + 1024;
    int na = 3;
    int nb = (227 * 1024 - aux - na * a_stage) / b_stage;
    if (nb > 6) nb = 6;
    if (nb < 2) return 0;
    if (227 * 1024 - aux - na * a_stage - nb * b_stage >= a_stage) na = 4;
    p.na_stages = na;
    p.nb_stages = nb;
    const int smem_bytes = na * a_stage + nb * b_stage + aux;
    const bool fuse = g_conv_fuse != 0 && (BN == 64 || g_conv_fuse == 1 || Cin >= 256);
    int rc = BN == 128 ? (fuse ? launch4_bn<128, true>(mA, mB, p, smem_bytes, st) : launch4_bn<128, false>(mA, mB, p, smem_bytes, st))
                       : (fuse ? launch4_bn<64, true>(mA, mB, p, smem_bytes, st) : launch4_bn<64, false>(mA, mB, p, smem_bytes, st));
    return rc ? rc : 1;
}

}  // namespace wctb
