// Conv2DReflect 3x3 (and 1x1 "apply") as an implicit GEMM on the 5th-gen tensor cores.
//
//   reference: Lambda(pad_reflect) -> Conv2D(valid) (+ReLU)   ops.py:12-19,
//              vgg_normalised.py:28-40, model.py:291; and the whitening/colouring
//              apply  M * fc  of ops.py:73,77 when taps == 1.
//
// Formulation.  Activations are SPF16: two fp16 planes (hi, lo) of the reflect-padded
// NHWC tensor, viewed as a 2-D matrix [P = N*(H+2)*(W+2) padded positions][C].  For an
// output position p (padded coordinates) and filter tap (ky,kx) the input row is simply
// p + (ky-1)*(W+2) + (kx-1): each tap of a 128-position output tile is ONE dense
// 128 x 64 TMA box at a shifted row coordinate (rows outside [0,P) are zero-filled by
// TMA and only feed halo/junk outputs, which the epilogue never stores).
//
//   D[128 pos][BN cout] = sum_{tap, cin-slice}  A_tap[128][64] * W_tap[BN][64]^T
//
// Precision.  fp32 accuracy from fp16 tensor-core inputs: x = x_hi + x_lo (22 bits),
//   x*w ~= x_hi*w_hi + x_hi*w_lo + x_lo*w_hi     (3 x tcgen05.mma kind::f16, fp32 accumulate in TMEM)
// the dropped lo*lo term is 2^-22 relative.
//
// Structure (one 128 x BN output tile per CTA, 192 threads):
//   warp 0   : TMA producer   (cp.async.bulk.tensor 3-D, 128B swizzle, mbarrier complete_tx)
//   warp 1   : MMA issuer     (one elected thread, tcgen05.mma / tcgen05.commit), owns TMEM alloc
//   warps 2-5: epilogue       (tcgen05.ld 32x32b -> +bias, ReLU -> split fp16 -> 16-byte stores
//                              of the interior pixel AND the halo cells that mirror it)
#include "common.cuh"

namespace wctb {

extern int g_conv_bn_override;
extern int g_conv_impl;

struct ConvParams {
    int N, H, W, Cin, Cout, Hp, Wp;
    long long P;
    int taps;             // 9 or 1
    int per_image;        // tiles never straddle images; weight/bias set = image index if nsets > 1
    int nsets;
    int tiles_per_image;
    int flags;
    const float* bias;    // [nsets][Cout] or nullptr
    __half* out;          // SPF16, Cout channels, same N,H,W
    unsigned int* err;
};

template <int BN>
struct ConvCfg {
    static constexpr int BM = 128;
    static constexpr int BK = 64;
    static constexpr int A_BYTES = BM * BK * 2;          // one plane of the A tile
    static constexpr int B_BYTES = BN * BK * 2;
    static constexpr int STAGE_BYTES = 2 * A_BYTES + 2 * B_BYTES;
    static constexpr int STAGES = BN == 64 ? 4 : (BN == 128 ? 3 : 2);
    static constexpr int AUX_BYTES = 256 + BN * 4;       // barriers, tmem slot, abort flag, bias tile
    static constexpr int SMEM_BYTES = STAGES * STAGE_BYTES + AUX_BYTES + 1024;  // + alignment slack
};

template <int BN>
__global__ void __launch_bounds__(192, 1)
conv_tc_kernel(const __grid_constant__ CUtensorMap mapA, const __grid_constant__ CUtensorMap mapB, const ConvParams p) {
    using Cfg = ConvCfg<BN>;
    extern __shared__ uint8_t smem_raw[];
    // SWIZZLE_128B tiles need 1024-byte alignment
    uint8_t* smem = smem_raw + ((1024u - (smem_u32(smem_raw) & 1023u)) & 1023u);
    uint8_t* aux = smem + Cfg::STAGES * Cfg::STAGE_BYTES;
    uint64_t* full = reinterpret_cast<uint64_t*>(aux);
    uint64_t* empty = full + Cfg::STAGES;
    uint64_t* tmem_full = empty + Cfg::STAGES;
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(tmem_full + 1);
    volatile int* abort_flag = reinterpret_cast<volatile int*>(tmem_slot + 1);
    float* sbias = reinterpret_cast<float*>(aux + 256);

    const int warp = threadIdx.x >> 5;
    const int lane = threadIdx.x & 31;

    {   // an earlier CTA already timed out: leave (uniformly) instead of waiting 2 s per CTA
        __shared__ unsigned int s_prev_err;
        if (threadIdx.x == 0) s_prev_err = *reinterpret_cast<volatile unsigned int*>(p.err);
        __syncthreads();
        if (s_prev_err != 0u) return;
    }

    const long long HpWp = (long long)p.Hp * p.Wp;
    long long p0, p_end;
    int set = 0;
    if (p.per_image) {
        const int img = blockIdx.x / p.tiles_per_image;
        const int t = blockIdx.x - img * p.tiles_per_image;
        p0 = img * HpWp + (long long)t * Cfg::BM;
        p_end = (img + 1) * HpWp;
        set = p.nsets > 1 ? img : 0;
    } else {
        p0 = (long long)blockIdx.x * Cfg::BM;
        p_end = p.P;
    }
    const int n0 = blockIdx.y * BN;

    if (threadIdx.x == 0) {
        for (int s = 0; s < Cfg::STAGES; ++s) {
            mbar_init(&full[s], 1);
            mbar_init(&empty[s], 1);
        }
        mbar_init(tmem_full, 1);
        *abort_flag = 0;
        fence_barrier_init();
        tma_prefetch_desc(&mapA);
        tma_prefetch_desc(&mapB);
    }
    if (warp == 1) tmem_alloc(tmem_slot, BN);
    if (warp >= 2) {
        for (int i = threadIdx.x - 64; i < BN; i += 128)
            sbias[i] = p.bias ? p.bias[(long long)set * p.Cout + n0 + i] : 0.f;
    }
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem_base = *tmem_slot;

    const int ksl = p.Cin / Cfg::BK;
    const int kiters = p.taps * ksl;

    if (warp == 0) {
        if (lane == 0) {
            for (int it = 0; it < kiters; ++it) {
                const int s = it % Cfg::STAGES;
                const uint32_t ph = (it / Cfg::STAGES) & 1;
                mbar_wait(&empty[s], ph ^ 1u, abort_flag, p.err, 0x100u + s);
                const int tap = it / ksl;
                const int ks = it - tap * ksl;
                const int off = p.taps == 9 ? (tap / 3 - 1) * p.Wp + (tap % 3 - 1) : 0;
                const int row = (int)(p0 + off);
                uint8_t* st = smem + s * Cfg::STAGE_BYTES;
                mbar_arrive_expect_tx(&full[s], Cfg::STAGE_BYTES);
                tma_load_3d(st, &mapA, &full[s], ks * Cfg::BK, row, 0);
                tma_load_3d(st + Cfg::A_BYTES, &mapA, &full[s], ks * Cfg::BK, row, 1);
                tma_load_3d(st + 2 * Cfg::A_BYTES, &mapB, &full[s], tap * p.Cin + ks * Cfg::BK, n0, set * 2);
                tma_load_3d(st + 2 * Cfg::A_BYTES + Cfg::B_BYTES, &mapB, &full[s], tap * p.Cin + ks * Cfg::BK, n0,
                            set * 2 + 1);
            }
        }
        __syncwarp();
    } else if (warp == 1) {
        if (lane == 0) {
            constexpr uint32_t idesc = umma_idesc_f16(Cfg::BM, BN);
            for (int it = 0; it < kiters; ++it) {
                const int s = it % Cfg::STAGES;
                const uint32_t ph = (it / Cfg::STAGES) & 1;
                mbar_wait(&full[s], ph, abort_flag, p.err, 0x200u + s);
                tc_fence_after();
                const uint32_t st = smem_u32(smem + s * Cfg::STAGE_BYTES);
                const uint64_t a_hi = umma_desc_sw128(st);
                const uint64_t a_lo = umma_desc_sw128(st + Cfg::A_BYTES);
                const uint64_t b_hi = umma_desc_sw128(st + 2 * Cfg::A_BYTES);
                const uint64_t b_lo = umma_desc_sw128(st + 2 * Cfg::A_BYTES + Cfg::B_BYTES);
#pragma unroll
                for (int k = 0; k < Cfg::BK / 16; ++k) {
                    const uint64_t ko = (uint64_t)(k * 32 >> 4);  // +32 bytes of K per UMMA_K=16 step
                    umma_f16(tmem_base, a_hi + ko, b_lo + ko, idesc, (it | k) != 0 ? 1u : 0u);
                    umma_f16(tmem_base, a_lo + ko, b_hi + ko, idesc, 1u);
                    umma_f16(tmem_base, a_hi + ko, b_hi + ko, idesc, 1u);
                }
                umma_commit(&empty[s]);   // frees the smem stage when these MMAs retire
            }
            umma_commit(tmem_full);       // accumulator complete
        }
        __syncwarp();
    } else {
        // ---- epilogue: TMEM -> regs -> bias/ReLU -> split fp16 -> global (+ reflect halo) ----
        mbar_wait(tmem_full, 0u, abort_flag, p.err, 0x300u);
        tc_fence_after();
        const int g = warp & 3;                 // TMEM lanes [32g, 32g+32) belong to this warp
        const long long pos = p0 + g * 32 + lane;
        bool valid = pos < p_end;
        int n = 0, y = 0, x = 0;
        if (valid) {
            n = (int)(pos / HpWp);
            const int r = (int)(pos - n * HpWp);
            const int yy = r / p.Wp;
            const int xx = r - yy * p.Wp;
            valid = (yy >= 1) && (yy <= p.H) && (xx >= 1) && (xx <= p.W);
            y = yy - 1;
            x = xx - 1;
        }
        const ActGeom go(p.N, p.H, p.W, p.Cout);
        const bool relu = (p.flags & WCTB200_RELU) != 0;
#pragma unroll 1
        for (int c0 = 0; c0 < BN; c0 += 32) {
            uint32_t r[32];
            tmem_ld32(tmem_base + ((uint32_t)(g * 32) << 16) + (uint32_t)c0, r);
            tmem_ld_wait();
            if (valid && !*abort_flag) {
#pragma unroll
                for (int q = 0; q < 4; ++q) {
                    float v[8];
#pragma unroll
                    for (int j = 0; j < 8; ++j) {
                        float t = __uint_as_float(r[q * 8 + j]) + sbias[c0 + q * 8 + j];
                        v[j] = relu ? fmaxf(t, 0.f) : t;
                    }
                    Half8 hi, lo;
                    split8(v, hi, lo);
                    store8_with_halo(p.out, go, n, y, x, n0 + c0 + q * 8, hi, lo);
                }
            }
        }
    }
    tc_fence_before();
    __syncthreads();
    if (warp == 1) tmem_dealloc(tmem_base, BN);
}

// ===========================================================================
// v2: persistent CTAs + chunked accumulation drained into registers
//
//   * grid = #SMs, every CTA walks a static tile list (cout tile fastest, so CTAs running
//     together share activation tiles in L2 and the whole weight tensor stays L2 resident);
//   * the accumulator ring: TMEM holds NBUF buffers of BN fp32 columns.  The MMA warp
//     accumulates CH k-iterations (CH*4*3 tcgen05.mma) into one buffer starting from zero,
//     commits it, and moves on to the next buffer; the epilogue warps drain each finished
//     buffer with tcgen05.ld and ADD IT INTO REGISTERS with round-to-nearest FADDs.
//     The tensor core adds into its fp32 accumulator with truncation (measured: -1.2e-5
//     relative bias at K=4608 when everything is accumulated in TMEM); short chunks summed
//     in registers bring the conv back to fp32-class error, and the drain of chunk i
//     overlaps the MMAs of chunk i+1 (also across tiles: the epilogue of tile t overlaps
//     the main loop of tile t+1).
// ===========================================================================
template <int BN, bool FUSE_>
struct Conv2Cfg {
    static constexpr int BM = 128;
    static constexpr int BK = 64;
    static constexpr int A_BYTES = BM * BK * 2;
    static constexpr int B_BYTES = BN * BK * 2;
    static constexpr int STAGE_BYTES = 2 * A_BYTES + 2 * B_BYTES;
    static constexpr int STAGES = BN == 64 ? 4 : (BN == 128 ? 3 : 2);
    // FUSE: the products a_hi*b_hi and a_hi*b_lo share the A operand, and the two weight planes sit back to back in a
    // stage, so ONE tcgen05.mma with N = 2*BN computes both into adjacent accumulator column ranges (summed when the
    // chunk is drained); a_lo*b_hi follows with N = BN into the first range.  8 instead of 12 MMAs per k-iteration and
    // half the A-operand shared-memory reads per flop (timeline probe: ~80 cycles per MMA regardless of N <= 128).
    static constexpr bool FUSE = FUSE_ && BN <= 128;
    static constexpr int ACC_COLS = FUSE ? 2 * BN : BN;         // TMEM columns of one accumulation buffer
    static constexpr int NBUF = 512 / ACC_COLS >= 4 ? 4 : 2;
    static constexpr int TMEM_COLS = NBUF * ACC_COLS;           // 512 / 512 / 512
    static constexpr int CH = 4;                                // k-iterations per accumulation chunk
    static constexpr int EPI_WARPS = BN == 256 ? 8 : 4;
    static constexpr int THREADS = 64 + 32 * EPI_WARPS;
    static constexpr int NACC = BN / (EPI_WARPS / 4);           // accumulators per epilogue thread (<= 128)
    static constexpr int AUX_BYTES = 256 + BN * 4;
    static constexpr int STG_BYTES = BN <= 128 ? 4 * 8192 : 0;  // per-epilogue-warp store staging (coalesced stores)
    static constexpr int SMEM_BYTES = STAGES * STAGE_BYTES + AUX_BYTES + STG_BYTES + 1024;
};

struct TileCoord {
    long long p0, p_end;
    int n0, set;
};

__device__ __forceinline__ TileCoord tile_coord(const ConvParams& p, int tile, int n_tiles, int BN) {
    TileCoord t;
    const int nt = tile % n_tiles;
    const int mt = tile / n_tiles;
    const long long HpWp = (long long)p.Hp * p.Wp;
    t.n0 = nt * BN;
    if (p.per_image) {
        const int img = mt / p.tiles_per_image;
        const int r = mt - img * p.tiles_per_image;
        t.p0 = img * HpWp + (long long)r * 128;
        t.p_end = (img + 1) * HpWp;
        t.set = p.nsets > 1 ? img : 0;
    } else {
        t.p0 = (long long)mt * 128;
        t.p_end = p.P;
        t.set = 0;
    }
    return t;
}

template <int BN, bool FUSE_>
__global__ void __launch_bounds__(Conv2Cfg<BN, FUSE_>::THREADS, 1)
conv_tc2_kernel(const __grid_constant__ CUtensorMap mapA, const __grid_constant__ CUtensorMap mapB, const ConvParams p,
                const int total_tiles, const int n_tiles) {
    using Cfg = Conv2Cfg<BN, FUSE_>;
    extern __shared__ uint8_t smem_raw[];
    uint8_t* smem = smem_raw + ((1024u - (smem_u32(smem_raw) & 1023u)) & 1023u);
    uint8_t* aux = smem + Cfg::STAGES * Cfg::STAGE_BYTES;
    uint64_t* full = reinterpret_cast<uint64_t*>(aux);
    uint64_t* empty = full + Cfg::STAGES;
    uint64_t* tfull = empty + Cfg::STAGES;
    uint64_t* tempty = tfull + Cfg::NBUF;
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(tempty + Cfg::NBUF);
    volatile int* abort_flag = reinterpret_cast<volatile int*>(tmem_slot + 1);
    float* sbias = reinterpret_cast<float*>(aux + 256);

    const int warp = threadIdx.x >> 5;
    const int lane = threadIdx.x & 31;
    {
        __shared__ unsigned int s_prev_err;
        if (threadIdx.x == 0) s_prev_err = *reinterpret_cast<volatile unsigned int*>(p.err);
        __syncthreads();
        if (s_prev_err != 0u) return;
    }
    if (threadIdx.x == 0) {
        for (int s = 0; s < Cfg::STAGES; ++s) {
            mbar_init(&full[s], 1);
            mbar_init(&empty[s], 1);
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
    tc_fence_after();
    const uint32_t tmem_base = *tmem_slot;

    const int ksl = p.Cin / Cfg::BK;
    const int kiters = p.taps * ksl;
    const int nchunks = (kiters + Cfg::CH - 1) / Cfg::CH;

    if (warp == 0) {
        if (lane == 0) {
            uint32_t itg = 0;
            for (int tile = blockIdx.x; tile < total_tiles; tile += gridDim.x) {
                const TileCoord tc = tile_coord(p, tile, n_tiles, BN);
                for (int it = 0; it < kiters; ++it, ++itg) {
                    const int s = itg % Cfg::STAGES;
                    const uint32_t ph = (itg / Cfg::STAGES) & 1;
                    mbar_wait(&empty[s], ph ^ 1u, abort_flag, p.err, 0x100u + s);
                    const int tap = it / ksl;
                    const int ks = it - tap * ksl;
                    const int off = p.taps == 9 ? (tap / 3 - 1) * p.Wp + (tap % 3 - 1) : 0;
                    const int row = (int)(tc.p0 + off);
                    uint8_t* st = smem + s * Cfg::STAGE_BYTES;
                    mbar_arrive_expect_tx(&full[s], Cfg::STAGE_BYTES);
                    tma_load_3d(st, &mapA, &full[s], ks * Cfg::BK, row, 0);
                    tma_load_3d(st + Cfg::A_BYTES, &mapA, &full[s], ks * Cfg::BK, row, 1);
                    tma_load_3d(st + 2 * Cfg::A_BYTES, &mapB, &full[s], tap * p.Cin + ks * Cfg::BK, tc.n0, tc.set * 2);
                    tma_load_3d(st + 2 * Cfg::A_BYTES + Cfg::B_BYTES, &mapB, &full[s], tap * p.Cin + ks * Cfg::BK,
                                tc.n0, tc.set * 2 + 1);
                }
            }
        }
        __syncwarp();
    } else if (warp == 1) {
        if (lane == 0) {
            constexpr uint32_t idesc = umma_idesc_f16(Cfg::BM, BN);
            constexpr uint32_t idesc2 = umma_idesc_f16(Cfg::BM, Cfg::FUSE ? 2 * BN : BN);
            uint32_t itg = 0, cg_ = 0;
            for (int tile = blockIdx.x; tile < total_tiles; tile += gridDim.x) {
                for (int c = 0; c < nchunks; ++c, ++cg_) {
                    const int b = cg_ % Cfg::NBUF;
                    const uint32_t bph = (cg_ / Cfg::NBUF) & 1;
                    mbar_wait(&tempty[b], bph ^ 1u, abort_flag, p.err, 0x400u + b);   // epilogue drained this buffer
                    tc_fence_after();
                    const uint32_t tacc = tmem_base + (uint32_t)(b * Cfg::ACC_COLS);
                    const int it_end = min(kiters, (c + 1) * Cfg::CH);
                    for (int it = c * Cfg::CH; it < it_end; ++it, ++itg) {
                        const int s = itg % Cfg::STAGES;
                        const uint32_t ph = (itg / Cfg::STAGES) & 1;
                        mbar_wait(&full[s], ph, abort_flag, p.err, 0x200u + s);
                        tc_fence_after();
                        const uint32_t st = smem_u32(smem + s * Cfg::STAGE_BYTES);
                        const uint64_t a_hi = umma_desc_sw128(st);
                        const uint64_t a_lo = umma_desc_sw128(st + Cfg::A_BYTES);
                        const uint64_t b_hi = umma_desc_sw128(st + 2 * Cfg::A_BYTES);
                        const uint64_t b_lo = umma_desc_sw128(st + 2 * Cfg::A_BYTES + Cfg::B_BYTES);
                        const bool first = (it == c * Cfg::CH);
#pragma unroll
                        for (int k = 0; k < Cfg::BK / 16; ++k) {
                            const uint64_t ko = (uint64_t)(k * 32 >> 4);
                            if (Cfg::FUSE) {
                                // [b_hi | b_lo] is one 2*BN-row K-major tile: columns [0,BN) += a_hi b_hi, [BN,2BN) += a_hi b_lo
                                umma_f16(tacc, a_hi + ko, b_hi + ko, idesc2, (first && k == 0) ? 0u : 1u);
                                umma_f16(tacc, a_lo + ko, b_hi + ko, idesc, 1u);
                            } else {
                                umma_f16(tacc, a_hi + ko, b_lo + ko, idesc, (first && k == 0) ? 0u : 1u);
                                umma_f16(tacc, a_lo + ko, b_hi + ko, idesc, 1u);
                                umma_f16(tacc, a_hi + ko, b_hi + ko, idesc, 1u);
                            }
                        }
                        umma_commit(&empty[s]);
                    }
                    umma_commit(&tfull[b]);
                }
            }
        }
        __syncwarp();
    } else {
        // ---- epilogue warps: drain chunks into registers, then bias/ReLU/split/store ----
        const int e = warp - 2;
        const int g = warp & 3;                              // TMEM lane quadrant of this warp
        const int colbase = (e >> 2) * Cfg::NACC;            // BN=256: warps 6..9 take the upper half
        const int et = threadIdx.x - 64;                     // 0 .. 32*EPI_WARPS-1
        constexpr int ETHREADS = 32 * Cfg::EPI_WARPS;
        const long long HpWp = (long long)p.Hp * p.Wp;
        const ActGeom go(p.N, p.H, p.W, p.Cout);
        const bool relu = (p.flags & WCTB200_RELU) != 0;
        uint32_t cg_ = 0;
        for (int tile = blockIdx.x; tile < total_tiles; tile += gridDim.x) {
            const TileCoord tc = tile_coord(p, tile, n_tiles, BN);
            // stage this tile's bias slice (named barrier 1: epilogue warps only)
            asm volatile("bar.sync 1, %0;" ::"r"(ETHREADS) : "memory");
            for (int i = et; i < BN; i += ETHREADS) sbias[i] = p.bias ? p.bias[(long long)tc.set * p.Cout + tc.n0 + i] : 0.f;
            asm volatile("bar.sync 1, %0;" ::"r"(ETHREADS) : "memory");

            float acc[Cfg::NACC];
#pragma unroll
            for (int i = 0; i < Cfg::NACC; ++i) acc[i] = 0.f;
            for (int c = 0; c < nchunks; ++c, ++cg_) {
                const int b = cg_ % Cfg::NBUF;
                const uint32_t bph = (cg_ / Cfg::NBUF) & 1;
                mbar_wait(&tfull[b], bph, abort_flag, p.err, 0x300u + b);
                tc_fence_after();
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
                if (lane == 0) mbar_arrive(&tempty[b]);      // buffer may be overwritten
            }
            // ---- store: interior pixel + the halo cells that mirror it ----
            const long long pos = tc.p0 + g * 32 + lane;
            bool valid = pos < tc.p_end;
            int n = 0, y = 0, x = 0;
            if (valid) {
                n = (int)(pos / HpWp);
                const int r = (int)(pos - n * HpWp);
                const int yy = r / p.Wp;
                const int xx = r - yy * p.Wp;
                valid = (yy >= 1) && (yy <= p.H) && (xx >= 1) && (xx <= p.W);
                y = yy - 1;
                x = xx - 1;
            }
            if (Cfg::STG_BYTES > 0) {
                uint8_t* stg = aux + Cfg::AUX_BYTES + e * 8192;
                store_tile_rows<Cfg::NACC>(acc, sbias + colbase, relu, stg, lane, valid && !*abort_flag, n, y, x, p.out, go,
                                           tc.n0 + colbase);
            } else if (valid && !*abort_flag) {
#pragma unroll
                for (int q = 0; q < Cfg::NACC / 8; ++q) {
                    float v[8];
#pragma unroll
                    for (int j = 0; j < 8; ++j) {
                        const float t = acc[q * 8 + j] + sbias[colbase + q * 8 + j];
                        v[j] = relu ? fmaxf(t, 0.f) : t;
                    }
                    Half8 hi, lo;
                    split8(v, hi, lo);
                    store8_with_halo(p.out, go, n, y, x, tc.n0 + colbase + q * 8, hi, lo);
                }
            }
        }
    }
    tc_fence_before();
    __syncthreads();
    if (warp == 1) tmem_dealloc(tmem_base, Cfg::TMEM_COLS);
}

// ===========================================================================
// v5: v2 on CTA PAIRS (tcgen05 cta_group::2).  A timeline probe of the MMA-issuing thread showed every
// single-CTA tcgen05.mma (M=128, K=16) occupying the tensor pipe for >= 80 cycles (N=64) .. ~115 cycles (N=128):
// 40-55 % of the nominal rate.  A pair of CTAs on the two SMs of a TPC executes ONE M=256 instruction: each CTA
// supplies its own 128 activation rows and HALF of the weight tile (N/2 rows), accumulators stay in each CTA's own
// TMEM.  Per CTA the weight traffic halves and the operand shared-memory reads per flop drop by a third.
//   * both CTAs run the TMA producer (own A tile + own half of B); every load completes on the LEADER's `full`
//     barrier (cp.async.bulk.tensor ... .cta_group::2, barrier address mapped with mapa);
//   * only the leader's elected thread issues tcgen05.mma.cta_group::2; stage release (`empty`) and chunk completion
//     (`tfull`) are multicast commits to both CTAs;
//   * both CTAs' epilogue warps drain their own TMEM and arrive on the leader's `tempty` (remote mbarrier arrive).
// Everything else is v2 (persistent pairs, 4-k-iteration chunks summed in registers with RN adds, coalesced store).
// ===========================================================================
template <int BN>
struct Conv5Cfg {
    static constexpr int BM = 128;                              // rows per CTA (the MMA is M = 256)
    static constexpr int BK = 64;
    static constexpr int A_BYTES = BM * BK * 2;                 // one plane of this CTA's A tile
    static constexpr int BH_BYTES = (BN / 2) * BK * 2;          // one plane of this CTA's HALF of the weight tile
    static constexpr int STAGE_BYTES = 2 * A_BYTES + 2 * BH_BYTES;
    static constexpr int STAGES = 4;
    static constexpr int NBUF = 4;
    static constexpr int TMEM_COLS = NBUF * BN;
    static constexpr int CH = 4;
    static constexpr int EPI_WARPS = 4;
    static constexpr int THREADS = 64 + 32 * EPI_WARPS;
    static constexpr int NACC = BN;
    static constexpr int AUX_BYTES = 256 + BN * 4;
    static constexpr int STG_BYTES = 4 * 8192;
    static constexpr int SMEM_BYTES = STAGES * STAGE_BYTES + AUX_BYTES + STG_BYTES + 1024;
};

__device__ __forceinline__ uint32_t cluster_rank() {
    uint32_t r;
    asm volatile("mov.u32 %0, %%cluster_ctarank;" : "=r"(r));
    return r;
}
__device__ __forceinline__ void cluster_barrier() {
    asm volatile("barrier.cluster.arrive.release.aligned;" ::: "memory");
    asm volatile("barrier.cluster.wait.acquire.aligned;" ::: "memory");
}
__device__ __forceinline__ uint32_t mapa_u32(uint32_t local_smem_addr, uint32_t rank) {
    uint32_t r;
    asm volatile("mapa.shared::cluster.u32 %0, %1, %2;" : "=r"(r) : "r"(local_smem_addr), "r"(rank));
    return r;
}
__device__ __forceinline__ void mbar_arrive_remote(uint32_t cluster_addr) {
    asm volatile("mbarrier.arrive.release.cluster.shared::cluster.b64 _, [%0];" ::"r"(cluster_addr) : "memory");
}
// TMA load whose completion is signalled on a barrier that may live in the peer CTA of the pair
__device__ __forceinline__ void tma_load_3d_2sm(void* smem_dst, const void* map, uint32_t bar_cluster_addr, int c0, int c1,
                                                int c2) {
    asm volatile(
        "cp.async.bulk.tensor.3d.cta_group::2.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5}], [%2];"
        ::"r"(smem_u32(smem_dst)), "l"(reinterpret_cast<uint64_t>(map)), "r"(bar_cluster_addr), "r"(c0), "r"(c1), "r"(c2)
        : "memory");
}
__device__ __forceinline__ void tmem_alloc_2sm(uint32_t* smem_slot, uint32_t ncols) {
    asm volatile("tcgen05.alloc.cta_group::2.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(smem_slot)), "r"(ncols)
                 : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::2.sync.aligned;" ::: "memory");
}
__device__ __forceinline__ void tmem_dealloc_2sm(uint32_t taddr, uint32_t ncols) {
    asm volatile("tcgen05.dealloc.cta_group::2.sync.aligned.b32 %0, %1;" ::"r"(taddr), "r"(ncols) : "memory");
}
__device__ __forceinline__ void umma_f16_2sm(uint32_t tmem_d, uint64_t desc_a, uint64_t desc_b, uint32_t idesc,
                                             uint32_t accumulate) {
    asm volatile(
        "{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\t"
        "tcgen05.mma.cta_group::2.kind::f16 [%0], %1, %2, %3, p;\n\t}"
        ::"r"(tmem_d), "l"(desc_a), "l"(desc_b), "r"(idesc), "r"(accumulate)
        : "memory");
}
__device__ __forceinline__ void umma_commit_2sm(uint64_t* bar) {   // arrives on `bar` in BOTH CTAs of the pair
    asm volatile("tcgen05.commit.cta_group::2.mbarrier::arrive::one.shared::cluster.multicast::cluster.b64 [%0], %1;"
                 ::"r"(smem_u32(bar)), "h"((uint16_t)3)
                 : "memory");
}

template <int BN>
__global__ void __launch_bounds__(Conv5Cfg<BN>::THREADS, 1)
conv_tc5_kernel(const __grid_constant__ CUtensorMap mapA, const __grid_constant__ CUtensorMap mapB, const ConvParams p,
                const int total_q, const int n_tiles) {
    using Cfg = Conv5Cfg<BN>;
    extern __shared__ uint8_t smem_raw[];
    uint8_t* smem = smem_raw + ((1024u - (smem_u32(smem_raw) & 1023u)) & 1023u);
    uint8_t* aux = smem + Cfg::STAGES * Cfg::STAGE_BYTES;
    uint64_t* full = reinterpret_cast<uint64_t*>(aux);
    uint64_t* empty = full + Cfg::STAGES;
    uint64_t* tfull = empty + Cfg::STAGES;
    uint64_t* tempty = tfull + Cfg::NBUF;
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(tempty + Cfg::NBUF);
    volatile int* abort_flag = reinterpret_cast<volatile int*>(tmem_slot + 1);
    float* sbias = reinterpret_cast<float*>(aux + 256);

    const int warp = threadIdx.x >> 5;
    const int lane = threadIdx.x & 31;
    const uint32_t rank = cluster_rank();                     // 0 = leader (issues the MMAs)
    // NOTE: no early exit on a previous error: both CTAs of a pair must take the same path
    if (threadIdx.x == 0) {
        for (int s = 0; s < Cfg::STAGES; ++s) {
            mbar_init(&full[s], 1);                            // leader's copy is the live one: 1 arrive (leader producer) + tx bytes of both CTAs
            mbar_init(&empty[s], 1);                           // multicast commit
        }
        for (int b = 0; b < Cfg::NBUF; ++b) {
            mbar_init(&tfull[b], 1);                           // multicast commit
            mbar_init(&tempty[b], 2 * Cfg::EPI_WARPS);         // leader's copy: epilogue warps of both CTAs
        }
        *abort_flag = 0;
        fence_barrier_init();
        tma_prefetch_desc(&mapA);
        tma_prefetch_desc(&mapB);
    }
    if (warp == 1) tmem_alloc_2sm(tmem_slot, Cfg::TMEM_COLS);
    tc_fence_before();
    __syncthreads();
    cluster_barrier();                                         // peer barriers initialised before any remote arrive / multicast
    tc_fence_after();
    const uint32_t tmem_base = *tmem_slot;

    const int ksl = p.Cin / Cfg::BK;
    const int kiters = p.taps * ksl;
    const int nchunks = (kiters + Cfg::CH - 1) / Cfg::CH;
    const int num_pairs = gridDim.x >> 1;
    const int pid = blockIdx.x >> 1;

    if (warp == 0) {
        if (lane == 0) {
            int s = 0;
            uint32_t ph = 0;
            for (int q = pid; q < total_q; q += num_pairs) {
                const int n0 = (q % n_tiles) * BN;
                const long long p0 = (long long)(q / n_tiles) * 256 + rank * 128;
                for (int it = 0; it < kiters; ++it) {
                    mbar_wait(&empty[s], ph ^ 1u, abort_flag, p.err, 0x500u + s);
                    const int tap = it / ksl;
                    const int ks = it - tap * ksl;
                    const int off = p.taps == 9 ? (tap / 3 - 1) * p.Wp + (tap % 3 - 1) : 0;
                    const int row = (int)(p0 + off);
                    uint8_t* st = smem + s * Cfg::STAGE_BYTES;
                    const uint32_t lbar = mapa_u32(smem_u32(&full[s]), 0u);
                    if (rank == 0) mbar_arrive_expect_tx(&full[s], 2 * Cfg::STAGE_BYTES);
                    const int kc = tap * p.Cin + ks * Cfg::BK;
                    const int nrow = n0 + (int)rank * (BN / 2);
                    tma_load_3d_2sm(st, &mapA, lbar, ks * Cfg::BK, row, 0);
                    tma_load_3d_2sm(st + Cfg::A_BYTES, &mapA, lbar, ks * Cfg::BK, row, 1);
                    tma_load_3d_2sm(st + 2 * Cfg::A_BYTES, &mapB, lbar, kc, nrow, 0);
                    tma_load_3d_2sm(st + 2 * Cfg::A_BYTES + Cfg::BH_BYTES, &mapB, lbar, kc, nrow, 1);
                    if (++s == Cfg::STAGES) { s = 0; ph ^= 1u; }
                }
            }
        }
        __syncwarp();
    } else if (warp == 1) {
        if (lane == 0 && rank == 0) {
            constexpr uint32_t idesc = umma_idesc_f16(256, BN);
            int s = 0, b = 0;
            uint32_t ph = 0, pht = 0;
            for (int q = pid; q < total_q; q += num_pairs) {
                for (int c = 0; c < nchunks; ++c) {
                    mbar_wait(&tempty[b], pht ^ 1u, abort_flag, p.err, 0x800u + b);   // both epilogues drained this buffer
                    tc_fence_after();
                    const uint32_t tacc = tmem_base + (uint32_t)(b * BN);
                    const int it_end = min(kiters, (c + 1) * Cfg::CH);
                    for (int it = c * Cfg::CH; it < it_end; ++it) {
                        mbar_wait(&full[s], ph, abort_flag, p.err, 0x600u + s);
                        tc_fence_after();
                        const uint32_t st = smem_u32(smem + s * Cfg::STAGE_BYTES);
                        const uint64_t a_hi = umma_desc_sw128(st);
                        const uint64_t a_lo = umma_desc_sw128(st + Cfg::A_BYTES);
                        const uint64_t b_hi = umma_desc_sw128(st + 2 * Cfg::A_BYTES);
                        const uint64_t b_lo = umma_desc_sw128(st + 2 * Cfg::A_BYTES + Cfg::BH_BYTES);
                        const bool first = (it == c * Cfg::CH);
#pragma unroll
                        for (int k = 0; k < Cfg::BK / 16; ++k) {
                            const uint64_t ko = (uint64_t)(k * 32 >> 4);
                            umma_f16_2sm(tacc, a_hi + ko, b_lo + ko, idesc, (first && k == 0) ? 0u : 1u);
                            umma_f16_2sm(tacc, a_lo + ko, b_hi + ko, idesc, 1u);
                            umma_f16_2sm(tacc, a_hi + ko, b_hi + ko, idesc, 1u);
                        }
                        umma_commit_2sm(&empty[s]);
                        if (++s == Cfg::STAGES) { s = 0; ph ^= 1u; }
                    }
                    umma_commit_2sm(&tfull[b]);
                    if (++b == Cfg::NBUF) { b = 0; pht ^= 1u; }
                }
            }
        }
        __syncwarp();
    } else {
        const int e = warp - 2;
        const int g = warp & 3;
        const int et = threadIdx.x - 64;
        constexpr int ETHREADS = 32 * Cfg::EPI_WARPS;
        const long long HpWp = (long long)p.Hp * p.Wp;
        const ActGeom go(p.N, p.H, p.W, p.Cout);
        const bool relu = (p.flags & WCTB200_RELU) != 0;
        int b = 0;
        uint32_t pht = 0;
        for (int q = pid; q < total_q; q += num_pairs) {
            const int n0 = (q % n_tiles) * BN;
            const long long p0 = (long long)(q / n_tiles) * 256 + rank * 128;
            asm volatile("bar.sync 1, %0;" ::"r"(ETHREADS) : "memory");
            for (int i = et; i < BN; i += ETHREADS) sbias[i] = p.bias ? p.bias[n0 + i] : 0.f;
            asm volatile("bar.sync 1, %0;" ::"r"(ETHREADS) : "memory");

            float acc[Cfg::NACC];
#pragma unroll
            for (int i = 0; i < Cfg::NACC; ++i) acc[i] = 0.f;
            for (int c = 0; c < nchunks; ++c) {
                mbar_wait(&tfull[b], pht, abort_flag, p.err, 0x700u + b);
                tc_fence_after();
                const uint32_t tsrc = tmem_base + ((uint32_t)(g * 32) << 16) + (uint32_t)(b * BN);
#pragma unroll
                for (int c0 = 0; c0 < Cfg::NACC; c0 += 64) {
                    uint32_t r0[32], r1[32];
                    tmem_ld32(tsrc + c0, r0);
                    tmem_ld32(tsrc + c0 + 32, r1);
                    tmem_ld_wait();
#pragma unroll
                    for (int j = 0; j < 32; ++j) acc[c0 + j] += __uint_as_float(r0[j]);
#pragma unroll
                    for (int j = 0; j < 32; ++j) acc[c0 + 32 + j] += __uint_as_float(r1[j]);
                }
                tc_fence_before();
                __syncwarp();
                if (lane == 0) mbar_arrive_remote(mapa_u32(smem_u32(&tempty[b]), 0u));   // the leader's barrier counts both CTAs
                if (++b == Cfg::NBUF) { b = 0; pht ^= 1u; }
            }
            const long long pos = p0 + g * 32 + lane;
            bool valid = pos < p.P;
            int n = 0, y = 0, x = 0;
            if (valid) {
                n = (int)(pos / HpWp);
                const int r = (int)(pos - n * HpWp);
                const int yy = r / p.Wp;
                const int xx = r - yy * p.Wp;
                valid = (yy >= 1) && (yy <= p.H) && (xx >= 1) && (xx <= p.W);
                y = yy - 1;
                x = xx - 1;
            }
            uint8_t* stg = aux + Cfg::AUX_BYTES + e * 8192;
            store_tile_rows<Cfg::NACC>(acc, sbias, relu, stg, lane, valid && !*abort_flag, n, y, x, p.out, go, n0);
        }
    }
    tc_fence_before();
    __syncthreads();
    cluster_barrier();                                          // no CTA exits (or frees TMEM) while its peer may still signal it
    if (warp == 1) tmem_dealloc_2sm(tmem_base, Cfg::TMEM_COLS);
}

// ---------------------------------------------------------------------------
// host side
// ---------------------------------------------------------------------------
typedef CUresult (*PFN_encodeTiled)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*,
                                    const cuuint64_t*, const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave,
                                    CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

static PFN_encodeTiled get_encode() {
    static PFN_encodeTiled fn = nullptr;
    if (!fn) {
        void* ptr = nullptr;
        cudaDriverEntryPointQueryResult q;
        if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &ptr, cudaEnableDefault, &q) == cudaSuccess &&
            q == cudaDriverEntryPointSuccess)
            fn = reinterpret_cast<PFN_encodeTiled>(ptr);
    }
    return fn;
}

// 3-D fp16 tensor map [d2][d1][d0] (d0 contiguous), box {64, box1, 1}, 128B swizzle, zero OOB fill
static int make_map(CUtensorMap* m, const void* base, uint64_t d0, uint64_t d1, uint64_t d2, uint64_t stride1_bytes,
                    uint64_t stride2_bytes, uint32_t box1) {
    PFN_encodeTiled enc = get_encode();
    if (!enc) {
        set_error("cuTensorMapEncodeTiled entry point not available");
        return WCTB200_ECUDA;
    }
    cuuint64_t dims[3] = {d0, d1, d2};
    cuuint64_t strides[2] = {stride1_bytes, stride2_bytes};
    cuuint32_t box[3] = {64, box1, 1};
    cuuint32_t estr[3] = {1, 1, 1};
    CUresult r = enc(m, CU_TENSOR_MAP_DATA_TYPE_FLOAT16, 3, const_cast<void*>(base), dims, strides, box, estr,
                     CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                     CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (r != CUDA_SUCCESS) {
        set_error("cuTensorMapEncodeTiled failed (%d): dims %llu %llu %llu strides %llu %llu box1 %u", (int)r,
                  (unsigned long long)d0, (unsigned long long)d1, (unsigned long long)d2,
                  (unsigned long long)stride1_bytes, (unsigned long long)stride2_bytes, box1);
        return WCTB200_ECUDA;
    }
    return 0;
}

template <int BN>
static int launch_bn(const CUtensorMap& mA, const CUtensorMap& mB, const ConvParams& p, dim3 grid, cudaStream_t st) {
    using Cfg = ConvCfg<BN>;
    static bool attr_done = false;
    if (!attr_done) {
        WCTB_CUDA(cudaFuncSetAttribute(conv_tc_kernel<BN>, cudaFuncAttributeMaxDynamicSharedMemorySize, Cfg::SMEM_BYTES));
        attr_done = true;
    }
    conv_tc_kernel<BN><<<grid, 192, Cfg::SMEM_BYTES, st>>>(mA, mB, p);
    WCTB_CHECK_LAUNCH("conv_tc_kernel");
    return 0;
}

int g_conv_oversub = 4;

template <int BN, bool FUSE_>
static int launch2_bn(const CUtensorMap& mA, const CUtensorMap& mB, const ConvParams& p, int total_tiles, int n_tiles,
                      cudaStream_t st) {
    using Cfg = Conv2Cfg<BN, FUSE_>;
    static bool attr_done = false;
    static int sms = 0;
    if (!attr_done) {
        WCTB_CUDA(cudaFuncSetAttribute(conv_tc2_kernel<BN, FUSE_>, cudaFuncAttributeMaxDynamicSharedMemorySize, Cfg::SMEM_BYTES));
        int dev = 0;
        WCTB_CUDA(cudaGetDevice(&dev));
        WCTB_CUDA(cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev));
        attr_done = true;
    }
    // Over-subscribed persistent grid: with g_conv_oversub x #SMs CTAs (1 resident per SM) the
    // hardware block scheduler hands queued CTAs to whichever SMs are free, so a conv launched
    // while the Jacobi clusters of the other stream hold half the SMs still balances its tiles
    // (a grid of exactly #SMs would run as two unbalanced waves).
    int grid = sms * (g_conv_oversub > 0 ? g_conv_oversub : 1);
    if (grid > total_tiles) grid = total_tiles;
    conv_tc2_kernel<BN, FUSE_><<<grid, Cfg::THREADS, Cfg::SMEM_BYTES, st>>>(mA, mB, p, total_tiles, n_tiles);
    WCTB_CHECK_LAUNCH("conv_tc2_kernel");
    return 0;
}

template <int BN>
static int launch5_bn(const CUtensorMap& mA, const CUtensorMap& mB, const ConvParams& p, int total_q, int n_tiles,
                      cudaStream_t st) {
    using Cfg = Conv5Cfg<BN>;
    static bool attr_done = false;
    static int sms = 0;
    if (!attr_done) {
        WCTB_CUDA(cudaFuncSetAttribute(conv_tc5_kernel<BN>, cudaFuncAttributeMaxDynamicSharedMemorySize, Cfg::SMEM_BYTES));
        int dev = 0;
        WCTB_CUDA(cudaGetDevice(&dev));
        WCTB_CUDA(cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev));
        attr_done = true;
    }
    int pairs = (sms / 2) * (g_conv_oversub > 0 ? g_conv_oversub : 1);
    if (pairs > total_q) pairs = total_q;
    cudaLaunchConfig_t cfg = {};
    cfg.gridDim = dim3((unsigned)(2 * pairs), 1, 1);
    cfg.blockDim = dim3(Cfg::THREADS, 1, 1);
    cfg.dynamicSmemBytes = Cfg::SMEM_BYTES;
    cfg.stream = st;
    cudaLaunchAttribute at[1];
    at[0].id = cudaLaunchAttributeClusterDimension;
    at[0].val.clusterDim.x = 2;
    at[0].val.clusterDim.y = 1;
    at[0].val.clusterDim.z = 1;
    cfg.attrs = at;
    cfg.numAttrs = 1;
    WCTB_CUDA(cudaLaunchKernelEx(&cfg, conv_tc5_kernel<BN>, mA, mB, p, total_q, n_tiles));
    return 0;
}

int g_conv_fuse = -1;        // -1 auto, 0 never, 1 whenever the tile allows (wctb200_debug_set_conv_fuse)
int g_conv_bn_override = 0;  // test/tuning hook: force the N tile (64/128/256)
int g_conv_impl = 2;         // 1 = one tile per CTA, all-TMEM accumulation; 2 = persistent + chunked register accumulation; 3 = 2 + tap reuse + multicast (conv_tc3.cu); 4 = aligned tap reuse (conv_tc4.cu); 5 = v2 only; 6 = v2 on CTA pairs (cta_group::2)

int launch_conv3x3_tc3(const __half* in, int N, int H, int W, int Cin, const __half* w_split, const float* bias, int Cout,
                       int flags, __half* out, int bn_override, cudaStream_t st);
int launch_conv3x3_tc4(const __half* in, int N, int H, int W, int Cin, const __half* w_split, const float* bias, int Cout,
                       int flags, __half* out, int bn_override, cudaStream_t st);
extern int g_conv4_cin_max;

int launch_conv3x3_tc(const __half* in, int N, int H, int W, int Cin, const __half* w_split, int taps, int nsets,
                      const float* bias, int Cout, int flags, __half* out, cudaStream_t st) {
    WCTB_REQUIRE(N >= 1 && H >= 2 && W >= 2, "conv3x3: bad geometry N=%d H=%d W=%d", N, H, W);
    WCTB_REQUIRE(Cin % 64 == 0 && Cout % 64 == 0 && Cin >= 64 && Cout >= 64, "conv3x3: Cin=%d Cout=%d must be multiples of 64", Cin, Cout);
    WCTB_REQUIRE(taps == 9 || taps == 1, "conv3x3: taps must be 9 or 1");
    WCTB_REQUIRE(nsets == 1 || nsets == N, "conv3x3: nsets must be 1 or N");
    ActGeom gi(N, H, W, Cin);
    WCTB_REQUIRE(gi.P < (1ll << 31) - 4096, "conv3x3: too many padded positions (%lld)", gi.P);

    if (g_conv_impl == 3 && taps == 9 && nsets == 1) {
        const int r = launch_conv3x3_tc3(in, N, H, W, Cin, w_split, bias, Cout, flags, out, g_conv_bn_override, st);
        if (r != 0) return r < 0 ? r : 0;      // 0 = shape not covered by v3 -> v2 below
    }
    // impl 4 = v4 for every shape it covers; impl 2 (default) sends only the L2-operand-bound layers (Cin <= 64) to v4
    if ((g_conv_impl == 4 || (g_conv_impl == 2 && Cin <= g_conv4_cin_max)) && taps == 9 && nsets == 1) {
        const int r = launch_conv3x3_tc4(in, N, H, W, Cin, w_split, bias, Cout, flags, out, g_conv_bn_override, st);
        if (r != 0) return r < 0 ? r : 0;      // 0 = shape not covered by v4 -> v2 below
    }
    if (g_conv_impl == 6 && taps == 9 && nsets == 1) {   // CTA pairs (cta_group::2)
        int BN5 = Cout % 128 == 0 ? 128 : 64;
        if (g_conv_bn_override == 64) BN5 = 64;
        CUtensorMap mA5, mB5;
        int rc5 = make_map(&mA5, in, (uint64_t)Cin, (uint64_t)gi.P, 2, (uint64_t)Cin * 2, (uint64_t)gi.plane * 2, 128);
        if (rc5) return rc5;
        const uint64_t K5 = (uint64_t)9 * Cin;
        rc5 = make_map(&mB5, w_split, K5, (uint64_t)Cout, 2, K5 * 2, K5 * Cout * 2, (uint32_t)(BN5 / 2));
        if (rc5) return rc5;
        ConvParams p5;
        p5.N = N; p5.H = H; p5.W = W; p5.Cin = Cin; p5.Cout = Cout; p5.Hp = gi.Hp; p5.Wp = gi.Wp; p5.P = gi.P;
        p5.taps = 9; p5.nsets = 1; p5.per_image = 0; p5.tiles_per_image = 0;
        p5.flags = flags; p5.bias = bias; p5.out = out; p5.err = device_error_word();
        const int n_tiles5 = Cout / BN5;
        const int total_q = (int)cdiv(gi.P, 256) * n_tiles5;
        return BN5 == 128 ? launch5_bn<128>(mA5, mB5, p5, total_q, n_tiles5, st) : launch5_bn<64>(mA5, mB5, p5, total_q, n_tiles5, st);
    }
    int BN = Cout % 256 == 0 ? 256 : (Cout % 128 == 0 ? 128 : 64);
    if (BN == 256) BN = 128;  // v1 default: 3-stage pipeline beats the 2-stage 256-wide tile until 2-CTA lands
    if (g_conv_bn_override && Cout % g_conv_bn_override == 0) BN = g_conv_bn_override;

    CUtensorMap mA, mB;
    int rc = make_map(&mA, in, (uint64_t)Cin, (uint64_t)gi.P, 2, (uint64_t)Cin * 2, (uint64_t)gi.plane * 2, 128);
    if (rc) return rc;
    const uint64_t K = (uint64_t)taps * Cin;
    rc = make_map(&mB, w_split, K, (uint64_t)Cout, (uint64_t)2 * nsets, K * 2, K * Cout * 2, (uint32_t)BN);
    if (rc) return rc;

    ConvParams p;
    p.N = N; p.H = H; p.W = W; p.Cin = Cin; p.Cout = Cout; p.Hp = gi.Hp; p.Wp = gi.Wp; p.P = gi.P;
    p.taps = taps;
    p.nsets = nsets;
    p.per_image = nsets > 1 ? 1 : 0;
    p.tiles_per_image = cdiv((long long)gi.Hp * gi.Wp, 128);
    p.flags = flags;
    p.bias = bias;
    p.out = out;
    p.err = device_error_word();
    dim3 grid(p.per_image ? (unsigned)(N * p.tiles_per_image) : (unsigned)cdiv(gi.P, 128), (unsigned)(Cout / BN));
    if (g_conv_impl >= 2) {
        const int n_tiles = Cout / BN;
        const int total = (int)grid.x * n_tiles;
        switch (BN) {
            // fused [b_hi|b_lo] MMAs (Conv2Cfg::FUSE): always at BN=64 (4 TMEM buffers stay); at BN=128 the ring shrinks to 2
            // buffers, which only pays for long K loops (measured: Cin=128 -9 %, Cin>=256 +3..5 %)
            case 64: return g_conv_fuse == 0 ? launch2_bn<64, false>(mA, mB, p, total, n_tiles, st)
                                             : launch2_bn<64, true>(mA, mB, p, total, n_tiles, st);
            case 128: return (g_conv_fuse == 1 || (g_conv_fuse < 0 && (long long)taps * Cin >= 9 * 256))
                                 ? launch2_bn<128, true>(mA, mB, p, total, n_tiles, st)
                                 : launch2_bn<128, false>(mA, mB, p, total, n_tiles, st);
            default: return launch2_bn<256, false>(mA, mB, p, total, n_tiles, st);
        }
    }
    switch (BN) {
        case 64: return launch_bn<64>(mA, mB, p, grid, st);
        case 128: return launch_bn<128>(mA, mB, p, grid, st);
        default: return launch_bn<256>(mA, mB, p, grid, st);
    }
}

}  // namespace wctb
