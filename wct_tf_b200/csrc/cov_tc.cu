// Covariance GEMM  S = sum_p x_p x_p^T  (C x HW . HW x C, ops.py:45,50,108,121) on the tensor cores.
//
// The features are SPF16 [pixel][channel]: for the product X^T X the contraction index K is the
// PIXEL, so both operands are "MN-major" (for a fixed k the 64 channels of a slice are the 128
// contiguous bytes of a row).  A TMA box of 64 pixels x 64 channels lands as the canonical
// MN-major SWIZZLE_128B tile (8-pixel groups of 1024 B), and tcgen05.mma reads A and B from the
// same staged tiles with a_major = b_major = MN.  The tensor map is built over the INTERIOR of the
// reflect-padded plane (base pointer at pixel (0,0), padded strides, dims W x H): boxes that stick
// out at a ragged edge are zero-filled by TMA and add nothing to the sums, halo cells are never read.
//
//   The input is the CENTRED feature copy fc = x - mean written by k_center (wct.cu): products
//   in fp32 (split-fp16 x3), drained from TMEM every 128 pixels into registers (the tensor core
//   accumulates with truncation and the diagonals are all-positive sums), per-CTA partials
//   combined with fp64 atomics.  (An uncentred variant with the HW m m^T term removed in fp64 was
//   measured first: it cancels in fp32 and produced a spurious eigenvalue above the 1e-5 cut on a
//   rank-deficient map, so centring happens before the product, like ops.py:44-45.)
//
// One CTA = one 128 x 128 block pair (bi <= bj) of the C x C matrix x one range of pixel tiles.
#include "common.cuh"

namespace wctb {

struct CovParams {
    int C, W, H, N;
    int tiles_x, tiles_y;        // 32 x 2 pixel tiles per image
    int ksplit, tiles_per_split;
    int nb;                      // 128-channel blocks (C=64: 1 block, rows duplicated)
    int lbo_bytes, sbo_bytes;    // MN-major descriptor strides (probe knobs)
    double* cov;                 // [N][C][C] fp64 partial sums (upper block triangle)
    unsigned int* err;
};

struct CovCfg {
    static constexpr int SLICE = 64 * 128;                  // 64 pixels x 64 channels fp16 = 8 KB
    static constexpr int OPER = 4 * SLICE;                  // 2 channel slices x 2 planes
    static constexpr int STAGE = 2 * OPER;                  // A + B
    static constexpr int STAGES = 3;
    static constexpr int NBUF = 4;
    static constexpr int CH = 2;                            // 64-pixel tiles per TMEM accumulation chunk (24 truncating adds)
    static constexpr int THREADS = 192;
    static constexpr int SMEM_BYTES = STAGES * STAGE + 512 + 1024;
};

__device__ __forceinline__ void tma_load_5d_cov(void* smem_dst, const void* map, uint64_t* bar, int c0, int c1, int c2,
                                                int c3, int c4) {
    asm volatile(
        "cp.async.bulk.tensor.5d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5, %6, %7}], [%2];"
        ::"r"(smem_u32(smem_dst)), "l"(reinterpret_cast<uint64_t>(map)), "r"(smem_u32(bar)), "r"(c0), "r"(c1), "r"(c2),
        "r"(c3), "r"(c4)
        : "memory");
}

// MN-major SWIZZLE_128B operand: 64 MN elements (128 B) per row, rows = K; LBO = stride between
// 64-wide MN groups, SBO = stride between 8-row K groups (cute::UMMA canonical layout
// ((8,n),(8,k)):((1,LBO),(8,SBO)) in 16-byte units)
__device__ __forceinline__ uint64_t umma_desc_mn_sw128(uint32_t smem_addr, int lbo_bytes, int sbo_bytes) {
    uint64_t d = 0;
    d |= (uint64_t)((smem_addr >> 4) & 0x3FFF);
    d |= (uint64_t)((lbo_bytes >> 4) & 0x3FFF) << 16;
    d |= (uint64_t)((sbo_bytes >> 4) & 0x3FFF) << 32;
    d |= (uint64_t)1 << 46;
    d |= (uint64_t)2 << 61;
    return d;
}
__host__ __device__ constexpr uint32_t umma_idesc_f16_mn(int M, int N) {
    return (1u << 4) | (1u << 15) | (1u << 16) | ((uint32_t)(N >> 3) << 17) | ((uint32_t)(M >> 4) << 24);
}

__global__ void __launch_bounds__(CovCfg::THREADS, 1)
cov_tc_kernel(const __grid_constant__ CUtensorMap mapX, const CovParams p) {
    using Cfg = CovCfg;
    extern __shared__ uint8_t smem_raw[];
    uint8_t* smem = smem_raw + ((1024u - (smem_u32(smem_raw) & 1023u)) & 1023u);
    uint8_t* aux = smem + Cfg::STAGES * Cfg::STAGE;
    uint64_t* full = reinterpret_cast<uint64_t*>(aux);
    uint64_t* empty = full + Cfg::STAGES;
    uint64_t* tfull = empty + Cfg::STAGES;
    uint64_t* tempty = tfull + Cfg::NBUF;
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(tempty + Cfg::NBUF);
    volatile int* abort_flag = reinterpret_cast<volatile int*>(tmem_slot + 1);

    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    // block pair / k-split / image of this CTA
    const int npairs = p.nb * (p.nb + 1) / 2;
    const int pair = blockIdx.x % npairs;
    const int split = blockIdx.x / npairs;
    int bi = 0, rem = pair;
    while (rem >= p.nb - bi) { rem -= p.nb - bi; ++bi; }
    const int bj = bi + rem;
    const bool diag = (bi == bj);
    const int img = blockIdx.y;
    const int tiles_img = p.tiles_x * p.tiles_y;
    const int t0 = split * p.tiles_per_split;
    const int t1 = min(t0 + p.tiles_per_split, tiles_img);
    const int ntiles = max(t1 - t0, 0);
    // C = 64: a single 64-channel slice.  The 128-row operand is [hi plane | lo plane] of that slice and is used as BOTH
    // A and B: ONE MMA per 16 pixels yields hi.hi, hi.lo, lo.hi (and lo.lo) as the four 64x64 quadrants of the 128x128
    // accumulator; the epilogue adds the quadrants.  (The first version loaded the slice twice to fill M = N = 128 and
    // issued the three split products separately: 3x the MMAs and 2x the TMA bytes for the same result.)
    const bool dup = (p.C == 64);

    if (threadIdx.x == 0) {
        for (int s = 0; s < Cfg::STAGES; ++s) { mbar_init(&full[s], 1); mbar_init(&empty[s], 1); }
        for (int b = 0; b < Cfg::NBUF; ++b) { mbar_init(&tfull[b], 1); mbar_init(&tempty[b], 4); }
        *abort_flag = 0;
        fence_barrier_init();
        tma_prefetch_desc(&mapX);
    }
    if (warp == 1) tmem_alloc(tmem_slot, 512);
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem_base = *tmem_slot;
    const int nchunks = (ntiles + Cfg::CH - 1) / Cfg::CH;
    const uint32_t stage_bytes = dup ? 2 * Cfg::SLICE : (diag ? Cfg::OPER : Cfg::STAGE);

    if (warp == 0) {
        if (lane == 0) {
            for (int it = 0; it < ntiles; ++it) {
                const int s = it % Cfg::STAGES;
                mbar_wait(&empty[s], ((it / Cfg::STAGES) & 1) ^ 1u, abort_flag, p.err, 0x510u + s);
                const int t = t0 + it;
                const int ty = t / p.tiles_x, tx = t - ty * p.tiles_x;
                uint8_t* st = smem + s * Cfg::STAGE;
                mbar_arrive_expect_tx(&full[s], stage_bytes);
                if (dup) {
                    tma_load_5d_cov(st, &mapX, &full[s], 0, tx * 32, ty * 2, img, 0);                  // rows 0..63   : hi
                    tma_load_5d_cov(st + Cfg::SLICE, &mapX, &full[s], 0, tx * 32, ty * 2, img, 1);     // rows 64..127 : lo
                    continue;
                }
                for (int op = 0; op < (diag ? 1 : 2); ++op) {
                    const int blk = op == 0 ? bi : bj;
                    for (int sl = 0; sl < 2; ++sl) {
                        const int ch = blk * 128 + sl * 64;
                        for (int pl = 0; pl < 2; ++pl)
                            tma_load_5d_cov(st + op * Cfg::OPER + (pl * 2 + sl) * Cfg::SLICE, &mapX, &full[s], ch, tx * 32,
                                            ty * 2, img, pl);
                    }
                }
            }
        }
        __syncwarp();
    } else if (warp == 1) {
        if (lane == 0) {
            constexpr uint32_t idesc = umma_idesc_f16_mn(128, 128);
            int it = 0;
            for (int c = 0; c < nchunks; ++c) {
                const int b = c % Cfg::NBUF;
                mbar_wait(&tempty[b], ((c / Cfg::NBUF) & 1) ^ 1u, abort_flag, p.err, 0x540u + b);
                tc_fence_after();
                const uint32_t tacc = tmem_base + (uint32_t)(b * 128);
                const int it_end = min(ntiles, (c + 1) * Cfg::CH);
                for (; it < it_end; ++it) {
                    const int s = it % Cfg::STAGES;
                    mbar_wait(&full[s], (it / Cfg::STAGES) & 1, abort_flag, p.err, 0x520u + s);
                    tc_fence_after();
                    const uint32_t a0 = smem_u32(smem + s * Cfg::STAGE);
                    const uint32_t b0 = diag ? a0 : a0 + Cfg::OPER;
                    const bool first = (it == c * Cfg::CH);
#pragma unroll
                    for (int k = 0; k < 4; ++k) {           // 64 pixels = 4 x UMMA_K(16): 16 rows of 128 B = 2048 B per step
                        const uint32_t ko = k * 2048;
                        if (dup) {
                            const uint64_t d = umma_desc_mn_sw128(a0 + ko, p.lbo_bytes, p.sbo_bytes);   // [hi | lo] x [hi | lo]^T
                            umma_f16(tacc, d, d, idesc, (first && k == 0) ? 0u : 1u);
                            continue;
                        }
                        const uint64_t a_hi = umma_desc_mn_sw128(a0 + ko, p.lbo_bytes, p.sbo_bytes);
                        const uint64_t a_lo = umma_desc_mn_sw128(a0 + 2 * Cfg::SLICE + ko, p.lbo_bytes, p.sbo_bytes);
                        const uint64_t b_hi = umma_desc_mn_sw128(b0 + ko, p.lbo_bytes, p.sbo_bytes);
                        const uint64_t b_lo = umma_desc_mn_sw128(b0 + 2 * Cfg::SLICE + ko, p.lbo_bytes, p.sbo_bytes);
                        umma_f16(tacc, a_hi, b_lo, idesc, (first && k == 0) ? 0u : 1u);
                        umma_f16(tacc, a_lo, b_hi, idesc, 1u);
                        umma_f16(tacc, a_hi, b_hi, idesc, 1u);
                    }
                    umma_commit(&empty[s]);
                }
                umma_commit(&tfull[b]);
            }
        }
        __syncwarp();
    } else {
        const int g = warp & 3;
        float acc[128];
#pragma unroll
        for (int i = 0; i < 128; ++i) acc[i] = 0.f;
        for (int c = 0; c < nchunks; ++c) {
            const int b = c % Cfg::NBUF;
            mbar_wait(&tfull[b], (c / Cfg::NBUF) & 1, abort_flag, p.err, 0x530u + b);
            tc_fence_after();
            const uint32_t tsrc = tmem_base + ((uint32_t)(g * 32) << 16) + (uint32_t)(b * 128);
#pragma unroll
            for (int c0 = 0; c0 < 128; c0 += 64) {
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
            if (lane == 0) mbar_arrive(&tempty[b]);
        }
        // this CTA's partial block -> fp64 accumulation buffer (row = channel bi*128 + m)
        const int m = g * 32 + lane;
        if (ntiles > 0 && !*abort_flag) {
            if (dup) {
                // accumulator rows m and m+64 both belong to channel m & 63; columns j and 64+j to channel j
                double* dst = p.cov + ((long long)img * p.C + (m & 63)) * p.C;
#pragma unroll 8
                for (int j = 0; j < 64; ++j) atomicAdd(dst + j, (double)acc[j] + (double)acc[64 + j]);
            } else {
                double* dst = p.cov + ((long long)img * p.C + bi * 128 + m) * p.C + bj * 128;
#pragma unroll 8
                for (int j = 0; j < 128; ++j) atomicAdd(dst + j, (double)acc[j]);
            }
        }
    }
    tc_fence_before();
    __syncthreads();
    if (warp == 1) tmem_dealloc(tmem_base, 512);
}

// ---------------------------------------------------------------------------
typedef CUresult (*PFN_encodeTiledC)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*,
                                     const cuuint64_t*, const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave,
                                     CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);


int launch_cov_tc(const __half* act, ActGeom g, double* cov, cudaStream_t st) {
    static PFN_encodeTiledC enc = nullptr;
    if (!enc) {
        void* ptr = nullptr;
        cudaDriverEntryPointQueryResult q;
        if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &ptr, cudaEnableDefault, &q) != cudaSuccess ||
            q != cudaDriverEntryPointSuccess) {
            set_error("cuTensorMapEncodeTiled entry point not available");
            return WCTB200_ECUDA;
        }
        enc = reinterpret_cast<PFN_encodeTiledC>(ptr);
    }
    WCTB_REQUIRE(g.C == 64 || g.C % 128 == 0, "cov_tc: C=%d must be 64 or a multiple of 128", g.C);
    CovParams p;
    p.C = g.C; p.W = g.W; p.H = g.H; p.N = g.N;
    p.tiles_x = (g.W + 31) / 32;
    p.tiles_y = (g.H + 1) / 2;
    p.nb = g.C == 64 ? 1 : g.C / 128;
    const int npairs = p.nb * (p.nb + 1) / 2;
    const int tiles_img = p.tiles_x * p.tiles_y;
    // aim for ~2 CTAs per SM in total (each CTA pays ~5 us of TMEM/barrier set-up and final atomics),
    // at least 8 tiles (512 pixels) per CTA
    int ksplit = (device_sm_count() * 2 + npairs * g.N - 1) / (npairs * g.N);
    if (ksplit < 1) ksplit = 1;
    int tps = (tiles_img + ksplit - 1) / ksplit;
    if (tps < 8) tps = 8;
    ksplit = (tiles_img + tps - 1) / tps;
    p.ksplit = ksplit;
    p.tiles_per_split = tps;
    p.lbo_bytes = 8192;      // MN-major descriptor strides, probed on B200 (profiles/r01_cov_mn_major_probe.txt)
    p.sbo_bytes = 1024;
    p.cov = cov;
    p.err = device_error_word();
    // tensor map over the interior pixels only (see header comment)
    CUtensorMap mX;
    const __half* base = act + ((long long)g.Wp + 1) * g.C;
    cuuint64_t dims[5] = {(cuuint64_t)g.C, (cuuint64_t)g.W, (cuuint64_t)g.H, (cuuint64_t)g.N, 2};
    cuuint64_t strides[4] = {(cuuint64_t)g.C * 2, (cuuint64_t)g.Wp * g.C * 2, (cuuint64_t)g.Hp * g.Wp * g.C * 2,
                             (cuuint64_t)g.plane * 2};
    cuuint32_t box[5] = {64, 32, 2, 1, 1};
    cuuint32_t estr[5] = {1, 1, 1, 1, 1};
    CUresult r = enc(&mX, CU_TENSOR_MAP_DATA_TYPE_FLOAT16, 5, const_cast<__half*>(base), dims, strides, box, estr,
                     CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                     CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (r != CUDA_SUCCESS) {
        set_error("cov_tc: cuTensorMapEncodeTiled failed (%d)", (int)r);
        return WCTB200_ECUDA;
    }
    WCTB_ENSURE_SMEM(cov_tc_kernel, CovCfg::SMEM_BYTES);
    dim3 grid((unsigned)(npairs * ksplit), (unsigned)g.N);
    cov_tc_kernel<<<grid, CovCfg::THREADS, CovCfg::SMEM_BYTES, st>>>(mX, p);
    WCTB_CHECK_LAUNCH("cov_tc_kernel");
    return 0;
}

}  // namespace wctb
