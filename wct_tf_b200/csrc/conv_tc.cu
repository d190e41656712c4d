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

int g_conv_bn_override = 0;  // test/tuning hook: force the N tile (64/128/256)

int launch_conv3x3_tc(const __half* in, int N, int H, int W, int Cin, const __half* w_split, int taps, int nsets,
                      const float* bias, int Cout, int flags, __half* out, cudaStream_t st) {
    WCTB_REQUIRE(N >= 1 && H >= 2 && W >= 2, "conv3x3: bad geometry N=%d H=%d W=%d", N, H, W);
    WCTB_REQUIRE(Cin % 64 == 0 && Cout % 64 == 0 && Cin >= 64 && Cout >= 64, "conv3x3: Cin=%d Cout=%d must be multiples of 64", Cin, Cout);
    WCTB_REQUIRE(taps == 9 || taps == 1, "conv3x3: taps must be 9 or 1");
    WCTB_REQUIRE(nsets == 1 || nsets == N, "conv3x3: nsets must be 1 or N");
    ActGeom gi(N, H, W, Cin);
    WCTB_REQUIRE(gi.P < (1ll << 31) - 4096, "conv3x3: too many padded positions (%lld)", gi.P);

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
    switch (BN) {
        case 64: return launch_bn<64>(mA, mB, p, grid, st);
        case 128: return launch_bn<128>(mA, mB, p, grid, st);
        default: return launch_bn<256>(mA, mB, p, grid, st);
    }
}

}  // namespace wctb
