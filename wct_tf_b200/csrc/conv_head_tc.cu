// Encoder head on the tensor cores: (1x1 preprocess conv folded into) conv1_1 3 -> 64 + ReLU  (vgg_normalised.py:25-40),
// fp32 image in, SPF16 activation (interior + reflect halo) out.
//
// K = 27 is too short for a TMA-fed implicit GEMM (the operand would be 9 boxes of 6 bytes), so the A operand is BUILT in
// shared memory by four producer warps: thread = pixel, its 3x3x3 reflect-indexed neighbourhood (27 fp32, L1 hits) is scaled
// by 256 (the [0,1] image would put the lo halves into the fp16 subnormals), split into fp16 hi/lo and written as ONE
// 128-byte K-major row  [hi k=0..31 | lo k=0..31]  in the 128B-swizzled layout the MMA expects.  With
//     B rows  0..63  = [w_hi | w_hi]        ->  D[:,  0:64 ) = x_hi w_hi + x_lo w_hi
//     B rows 64..127 = [w_lo |  0  ]        ->  D[:, 64:128) = x_hi w_lo
// one tcgen05.mma (M = 128, N = 128) per 16-wide k-step -- 4 per 128-pixel tile -- computes all three split products.
//   warp 0    : MMA issuer, owns the TMEM allocation (2 buffers x 128 columns)
//   warps 1-4 : producers (global -> split fp16 -> swizzled smem, fence.proxy.async, mbarrier)
//   warps 5-8 : epilogue (tcgen05.ld, undo the scales, bias, ReLU, split, coalesced stores incl. the halo cells)
// Two CTAs per SM.  The SIMT kernel (layers.cu k_conv_head) spent 1728 FMAs per pixel and ran at 31 TFLOP/s; this one is
// bound by its 258 B/pixel of output.
#include "common.cuh"

namespace wctb {

struct HeadCfg {
    static constexpr int A_BYTES = 128 * 128;              // one tile: 128 pixels x 128 B
    static constexpr int STAGES = 2;
    static constexpr int B_BYTES = 128 * 128;              // 128 rows (n) x 64 fp16 (k)
    static constexpr int STG_BYTES = 4 * 8192;
    static constexpr int AUX_BYTES = 512;                  // barriers + bias[64]
    static constexpr int SMEM_BYTES = STAGES * A_BYTES + B_BYTES + STG_BYTES + AUX_BYTES + 1024;
    static constexpr int THREADS = 288;
    static constexpr int TMEM_COLS = 256;
};

struct HeadParams {
    int N, H, W;
    long long total_px;
    int tiles;
    const float* img;
    const float* w;        // [27][64], k = (ky*3 + kx)*3 + ci
    const float* bias;     // [64]
    __half* out;
    unsigned int* err;
};

__device__ __forceinline__ int reflect1(int i, int n) { return i < 0 ? -i : (i >= n ? 2 * n - 2 - i : i); }

__global__ void __launch_bounds__(HeadCfg::THREADS, 2)
conv_head_tc_kernel(const HeadParams p) {
    using Cfg = HeadCfg;
    extern __shared__ uint8_t smem_raw[];
    uint8_t* smem = smem_raw + ((1024u - (smem_u32(smem_raw) & 1023u)) & 1023u);
    uint8_t* bsm = smem + Cfg::STAGES * Cfg::A_BYTES;
    uint8_t* stg = bsm + Cfg::B_BYTES;
    uint8_t* aux = stg + Cfg::STG_BYTES;
    uint64_t* full = reinterpret_cast<uint64_t*>(aux);
    uint64_t* empty = full + Cfg::STAGES;
    uint64_t* tfull = empty + Cfg::STAGES;
    uint64_t* tempty = tfull + 2;
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(tempty + 2);
    volatile int* abort_flag = reinterpret_cast<volatile int*>(tmem_slot + 1);
    float* s_oscale = reinterpret_cast<float*>(tmem_slot + 2);
    float* sbias = reinterpret_cast<float*>(aux + 256);

    const int warp = threadIdx.x >> 5;
    const int lane = threadIdx.x & 31;
    __shared__ unsigned int s_prev_err;
    __shared__ float s_red[16];
    if (threadIdx.x == 0) s_prev_err = *reinterpret_cast<volatile unsigned int*>(p.err);
    __syncthreads();
    if (s_prev_err != 0u) return;

    if (threadIdx.x == 0) {
        for (int s = 0; s < Cfg::STAGES; ++s) {
            mbar_init(&full[s], 4);          // one arrival per producer warp
            mbar_init(&empty[s], 1);
        }
        for (int b = 0; b < 2; ++b) {
            mbar_init(&tfull[b], 1);
            mbar_init(&tempty[b], 4);
        }
        *abort_flag = 0;
        fence_barrier_init();
    }
    if (warp == 0) tmem_alloc(tmem_slot, Cfg::TMEM_COLS);

    // ---- B operand (all threads): power-of-two weight scale, split, swizzled K-major rows ----
    float m = 0.f;
    for (int i = threadIdx.x; i < 27 * 64; i += Cfg::THREADS) m = fmaxf(m, fabsf(p.w[i]));
    for (int o = 16; o > 0; o >>= 1) m = fmaxf(m, __shfl_xor_sync(0xffffffffu, m, o));
    if (lane == 0) s_red[warp] = m;
    if (threadIdx.x < 64) sbias[threadIdx.x] = p.bias[threadIdx.x];
    __syncthreads();
    float wsc = 1.f;
    {
        float mm = 0.f;
        for (int i = 0; i < Cfg::THREADS / 32; ++i) mm = fmaxf(mm, s_red[i]);
        if (mm > 0.f && isfinite(mm)) {
            int e = 0;
            frexpf(mm, &e);
            wsc = ldexpf(1.f, 10 - e);                     // max |w| * wsc in [512, 1024)
        }
    }
    if (threadIdx.x == 0) *s_oscale = 1.f / (wsc * 256.f);  // undoes the weight scale and the x256 of the image
    for (int i = threadIdx.x; i < 128 * 8; i += Cfg::THREADS) {
        const int row = i >> 3, c = i & 7;                 // 16-byte chunk c of B row `row`
        const int n = row & 63;
        const bool lo_row = row >= 64;
        Half8 v;
#pragma unroll
        for (int j = 0; j < 4; ++j) {
            __half h2[2];
#pragma unroll
            for (int t = 0; t < 2; ++t) {
                const int k = (c & 3) * 8 + j * 2 + t;
                float x = k < 27 ? p.w[k * 64 + n] * wsc : 0.f;
                __half hi, lo;
                split_f32(x, hi, lo);
                h2[t] = lo_row ? (c < 4 ? lo : __float2half(0.f)) : hi;
            }
            v.v[j] = __halves2half2(h2[0], h2[1]);
        }
        *reinterpret_cast<Half8*>(bsm + row * 128 + ((c ^ (row & 7)) * 16)) = v;
    }
    fence_proxy_async();
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem_base = *tmem_slot;
    const unsigned int HW = (unsigned int)p.H * (unsigned int)p.W;

    if (warp == 0) {
        if (lane == 0) {
            constexpr uint32_t idesc = umma_idesc_f16(128, 128);
            const uint64_t b_desc = umma_desc_sw128(smem_u32(bsm));
            uint32_t tg = 0;
            for (int tile = blockIdx.x; tile < p.tiles; tile += gridDim.x, ++tg) {
                const int s = tg % Cfg::STAGES, b = tg & 1;
                mbar_wait(&tempty[b], ((tg >> 1) & 1) ^ 1u, abort_flag, p.err, 0x420u + b);
                mbar_wait(&full[s], (tg / Cfg::STAGES) & 1, abort_flag, p.err, 0x230u + s);
                tc_fence_after();
                const uint64_t a_desc = umma_desc_sw128(smem_u32(smem + s * Cfg::A_BYTES));
                const uint32_t tacc = tmem_base + (uint32_t)(b * 128);
#pragma unroll
                for (int k = 0; k < 4; ++k) {
                    const uint64_t ko = (uint64_t)(k * 32 >> 4);
                    umma_f16(tacc, a_desc + ko, b_desc + ko, idesc, k == 0 ? 0u : 1u);
                }
                umma_commit(&empty[s]);
                umma_commit(&tfull[b]);
            }
        }
        __syncwarp();
    } else if (warp <= 4) {
        // ---- producers: one pixel per thread ----
        const int i = (warp - 1) * 32 + lane;              // row of the A tile
        uint32_t tg = 0;
        for (int tile = blockIdx.x; tile < p.tiles; tile += gridDim.x, ++tg) {
            const int s = tg % Cfg::STAGES;
            mbar_wait(&empty[s], ((tg / Cfg::STAGES) & 1) ^ 1u, abort_flag, p.err, 0x120u + s);
            const long long q = (long long)tile * 128 + i;
            float v[32];
#pragma unroll
            for (int k = 27; k < 32; ++k) v[k] = 0.f;
            if (q < p.total_px) {
                const unsigned int n = (unsigned int)(q / HW);
                const unsigned int r = (unsigned int)(q - (long long)n * HW);
                const int y = (int)(r / (unsigned int)p.W);
                const int x = (int)(r - (unsigned int)y * (unsigned int)p.W);
                const float* base = p.img + (long long)n * HW * 3;
#pragma unroll
                for (int ky = 0; ky < 3; ++ky) {
                    const float* row = base + (long long)reflect1(y + ky - 1, p.H) * p.W * 3;
#pragma unroll
                    for (int kx = 0; kx < 3; ++kx) {
                        const float* px = row + reflect1(x + kx - 1, p.W) * 3;
#pragma unroll
                        for (int ci = 0; ci < 3; ++ci) v[(ky * 3 + kx) * 3 + ci] = __ldg(px + ci) * 256.f;
                    }
                }
            } else {
#pragma unroll
                for (int k = 0; k < 27; ++k) v[k] = 0.f;
            }
            uint8_t* rowp = smem + s * Cfg::A_BYTES + i * 128;
#pragma unroll
            for (int c = 0; c < 4; ++c) {
                Half8 hi, lo;
                split8(v + c * 8, hi, lo);
                *reinterpret_cast<Half8*>(rowp + ((c ^ (i & 7)) * 16)) = hi;
                *reinterpret_cast<Half8*>(rowp + (((c + 4) ^ (i & 7)) * 16)) = lo;
            }
            fence_proxy_async();
            __syncwarp();
            if (lane == 0) mbar_arrive(&full[s]);
        }
    } else {
        // ---- epilogue ----
        const int g = warp & 3;                            // TMEM lane quadrant (warps 5,6,7,8 -> 1,2,3,0)
        const ActGeom go(p.N, p.H, p.W, 64);
        const float osc = *s_oscale;
        uint8_t* wstg = stg + (warp - 5) * 8192;
        uint32_t tg = 0;
        for (int tile = blockIdx.x; tile < p.tiles; tile += gridDim.x, ++tg) {
            const int b = tg & 1;
            mbar_wait(&tfull[b], (tg >> 1) & 1, abort_flag, p.err, 0x320u + b);
            tc_fence_after();
            const uint32_t tsrc = tmem_base + ((uint32_t)(g * 32) << 16) + (uint32_t)(b * 128);
            float acc[64];
#pragma unroll
            for (int c0 = 0; c0 < 64; c0 += 16) {           // 16-column loads: 2 CTAs/SM leave ~110 registers per thread
                uint32_t r0[16], r1[16];
                tmem_ld16(tsrc + c0, r0);
                tmem_ld16(tsrc + 64 + c0, r1);
                tmem_ld_wait();
#pragma unroll
                for (int j = 0; j < 16; ++j) acc[c0 + j] = __uint_as_float(r0[j]) + __uint_as_float(r1[j]);
            }
            tc_fence_before();
            __syncwarp();
            if (lane == 0) mbar_arrive(&tempty[b]);
            const long long q = (long long)tile * 128 + g * 32 + lane;
            int flags = -1;
            unsigned int ppos = 0;
            if (q < p.total_px && !*abort_flag) {
                const unsigned int n = (unsigned int)(q / HW);
                const unsigned int r = (unsigned int)(q - (long long)n * HW);
                const int y = (int)(r / (unsigned int)p.W);
                const int x = (int)(r - (unsigned int)y * (unsigned int)p.W);
                flags = halo_flags(go, y, x);
                ppos = (n * (unsigned int)go.Hp + (unsigned int)(y + 1)) * (unsigned int)go.Wp + (unsigned int)(x + 1);
            }
            store_tile_rows<64>(acc, osc, sbias, true, wstg, lane, ppos, flags, p.out, go, 0);
        }
    }
    tc_fence_before();
    __syncthreads();
    if (warp == 0) tmem_dealloc(tmem_base, Cfg::TMEM_COLS);
}

int g_conv_head_tc = 1;          // 0: the SIMT kernel of layers.cu (wctb200_debug_set_conv_head_tc)

// returns 1 when not handled here (caller falls through to the SIMT kernel), 0 on success, < 0 on error
int launch_conv_head_tc(const float* img, int N, int H, int W, const float* w, const float* b, __half* out, cudaStream_t st) {
    using Cfg = HeadCfg;
    const ActGeom go(N, H, W, 64);
    if (!g_conv_head_tc || H < 2 || W < 2 || go.P >= (1ll << 31)) return 1;
    HeadParams p;
    p.N = N; p.H = H; p.W = W;
    p.total_px = (long long)N * H * W;
    p.tiles = cdiv(p.total_px, 128);
    p.img = img;
    p.w = w;
    p.bias = b;
    p.out = out;
    p.err = device_error_word();
    WCTB_ENSURE_SMEM(conv_head_tc_kernel, Cfg::SMEM_BYTES);
    int grid = device_sm_count() * 2;
    if (grid > p.tiles) grid = p.tiles;
    conv_head_tc_kernel<<<grid, Cfg::THREADS, Cfg::SMEM_BYTES, st>>>(p);
    WCTB_CHECK_LAUNCH("conv_head_tc_kernel");
    return 0;
}

}  // namespace wctb
