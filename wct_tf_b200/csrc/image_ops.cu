// Image pre/post steps of the CLI on the device (scope row 8f-3): short-side / arbitrary bilinear resize with an optional
// crop window, and the pixel-sized parts of --keep-colors (CORAL).
//
//   reference: utils.resize_to / center_crop / center_crop_to (utils.py:29-67) -> scipy.misc.imresize(interp='bilinear')
//              = Pillow's ImagingResample on the uint8 image; utils.preserve_colors_np (utils.py:87-90) ->
//              coral.coral_numpy (coral.py:13-39).
//
// Resize.  Bit-exact with Pillow's 8-bit path (src/libImaging/Resample.c): a separable triangle filter whose support is
// max(1, in/out) input pixels, coefficients computed in double precision, normalised by their sum, converted to 22-bit fixed
// point; the HORIZONTAL pass runs first and rounds to uint8, the vertical pass follows.  The double-precision coefficient
// arithmetic uses the explicit round-to-nearest intrinsics so that ptxas cannot contract it into FMAs (x86 Pillow builds do
// not).  This is byte/integer work bound by HBM: one thread per output pixel, the (<= ksize) taps of a pixel are contiguous
// in the input row (horizontal) or a strided column walk that neighbouring threads coalesce (vertical).
//
// CORAL.  The pixel-sized work is (i) nine integer sums per image (sum x_c, sum x_c x_d -- exact in uint64) and (ii) the
// per-pixel affine map out = (A ((x/255 - m_s) / s_s)) * s_t + m_t in double precision followed by the reference's
// clip / *255 / truncation.  The 3x3 algebra in between (matSqrt through numpy's SVD, whose sign conventions the
// reference's non-symmetric U sqrt(D) U depends on) stays on the host in wct_tf_b200/device_image.py.
#include "common.cuh"

namespace wctb {

static constexpr int RS_PRECISION_BITS = 32 - 8 - 2;
static inline size_t align_up(size_t v, size_t a) { return (v + a - 1) / a * a; }

__device__ __forceinline__ double tri_filter(double x) {
    if (x < 0.0) x = -x;
    return x < 1.0 ? __dsub_rn(1.0, x) : 0.0;
}

// precompute_coeffs + normalize_coeffs_8bpc (Resample.c) for the box [0, in_size): bounds[xx] = (xmin, count),
// kk[xx][ksize] fixed-point weights.
__global__ void k_resample_coeffs(int in_size, int out_size, int ksize, int* __restrict__ bounds, int* __restrict__ kk) {
    const int xx = blockIdx.x * blockDim.x + threadIdx.x;
    if (xx >= out_size) return;
    const double scale = __ddiv_rn((double)in_size, (double)out_size);
    const double filterscale = scale < 1.0 ? 1.0 : scale;
    const double support = filterscale;                     // BILINEAR.support = 1.0
    const double ss = __ddiv_rn(1.0, filterscale);
    const double center = __dmul_rn(__dadd_rn((double)xx, 0.5), scale);
    int xmin = (int)__dadd_rn(__dsub_rn(center, support), 0.5);
    if (xmin < 0) xmin = 0;
    int xmax = (int)__dadd_rn(__dadd_rn(center, support), 0.5);
    if (xmax > in_size) xmax = in_size;
    xmax -= xmin;
    if (xmax > ksize) xmax = ksize;                         // cannot happen (ksize = 2 ceil(support) + 1); keeps the table in bounds
    double ww = 0.0;
    for (int x = 0; x < xmax; ++x) {
        const double w = tri_filter(__dmul_rn(__dadd_rn(__dsub_rn((double)(x + xmin), center), 0.5), ss));
        ww = __dadd_rn(ww, w);
    }
    int* k = kk + (long long)xx * ksize;
    for (int x = 0; x < xmax; ++x) {
        double w = tri_filter(__dmul_rn(__dadd_rn(__dsub_rn((double)(x + xmin), center), 0.5), ss));
        if (ww != 0.0) w = __ddiv_rn(w, ww);
        const double kf = __dmul_rn(w, (double)(1 << RS_PRECISION_BITS));
        k[x] = (int)(kf < 0.0 ? __dadd_rn(-0.5, kf) : __dadd_rn(0.5, kf));
    }
    for (int x = xmax; x < ksize; ++x) k[x] = 0;
    bounds[2 * xx] = xmin;
    bounds[2 * xx + 1] = xmax;
}

__device__ __forceinline__ uint8_t clip8(int v) {
    v >>= RS_PRECISION_BITS;
    return (uint8_t)(v < 0 ? 0 : (v > 255 ? 255 : v));
}

// One pass along `axis_stride`: out[n][o][i] = sum_t in[n][(xmin_o + t)][i] * k_o[t] for the output window
// o in [o0, o0 + n_out); the other axis is walked as-is.  in: [N][in_len][row] (horizontal: a row is C bytes and the lines
// are the image rows; vertical: a "row" is Wout*C bytes).
//   horizontal: thread = (line y, output column xx, channel c):  in index  (y*Win + x)*C + c
//   vertical  : thread = (output row yy, byte j of the row)    :  in index  y*rowbytes + j
template <bool HORIZONTAL>
__global__ void k_resample_pass(const uint8_t* __restrict__ in, uint8_t* __restrict__ out, const int* __restrict__ bounds,
                                const int* __restrict__ kk, int ksize, int N, int lines, int in_len, int o0, int n_out, int C) {
    // horizontal: lines = image rows, in_len = input width, output [N][lines][n_out][C]
    // vertical  : lines = bytes per row (W*C),  in_len = input height, output [N][n_out][lines]
    const long long total = (long long)N * lines * n_out * (HORIZONTAL ? C : 1);
    for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x) {
        int o, c = 0;
        long long base;                                        // byte offset of tap 0 in `in`
        int step;
        if (HORIZONTAL) {
            c = (int)(i % C);
            const long long r = i / C;
            o = (int)(r % n_out);
            const long long ln = r / n_out;                    // n*lines + y
            const int xmin = bounds[2 * (o0 + o)];
            base = (ln * in_len + xmin) * C + c;
            step = C;
        } else {
            const int j = (int)(i % lines);
            const long long r = i / lines;
            o = (int)(r % n_out);
            const long long n = r / n_out;
            const int ymin = bounds[2 * (o0 + o)];
            base = (n * in_len + ymin) * lines + j;
            step = lines;
        }
        const int cnt = bounds[2 * (o0 + o) + 1];
        const int* k = kk + (long long)(o0 + o) * ksize;
        int acc = 1 << (RS_PRECISION_BITS - 1);
        for (int t = 0; t < cnt; ++t) acc += (int)in[base + (long long)t * step] * __ldg(k + t);
        out[i] = clip8(acc);
    }
}

__global__ void k_crop_u8(const uint8_t* __restrict__ in, uint8_t* __restrict__ out, int N, int H, int W, int C, int y0, int x0,
                          int Ho, int Wo) {
    const long long total = (long long)N * Ho * Wo * C;
    const int rowb = Wo * C;
    for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x) {
        const int j = (int)(i % rowb);
        const long long r = i / rowb;
        const int y = (int)(r % Ho);
        const long long n = r / Ho;
        out[i] = in[((n * H + y0 + y) * W + x0) * C + j];
    }
}

static int grid_for(long long total, int block) {
    long long g = (total + block - 1) / block;
    const long long cap = (long long)device_sm_count() * 16;
    return (int)(g < 1 ? 1 : (g > cap ? cap : g));
}

static int resample_ksize(int in_size, int out_size) {
    double scale = (double)in_size / (double)out_size;
    if (scale < 1.0) scale = 1.0;
    return (int)ceil(scale) * 2 + 1;
}

struct ResizeWs {
    size_t bx, kx, by, ky, tmp, total;
    int ksx, ksy;
};
static ResizeWs resize_layout(int N, int Hs, int Ws, int C, int Hd, int Wd, int Wout) {
    ResizeWs L;
    L.ksx = resample_ksize(Ws, Wd);
    L.ksy = resample_ksize(Hs, Hd);
    size_t o = 0;
    L.bx = o; o = align_up(o + (size_t)Wd * 2 * 4, 256);
    L.kx = o; o = align_up(o + (size_t)Wd * L.ksx * 4, 256);
    L.by = o; o = align_up(o + (size_t)Hd * 2 * 4, 256);
    L.ky = o; o = align_up(o + (size_t)Hd * L.ksy * 4, 256);
    L.tmp = o; o = align_up(o + (size_t)N * Hs * Wout * C, 256);
    L.total = o;
    return L;
}
size_t resize_workspace_bytes(int N, int Hs, int Ws, int C, int Hd, int Wd, int Wout) {
    return resize_layout(N, Hs, Ws, C, Hd, Wd, Wout).total;
}

// src [N][Hs][Ws][C] -> the window [y0, y0+Hout) x [x0, x0+Wout) of the Hd x Wd resample, dst [N][Hout][Wout][C]
int launch_resize_u8(const uint8_t* src, int N, int Hs, int Ws, int C, int Hd, int Wd, int y0, int x0, int Hout, int Wout,
                     uint8_t* dst, void* ws, size_t ws_bytes, cudaStream_t st) {
    const ResizeWs L = resize_layout(N, Hs, Ws, C, Hd, Wd, Wout);
    if (ws_bytes < L.total) {
        set_error("resize: workspace %zu < %zu bytes", ws_bytes, L.total);
        return WCTB200_EWS;
    }
    uint8_t* w = static_cast<uint8_t*>(ws);
    int* bx = reinterpret_cast<int*>(w + L.bx);
    int* kx = reinterpret_cast<int*>(w + L.kx);
    int* by = reinterpret_cast<int*>(w + L.by);
    int* ky = reinterpret_cast<int*>(w + L.ky);
    uint8_t* tmp = w + L.tmp;
    const bool need_h = Wd != Ws, need_v = Hd != Hs;      // Resample.c ImagingResampleInner: a pass runs only when its size changes
    if (!need_h && !need_v) {
        k_crop_u8<<<grid_for((long long)N * Hout * Wout * C, 256), 256, 0, st>>>(src, dst, N, Hs, Ws, C, y0, x0, Hout, Wout);
        WCTB_CHECK_LAUNCH("k_crop_u8");
        return 0;
    }
    const uint8_t* vin = src;      // input of the vertical pass: [N][Hs][vW][C]
    int vW = Ws;
    if (need_h) {
        k_resample_coeffs<<<cdiv(Wd, 128), 128, 0, st>>>(Ws, Wd, L.ksx, bx, kx);
        WCTB_CHECK_LAUNCH("k_resample_coeffs(x)");
        uint8_t* hout = need_v ? tmp : dst;
        if (!need_v && (y0 != 0 || Hout != Hs)) hout = tmp;
        k_resample_pass<true><<<grid_for((long long)N * Hs * Wout * C, 256), 256, 0, st>>>(src, hout, bx, kx, L.ksx, N, Hs, Ws, x0,
                                                                                           Wout, C);
        WCTB_CHECK_LAUNCH("k_resample_pass(h)");
        if (!need_v) {
            if (hout == tmp) {
                k_crop_u8<<<grid_for((long long)N * Hout * Wout * C, 256), 256, 0, st>>>(tmp, dst, N, Hs, Wout, C, y0, 0, Hout, Wout);
                WCTB_CHECK_LAUNCH("k_crop_u8");
            }
            return 0;
        }
        vin = tmp;
        vW = Wout;
    } else if (x0 != 0 || Wout != Ws) {
        // vertical pass only, with a column window: crop the columns first
        k_crop_u8<<<grid_for((long long)N * Hs * Wout * C, 256), 256, 0, st>>>(src, tmp, N, Hs, Ws, C, 0, x0, Hs, Wout);
        WCTB_CHECK_LAUNCH("k_crop_u8");
        vin = tmp;
        vW = Wout;
    }
    k_resample_coeffs<<<cdiv(Hd, 128), 128, 0, st>>>(Hs, Hd, L.ksy, by, ky);
    WCTB_CHECK_LAUNCH("k_resample_coeffs(y)");
    k_resample_pass<false><<<grid_for((long long)N * Hout * vW * C, 256), 256, 0, st>>>(vin, dst, by, ky, L.ksy, N, vW * C, Hs, y0,
                                                                                        Hout, 1);
    WCTB_CHECK_LAUNCH("k_resample_pass(v)");
    return 0;
}

// ---------------------------------------------------------------------------
// CORAL
// ---------------------------------------------------------------------------
// sums[0..2] = sum x_c, sums[3..8] = sum x_0x_0, x_0x_1, x_0x_2, x_1x_1, x_1x_2, x_2x_2   (exact integers)
__global__ void k_rgb_moments(const uint8_t* __restrict__ img, long long npix, unsigned long long* __restrict__ sums) {
    unsigned long long s[9];
#pragma unroll
    for (int i = 0; i < 9; ++i) s[i] = 0ull;
    for (long long p = (long long)blockIdx.x * blockDim.x + threadIdx.x; p < npix; p += (long long)gridDim.x * blockDim.x) {
        const unsigned int a = img[3 * p], b = img[3 * p + 1], c = img[3 * p + 2];
        s[0] += a; s[1] += b; s[2] += c;
        s[3] += a * a; s[4] += a * b; s[5] += a * c;
        s[6] += b * b; s[7] += b * c; s[8] += c * c;
    }
#pragma unroll
    for (int i = 0; i < 9; ++i) {
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) s[i] += __shfl_xor_sync(0xffffffffu, s[i], o);
    }
    if ((threadIdx.x & 31) == 0) {
#pragma unroll
        for (int i = 0; i < 9; ++i) atomicAdd(sums + i, s[i]);     // integer: order-independent, deterministic
    }
}

struct CoralParams {
    double A[9], sm[3], ss[3], ts[3], tm[3];
};

// out = uint8(clip((A ((x/255 - sm)/ss)) * ts + tm, 0, 1) * 255)    coral.py:24,35-36 + utils.py:88-89, in double like NumPy
__global__ void k_coral_apply(const uint8_t* __restrict__ src, long long npix, const CoralParams P, uint8_t* __restrict__ dst) {
    for (long long p = (long long)blockIdx.x * blockDim.x + threadIdx.x; p < npix; p += (long long)gridDim.x * blockDim.x) {
        double sn[3];
#pragma unroll
        for (int j = 0; j < 3; ++j)
            sn[j] = __ddiv_rn(__dsub_rn(__ddiv_rn((double)src[3 * p + j], 255.0), P.sm[j]), P.ss[j]);
#pragma unroll
        for (int i = 0; i < 3; ++i) {
            double v = __dmul_rn(P.A[3 * i], sn[0]);
            v = fma(P.A[3 * i + 1], sn[1], v);
            v = fma(P.A[3 * i + 2], sn[2], v);
            v = __dadd_rn(__dmul_rn(v, P.ts[i]), P.tm[i]);
            v = v < 0.0 ? 0.0 : (v > 1.0 ? 1.0 : v);
            dst[3 * p + i] = (uint8_t)(int)__dmul_rn(v, 255.0);
        }
    }
}

int launch_rgb_moments(const uint8_t* img, long long npix, unsigned long long* sums, cudaStream_t st) {
    WCTB_CUDA(cudaMemsetAsync(sums, 0, 9 * sizeof(unsigned long long), st));
    k_rgb_moments<<<grid_for(npix, 256), 256, 0, st>>>(img, npix, sums);
    WCTB_CHECK_LAUNCH("k_rgb_moments");
    return 0;
}

int launch_coral_apply(const uint8_t* src, long long npix, const double* A, const double* sm, const double* ss, const double* tm,
                       const double* ts, uint8_t* dst, cudaStream_t st) {
    CoralParams P;
    for (int i = 0; i < 9; ++i) P.A[i] = A[i];
    for (int i = 0; i < 3; ++i) {
        P.sm[i] = sm[i];
        P.ss[i] = ss[i];
        P.tm[i] = tm[i];
        P.ts[i] = ts[i];
    }
    k_coral_apply<<<grid_for(npix, 256), 256, 0, st>>>(src, npix, P, dst);
    WCTB_CHECK_LAUNCH("k_coral_apply");
    return 0;
}

}  // namespace wctb
