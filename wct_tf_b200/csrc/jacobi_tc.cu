// C = 512 symmetric eigensolver on the tensor cores (ops.py:53-55,110,123 call sites: the SVD of the feature covariance).
//
// One-sided (Hestenes) Jacobi with the block tournament of k_jacobi<512> (wct.cu): a cluster of 8 CTAs per matrix, 16 blocks of
// 32 columns paired round-robin, blocks exchanged through L2.  What is different is how a CTA orthogonalises its 64 columns
// G_p (512 x 64) in a round:
//   1. the Gram matrix S = G_p^T G_p (64 x 64) is computed ONCE on tcgen05 (split fp16, columns scaled by powers of two; the
//      row-major tile [512 rows][64 columns] is the MN-major operand, A = B = [hi | lo], like the C = 64 covariance kernel);
//   2. all rotations of the round (32 steps of 32 disjoint pairs; 63 in round 0, which also covers the pairs inside the two
//      blocks) are applied to S two-sidedly (S <- J^T S J: one 2 x 2 block per thread, conflict free, in place) and accumulated
//      into a 64 x 64 orthogonal V (V <- V J) in shared memory -- the 512-long columns are NOT touched;  one warp computes the
//      32 rotation parameters of the next step from S while the other 15 warps apply the previous step to V;
//   3. the columns are updated once per round as a GEMM on tcgen05:  G_p <- G_p V  (the SAME shared-memory tile is now the
//      K-major A operand; V, rescaled to the new column norms, is the B operand; split fp16 x 3 products, fp32 accumulate).
// The FFMA kernel spends ~120 warp-instructions per 512-element rotation (40 of them the packed FMAs of the update); here the
// update is tensor-core work and a rotation costs the S / V bookkeeping only (profiles/r02_jacobi_s_experiment.txt).
//
// Accuracy.  S decides the rotation ANGLES only and is recomputed from the columns every round.  V is a product of exact
// Givens rotations in fp32 (orthogonal to ~1e-7); the GEMM adds ~3e-7 relative noise per round to each column (relative to
// that column's own norm: both operands are scaled per column), i.e. an orthogonality floor of ~1e-6 per sweep, below the
// convergence threshold 2.7e-6, and a multiplicative drift of ~4e-6 over a whole run, which perturbs f(A) = E f(L) E^T by the
// same relative amount (eigenvalues are Rayleigh quotients against the pristine matrix anyway).  Deterministic: no atomics.
#include <cooperative_groups.h>

#include "common.cuh"
#include "jacobi_common.cuh"

namespace cg = cooperative_groups;

namespace wctb {

struct JtCfg {
    static constexpr int NN = 512;
    static constexpr int P = 8;                     // CTAs per cluster
    static constexpr int NB = 16, M = 15;           // 32-column blocks, rounds per sweep
    static constexpr int PLANE = 512 * 128;         // one fp16 plane of the tile: 512 rows x 64 columns
    static constexpr int TILE_BYTES = 2 * PLANE;    // hi plane | lo plane
    static constexpr int SLD = 65;                  // row stride of S and V (floats)
    static constexpr int SV_BYTES = 64 * SLD * 4;   // 16640
    static constexpr int SV_REGION = (2 * SV_BYTES + 1023) / 1024 * 1024;   // S and V; the B operand behind them must stay 1024-aligned
    static constexpr int B_BYTES = 2 * 64 * 128;    // V as split-fp16 B operand: hi 8 KB | lo 8 KB
    static constexpr int AUX_BYTES = 3072 + 4096;
    static constexpr int SMEM_BYTES = TILE_BYTES + SV_REGION + B_BYTES + AUX_BYTES + 1024;
};

__device__ float g_jacobi_tc_tolq = 1e-4f;

// MN-major SWIZZLE_128B descriptor with the two 64-wide MN groups (hi plane, lo plane) one PLANE apart
__device__ __forceinline__ uint64_t jt_desc_mn(uint32_t smem_addr) {
    uint64_t d = 0;
    d |= (uint64_t)((smem_addr >> 4) & 0x3FFF);
    d |= (uint64_t)((JtCfg::PLANE >> 4) & 0x3FFF) << 16;      // LBO
    d |= (uint64_t)((1024 >> 4) & 0x3FFF) << 32;              // SBO: 8-row K groups
    d |= (uint64_t)1 << 46;
    d |= (uint64_t)2 << 61;
    return d;
}

struct JtShared {
    float* S;          // [64][SLD] Gram matrix of the block pair (true scale)
    float* V;          // [64][SLD] accumulated rotation of the round
    float4* rec;       // [2][32] (c, s, bits(i | j << 8), -) of column pair p of the current / previous step
    unsigned short* sched;   // [63][32] i | j << 8: the pairs of every step of a round (built once per kernel)
    int* slotcol;      // [32][2] scratch of the schedule builder (warp 0 only)
};

// ---- pair schedule (warp 0 only).  32 slots of two columns; at level H the slots [base, base+H) meet the slots
// [base+H, base+2H): aligned halves (sub 0), crossed halves (sub 1: the halves of every upper slot are swapped first), then
// the upper slots rotate by one.  H = 16: the 32 x 32 pairs between the two blocks; H = 8,4,2,1 and the in-slot step (H = 0):
// the pairs inside the blocks.  The table is private to warp 0: __syncwarp is enough.
__device__ __forceinline__ void jt_swap_upper(int* slotcol, int H, int lane) {
    if (lane < 16) {
        const int grp = lane / H, k = lane - grp * H;
        const int v = grp * 2 * H + H + k;
        const int c0 = slotcol[2 * v], c1 = slotcol[2 * v + 1];
        slotcol[2 * v] = c1;
        slotcol[2 * v + 1] = c0;
    }
    __syncwarp();
}
__device__ __forceinline__ void jt_rotate_upper(int* slotcol, int H, int lane) {
    int c0 = 0, c1 = 0, v = 0;
    if (lane < 16) {
        const int grp = lane / H, k = lane - grp * H;
        v = grp * 2 * H + H + k;
        const int vn = grp * 2 * H + H + (k + 1 == H ? 0 : k + 1);
        c0 = slotcol[2 * vn];
        c1 = slotcol[2 * vn + 1];
    }
    __syncwarp();
    if (lane < 16) { slotcol[2 * v] = c0; slotcol[2 * v + 1] = c1; }
    __syncwarp();
}
__device__ __forceinline__ void jt_pair(const int* slotcol, int H, int lane, int& i, int& j) {
    if (H == 0) {
        i = slotcol[2 * lane];
        j = slotcol[2 * lane + 1];
    } else {
        const int q = lane >> 1, e = lane & 1;
        const int grp = q / H, k = q - grp * H;
        const int u = grp * 2 * H + k, v = u + H;
        i = slotcol[2 * u + e];
        j = slotcol[2 * v + e];
    }
}

// ---- V <- V J with the rows of V in REGISTERS (warps 1-2: thread = one row, 32 slots of two columns = 32 packed pairs).
// Slot pair q = (u, v) of level H carries the column pairs p = 2q (lower halves) and p = 2q+1 (upper halves) of the schedule;
// the swap / rotate transitions are the ones the schedule builder applied to its table, so register indices stay static.
__device__ __forceinline__ f32x2 jt_swap2(f32x2 v) {
    float a, b;
    unpack2(v, a, b);
    return pack2(b, a);
}
template <int H>
__device__ __forceinline__ void jt_v_apply(f32x2 (&slot)[32], const float4* __restrict__ rec) {
    const f32x2 zero = 0ull;
    if constexpr (H == 0) {
#pragma unroll
        for (int k = 0; k < 32; ++k) {
            const float4 r = rec[k];
            float x, y;
            unpack2(slot[k], x, y);
            slot[k] = pack2(r.x * x - r.y * y, r.y * x + r.x * y);
        }
    } else {
#pragma unroll
        for (int q = 0; q < 16; ++q) {
            const int grp = q / H, k = q - grp * H;
            const int u = grp * 2 * H + k, v = u + H;
            const float4 ra = rec[2 * q], rb = rec[2 * q + 1];
            const f32x2 c2 = pack2(ra.x, rb.x), s2 = pack2(ra.y, rb.y), ns2 = pack2(-ra.y, -rb.y);
            const f32x2 x = slot[u], y = slot[v];
            slot[u] = fma2(ns2, y, fma2(c2, x, zero));      // x' = c x - s y
            slot[v] = fma2(s2, x, fma2(c2, y, zero));       // y' = s x + c y
        }
    }
}
template <int H>
__device__ __forceinline__ void jt_v_swap(f32x2 (&slot)[32]) {
    if constexpr (H > 0) {
#pragma unroll
        for (int q = 0; q < 16; ++q) {
            const int grp = q / H, k = q - grp * H;
            const int v = grp * 2 * H + H + k;
            slot[v] = jt_swap2(slot[v]);
        }
    }
}
template <int H>
__device__ __forceinline__ void jt_v_rotate(f32x2 (&slot)[32]) {
    if constexpr (H > 1) {
#pragma unroll
        for (int grp = 0; grp < 16 / H; ++grp) {
            const int b0 = grp * 2 * H + H;
            const f32x2 first = slot[b0];
#pragma unroll
            for (int k = 0; k < H - 1; ++k) slot[b0 + k] = slot[b0 + k + 1];
            slot[b0 + H - 1] = first;
        }
    }
}
// apply step `n` of the schedule to the register rows: first the transition the builder made before that step, then the rotations
__device__ __forceinline__ void jt_v_step(f32x2 (&slot)[32], const float4* __restrict__ rec, int n, int& H, int& sg, int& sub) {
    if (n > 0) {
        if (H > 0 && sub == 0) {
            switch (H) { case 16: jt_v_swap<16>(slot); break; case 8: jt_v_swap<8>(slot); break; case 4: jt_v_swap<4>(slot); break;
                         case 2: jt_v_swap<2>(slot); break; default: jt_v_swap<1>(slot); break; }
            sub = 1;
        } else {
            switch (H) { case 16: jt_v_rotate<16>(slot); break; case 8: jt_v_rotate<8>(slot); break; case 4: jt_v_rotate<4>(slot); break;
                         case 2: jt_v_rotate<2>(slot); break; default: break; }
            sub = 0;
            if (++sg >= H) { sg = 0; H = H > 1 ? H / 2 : 0; }
        }
    }
    switch (H) { case 16: jt_v_apply<16>(slot, rec); break; case 8: jt_v_apply<8>(slot, rec); break; case 4: jt_v_apply<4>(slot, rec); break;
                 case 2: jt_v_apply<2>(slot, rec); break; case 1: jt_v_apply<1>(slot, rec); break; default: jt_v_apply<0>(slot, rec); break; }
}

__global__ void __launch_bounds__(512, 1)
k_jacobi_tc(float* __restrict__ Gall, float* __restrict__ nrm_all, float* __restrict__ conv_ws, int* __restrict__ sweeps_out,
            int max_sweeps, float tol, unsigned int* err) {
    using Cfg = JtCfg;
    constexpr int NN = Cfg::NN, P = Cfg::P, NB = Cfg::NB, M = Cfg::M, SLD = Cfg::SLD;
    extern __shared__ uint8_t smem_raw[];
    uint8_t* smem = smem_raw + ((1024u - (smem_u32(smem_raw) & 1023u)) & 1023u);
    uint8_t* tile = smem;                                           // [hi plane | lo plane], row r at r*128, SWIZZLE_128B
    JtShared sh;
    sh.S = reinterpret_cast<float*>(smem + Cfg::TILE_BYTES);
    sh.V = reinterpret_cast<float*>(smem + Cfg::TILE_BYTES + Cfg::SV_BYTES);
    uint8_t* Bop = smem + Cfg::TILE_BYTES + Cfg::SV_REGION;         // 1024-aligned (SWIZZLE_128B operands)
    uint8_t* aux = Bop + Cfg::B_BYTES;
    sh.rec = reinterpret_cast<float4*>(aux);                        // 1024 B
    sh.slotcol = reinterpret_cast<int*>(aux + 1024);                // 256 B
    sh.sched = reinterpret_cast<unsigned short*>(aux + 3072);       // 4032 B
    unsigned char* fin = reinterpret_cast<unsigned char*>(aux + 2560);   // [2][64] slot half -> column at the end of a round
    float* scl = reinterpret_cast<float*>(aux + 1280);              // [64] power-of-two scale of the Gram / apply operand
    float* iscl = reinterpret_cast<float*>(aux + 1536);             // [64] its inverse
    float* scl2 = reinterpret_cast<float*>(aux + 1792);             // [64] scale of the NEW columns (apply output)
    float* iscl2 = reinterpret_cast<float*>(aux + 2048);            // [64]
    uint64_t* mma_bar = reinterpret_cast<uint64_t*>(aux + 2304);
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(aux + 2312);
    volatile int* abort_flag = reinterpret_cast<volatile int*>(aux + 2316);
    unsigned int* s_flag = reinterpret_cast<unsigned int*>(aux + 2320);
    unsigned int* s_amax = reinterpret_cast<unsigned int*>(aux + 2324);

    const int t = threadIdx.x, warp = t >> 5, lane = t & 31;
    const int rank = blockIdx.x, prob = blockIdx.y;
    float* G = Gall + (long long)prob * NN * NN;
    float* nrm = nrm_all + (long long)prob * NN;
    float* cw = conv_ws + (long long)prob * 16;
    cg::cluster_group cluster = cg::this_cluster();

    if (t == 0) {
        mbar_init(mma_bar, 1);
        *abort_flag = 0;
        fence_barrier_init();
    }
    if (warp == 1) tmem_alloc(tmem_slot, 512);
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem_base = *tmem_slot;

    // ---- the pair schedule of a round, once: 63 steps x 32 disjoint pairs (the first 32 steps are the cross pairs) ----
    if (warp == 0) {
        sh.slotcol[lane] = lane;
        sh.slotcol[32 + lane] = 32 + lane;                              // slot k holds columns 2k, 2k+1
        __syncwarp();
        int H = 16, sg = 0, sub = 0;
        for (int n = 0; n < 63; ++n) {
            if (n > 0) {
                if (H > 0 && sub == 0) { jt_swap_upper(sh.slotcol, H, lane); sub = 1; }
                else {
                    if (H > 1) jt_rotate_upper(sh.slotcol, H, lane);
                    sub = 0;
                    if (++sg >= H) { sg = 0; H = H > 1 ? H / 2 : 0; }       // 16 -> 8 -> 4 -> 2 -> 1 -> 0 (in-slot pairs)
                }
            }
            int i, j;
            jt_pair(sh.slotcol, H, lane, i, j);
            sh.sched[n * 32 + lane] = (unsigned short)(i | (j << 8));
            // the column held by every slot half when a round ends after this step (cross-only rounds: 32 steps; round 0: 63)
            if (n == 31) { fin[lane] = (unsigned char)sh.slotcol[lane]; fin[32 + lane] = (unsigned char)sh.slotcol[32 + lane]; }
            if (n == 62) { fin[64 + lane] = (unsigned char)sh.slotcol[lane]; fin[96 + lane] = (unsigned char)sh.slotcol[32 + lane]; }
            __syncwarp();
        }
    }
    __syncthreads();

    const float tol2 = tol * tol;
    const float tolq2 = g_jacobi_tc_tolq * g_jacobi_tc_tolq;
    float null2 = 0.f;
    uint32_t mma_phase = 0;
    int sweep = 0;
    for (; sweep < max_sweeps; ++sweep) {
        if (t == 0) { *s_flag = 0u; *s_amax = 0u; }
        float flag = 0.f, amax = 0.f;
        for (int r = 0; r < M; ++r) {
            int bt, bb;
            if (rank == 0) { bt = NB - 1; bb = r; }
            else { bt = (r + rank) % M; bb = (r - rank + M) % M; }
            // ---- scales of the 64 columns (from the carried norms), V = I, schedule table ----
            if (t < 64) {
                const int col = (t < 32 ? bt * 32 + t : bb * 32 + t - 32);
                const float n2 = __ldcg(nrm + col);
                float sc = 1.f;
                if (n2 > 0.f) {
                    int e2;
                    frexpf(n2, &e2);                                   // n2 = m 2^e2: |column| ~ 2^(e2/2)
                    int ex = 8 - ((e2 + (e2 >= 0 ? 1 : 0)) / 2);
                    ex = ex < -60 ? -60 : (ex > 60 ? 60 : ex);
                    sc = exp2f((float)ex);
                }
                scl[t] = sc;
                iscl[t] = 1.f / sc;
            }
            __syncthreads();
            // ---- load row t of the 64 columns (coalesced: a warp reads 128 contiguous bytes per column), scale, split,
            //      store as the shared-memory tile: row t at t*128 in both planes, 16-byte chunk c at c ^ (t & 7) ----
            {
                uint8_t* hi_row = tile + t * 128;
                uint8_t* lo_row = hi_row + Cfg::PLANE;
#pragma unroll
                for (int c = 0; c < 8; ++c) {
                    const int blk = c < 4 ? bt : bb;
                    const float* src = G + (long long)(blk * 32 + (c & 3) * 8) * NN + t;
                    float v[8];
#pragma unroll
                    for (int j = 0; j < 8; ++j) v[j] = __ldcg(src + (long long)j * NN) * scl[8 * c + j];
                    Half8 h, l;
                    split8(v, h, l);
                    const int off = (c ^ (t & 7)) << 4;
                    *reinterpret_cast<Half8*>(hi_row + off) = h;
                    *reinterpret_cast<Half8*>(lo_row + off) = l;
                }
            }
            fence_proxy_async();
            __syncthreads();
            // ---- Gram: D[128][128] = [hi | lo]^T [hi | lo] over the 512 rows (32 k-steps of 16 rows) ----
            if (t == 0) {
                tc_fence_after();
                constexpr uint32_t idesc = umma_idesc_f16_mn(128, 128);
                const uint32_t base = smem_u32(tile);
#pragma unroll 1
                for (int ks = 0; ks < 32; ++ks) {
                    const uint64_t d = jt_desc_mn(base + ks * 2048);
                    umma_f16(tmem_base, d, d, idesc, ks == 0 ? 0u : 1u);
                }
                umma_commit(mma_bar);
            }
            mbar_wait(mma_bar, mma_phase, abort_flag, err, 0x600u);
            mma_phase ^= 1u;
            tc_fence_after();
            // ---- S[i][j] = (D[i][j] + D[i][64+j] + D[64+i][j] + D[64+i][64+j]) / (scl_i scl_j) ----
            if (warp < 4) {
                const int m = warp * 32 + lane;                        // accumulator row (TMEM lane)
                const uint32_t tsrc = tmem_base + ((uint32_t)(warp * 32) << 16);
                float part[64];
#pragma unroll
                for (int cc = 0; cc < 4; ++cc) {
                    uint32_t r0[16], r1[16];
                    tmem_ld16(tsrc + cc * 16, r0);
                    tmem_ld16(tsrc + 64 + cc * 16, r1);
                    tmem_ld_wait();
#pragma unroll
                    for (int j = 0; j < 16; ++j) part[cc * 16 + j] = __uint_as_float(r0[j]) + __uint_as_float(r1[j]);
                }
                tc_fence_before();
                if (warp < 2) {
#pragma unroll
                    for (int j = 0; j < 64; ++j) sh.S[m * SLD + j] = part[j];
                }
                asm volatile("bar.sync 1, 128;" ::: "memory");       // warps 0-3: rows 0..63 are stored
                if (warp >= 2) {
#pragma unroll
                    for (int j = 0; j < 64; ++j) sh.S[(m - 64) * SLD + j] += part[j];
                }
            }
            __syncthreads();
            for (int e = t; e < 4096; e += 512) {
                const int i = e >> 6, j = e & 63;
                sh.S[i * SLD + j] *= iscl[i] * iscl[j];
            }
            __syncthreads();
            if (t < 64) amax = fmaxf(amax, sh.S[t * SLD + t]);

            // ---- the rotations of the round: S <- J^T S J, V <- V J ----
            // sub-step n: phase A  warp 0: schedule + parameters of step n  ||  warps 1-15: V update of step n-1
            //             phase B  all warps: S update of step n
            const int nsub = (r == 0) ? 63 : 32;
            f32x2 vrow[32];                                             // warps 1-2: row (t - 32) of V, slot k = columns 2k, 2k+1
            int vH = 16, vsg = 0, vsub = 0;
            if (warp == 1 || warp == 2) {
                const int row = t - 32;
#pragma unroll
                for (int k = 0; k < 32; ++k) vrow[k] = pack2(row == 2 * k ? 1.f : 0.f, row == 2 * k + 1 ? 1.f : 0.f);
            }
            for (int n = 0; n <= nsub; ++n) {
                const int cur = n & 1, prv = cur ^ 1;
                if (warp == 0) {
                    if (n < nsub) {
                        const unsigned int ij = sh.sched[n * 32 + lane];
                        const int i = ij & 255, j = ij >> 8;
                        const float a = sh.S[i * SLD + i], b = sh.S[j * SLD + j], g = sh.S[i * SLD + j];
                        float tt, s, cm1;
                        rot_scalars(g, a, b, tol2, tolq2, null2, flag, tt, s, cm1);
                        sh.rec[cur * 32 + lane] = make_float4(1.f + cm1, s, __uint_as_float(ij), 0.f);
                    }
                } else if ((warp == 1 || warp == 2) && n > 0) {
                    jt_v_step(vrow, sh.rec + prv * 32, n - 1, vH, vsg, vsub);       // V <- V J for step n-1, thread-local
                }
                __syncthreads();
                if (n < nsub) {
                    // S <- J^T S J : thread (warp w, lane l) owns the 2 x 2 blocks (pair 2w, pair l) and (pair 2w+1, pair l)
                    const float4 rq = sh.rec[cur * 32 + lane];
                    const unsigned int ijq = __float_as_uint(rq.z);
                    const int iq = ijq & 255, jq = ijq >> 8;
#pragma unroll
                    for (int h = 0; h < 2; ++h) {
                        const float4 rp = sh.rec[cur * 32 + 2 * warp + h];
                        const unsigned int ijp = __float_as_uint(rp.z);
                        float* r0 = sh.S + (ijp & 255) * SLD;
                        float* r1 = sh.S + (ijp >> 8) * SLD;
                        const float a = r0[iq], b = r0[jq], c = r1[iq], d = r1[jq];
                        const float a1 = a * rq.x - b * rq.y, b1 = a * rq.y + b * rq.x;     // columns
                        const float c1 = c * rq.x - d * rq.y, d1 = c * rq.y + d * rq.x;
                        r0[iq] = a1 * rp.x - c1 * rp.y;                                      // rows
                        r0[jq] = b1 * rp.x - d1 * rp.y;
                        r1[iq] = a1 * rp.y + c1 * rp.x;
                        r1[jq] = b1 * rp.y + d1 * rp.x;
                    }
                    __syncthreads();
                }
            }
            if (warp == 1 || warp == 2) {
                const unsigned char* fc = fin + (r == 0 ? 64 : 0);
                float* vr = sh.V + (t - 32) * SLD;
#pragma unroll
                for (int k = 0; k < 32; ++k) {
                    float a, b;
                    unpack2(vrow[k], a, b);
                    vr[fc[2 * k]] = a;
                    vr[fc[2 * k + 1]] = b;
                }
            }
            // ---- scales of the NEW columns (their norms are the diagonal of the updated S), carried norms ----
            if (t < 64) {
                const float n2 = fmaxf(sh.S[t * SLD + t], 0.f);
                float sc = 1.f;
                if (n2 > 0.f) {
                    int e2;
                    frexpf(n2, &e2);
                    int ex = 8 - ((e2 + (e2 >= 0 ? 1 : 0)) / 2);
                    ex = ex < -60 ? -60 : (ex > 60 ? 60 : ex);
                    sc = exp2f((float)ex);
                }
                scl2[t] = sc;
                iscl2[t] = 1.f / sc;
                nrm[t < 32 ? bt * 32 + t : bb * 32 + t - 32] = n2;
            }
            __syncthreads();
            // ---- B operand: Bhat[j][k] = V[k][j] * scl2[j] / scl[k]  (K-major: row = new column j, 64 old columns k) ----
            {
                const int j = t >> 3, c = t & 7;
                float v[8];
#pragma unroll
                for (int q = 0; q < 8; ++q) {
                    const int k = 8 * c + q;
                    v[q] = sh.V[k * SLD + j] * (scl2[j] * iscl[k]);
                }
                Half8 h, l;
                split8(v, h, l);
                const int off = j * 128 + ((c ^ (j & 7)) << 4);
                *reinterpret_cast<Half8*>(Bop + off) = h;
                *reinterpret_cast<Half8*>(Bop + 8192 + off) = l;
            }
            fence_proxy_async();
            tc_fence_before();                                          // the TMEM reads of the Gram are ordered before the apply MMAs
            __syncthreads();
            // ---- apply: D_mt[128 rows][128] = A_hi [B_hi | B_lo]^T  (+ A_lo B_hi^T into the first 64 columns), mt = 0..3 ----
            if (t == 0) {
                tc_fence_after();
                constexpr uint32_t idesc128 = umma_idesc_f16(128, 128), idesc64 = umma_idesc_f16(128, 64);
                const uint32_t a_base = smem_u32(tile), b_base = smem_u32(Bop);
#pragma unroll 1
                for (int mt = 0; mt < 4; ++mt) {
                    const uint64_t a_hi = umma_desc_sw128(a_base + mt * 16384);
                    const uint64_t a_lo = umma_desc_sw128(a_base + Cfg::PLANE + mt * 16384);
                    const uint64_t b_hi = umma_desc_sw128(b_base);
                    const uint32_t tacc = tmem_base + (uint32_t)(mt * 128);
#pragma unroll
                    for (int ks = 0; ks < 4; ++ks) {
                        const uint64_t ko = (uint64_t)(ks * 32 >> 4);
                        umma_f16(tacc, a_hi + ko, b_hi + ko, idesc128, ks == 0 ? 0u : 1u);
                        umma_f16(tacc, a_lo + ko, b_hi + ko, idesc64, 1u);
                    }
                }
                umma_commit(mma_bar);
            }
            mbar_wait(mma_bar, mma_phase, abort_flag, err, 0x601u);
            mma_phase ^= 1u;
            tc_fence_after();
            // ---- read out row t of the new columns, undo the scale, store (a warp writes 128 contiguous bytes per column) ----
            {
                const uint32_t tsrc = tmem_base + ((uint32_t)((warp & 3) * 32) << 16) + (uint32_t)((warp >> 2) * 128);
#pragma unroll
                for (int cc = 0; cc < 4; ++cc) {
                    uint32_t r0[16], r1[16];
                    tmem_ld16(tsrc + cc * 16, r0);
                    tmem_ld16(tsrc + 64 + cc * 16, r1);
                    tmem_ld_wait();
#pragma unroll
                    for (int q = 0; q < 16; ++q) {
                        const int j = cc * 16 + q;
                        const float val = (__uint_as_float(r0[q]) + __uint_as_float(r1[q])) * iscl2[j];
                        G[(long long)(j < 32 ? bt * 32 + j : bb * 32 + j - 32) * NN + t] = val;
                    }
                }
                tc_fence_before();
            }
            __threadfence();
            cluster.sync();   // release/acquire: next round reads what the peers just wrote
        }
        // ---- convergence: worst pair class seen in this sweep (0 / 1 / 2), agreed across the cluster ----
        if (warp == 0) {
#pragma unroll
            for (int o = 16; o >= 1; o >>= 1) flag = fmaxf(flag, __shfl_xor_sync(0xffffffffu, flag, o));
            if (lane == 0) *s_flag = __float_as_uint(flag);
        }
        if (t < 64) atomicMax(s_amax, __float_as_uint(amax));          // max of non-negative floats: order independent
        __syncthreads();
        float gmax = __uint_as_float(*s_flag);
        float amx = __uint_as_float(*s_amax);
        if (t == 0) {
            reinterpret_cast<volatile float*>(cw)[rank] = gmax;
            reinterpret_cast<volatile float*>(cw)[8 + rank] = amx;
            __threadfence();
        }
        cluster.sync();
        gmax = 0.f;
        amx = 0.f;
        for (int i = 0; i < P; ++i) {
            gmax = fmaxf(gmax, reinterpret_cast<volatile float*>(cw)[i]);
            amx = fmaxf(amx, reinterpret_cast<volatile float*>(cw)[8 + i]);
        }
        cluster.sync();   // everyone has read before the next sweep overwrites
        null2 = 1e-11f * amx;
        __syncthreads();
        if (gmax < 2.f) { ++sweep; break; }
    }
    if (rank == 0 && t == 0 && sweeps_out) sweeps_out[prob] = sweep;
    tc_fence_before();
    __syncthreads();
    if (warp == 1) tmem_dealloc(tmem_base, 512);
}

// squared column norms of `count` n x n matrices (columns contiguous): one warp per column, fixed-order reduction
__global__ void k_col_norms(const float* __restrict__ G, int n, int total_cols, float* __restrict__ nrm) {
    const int col = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
    const int lane = threadIdx.x & 31;
    if (col >= total_cols) return;
    const float* c = G + (long long)col * n;
    float s = 0.f;
    for (int i = lane; i < n; i += 32) s = fmaf(c[i], c[i], s);
#pragma unroll
    for (int o = 16; o >= 1; o >>= 1) s += __shfl_xor_sync(0xffffffffu, s, o);
    if (lane == 0) nrm[col] = s;
}

int g_jacobi_impl = 2;      // 1 = k_jacobi<512> (wct.cu), 2 = k_jacobi_tc (C = 512 only)

// G: [count][512][512] symmetric (overwritten); conv_ws: [count][16] floats; sweeps: [count] or null
int launch_jacobi_tc(float* G, int count, float* conv_ws, int* sweeps, cudaStream_t st) {
    using Cfg = JtCfg;
    float* nrm = nullptr;
    { int rc0 = scratch_alloc(reinterpret_cast<void**>(&nrm), (size_t)count * Cfg::NN * sizeof(float), st, 3); if (rc0) return rc0; }
    k_col_norms<<<cdiv((long long)count * Cfg::NN, 8), 256, 0, st>>>(G, Cfg::NN, count * Cfg::NN, nrm);
    WCTB_CHECK_LAUNCH("k_col_norms");
    const float tol = 2.f * sqrtf((float)Cfg::NN) * 5.96e-8f;
    WCTB_ENSURE_SMEM(k_jacobi_tc, Cfg::SMEM_BYTES);
    cudaLaunchConfig_t cfg = {};
    cfg.gridDim = dim3((unsigned)Cfg::P, (unsigned)count, 1);
    cfg.blockDim = dim3(512, 1, 1);
    cfg.dynamicSmemBytes = Cfg::SMEM_BYTES;
    cfg.stream = st;
    cudaLaunchAttribute at[1];
    at[0].id = cudaLaunchAttributeClusterDimension;
    at[0].val.clusterDim.x = (unsigned)Cfg::P;
    at[0].val.clusterDim.y = 1;
    at[0].val.clusterDim.z = 1;
    cfg.attrs = at;
    cfg.numAttrs = 1;
    WCTB_CUDA(cudaLaunchKernelEx(&cfg, k_jacobi_tc, G, nrm, conv_ws, sweeps, 40, tol, device_error_word()));
    return 0;
}

}  // namespace wctb
