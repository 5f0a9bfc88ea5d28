// Fused attention for the spatial self / cross attention of Transformer2DModel (head_dim 64), forward and backward:
// replaces  bgemm(QK^T, fp32 scores) -> softmax -> bgemm(PV)  and its 4-GEMM + softmax backward, which are store-bound on
// the fp32 score tensor (335 MB per 1024-token layer of a 16-frame clip).  Reference call site:
// diffusers AttnProcessor2_0 -> F.scaled_dot_product_attention, enabled by train.py:138-152.
//
// STATUS (end of round 1): verified on B200 through this C ABI against a PyTorch fp32 attention - values, log-sum-exp and
// all three gradients, fused-pitch operands, ragged lengths, split-query cross-attention; worst rel-to-max error 4.7e-3,
// 96 us forward / 328 us backward for 16 frames x 1024 tokens x 5 heads vs 268 / 431 us unfused
// (profiles/r1_flash_attn_experiment.txt), and with T2V_FLASH_ATTN=1 the end-to-end UNet parity tests pass on B200
// (profiles/r1_flash_attn_suite.txt).  Not yet benchmarked inside the step: the model uses it only when T2V_FLASH_ATTN=1
// (ops._use_flash).
//
// FlashAttention-2 blocking on warp-level tensor-core MMAs (mma.sync.m16n8k16 + ldmatrix, the helpers of attn_small.cu),
// cp.async double-buffered K/V tiles, online softmax in the accumulator registers.
//   forward : CTA = 64 query rows x one (batch, head); 4 warps x 16 rows; key blocks of 64; writes O and LSE
//   backward: delta = rowsum(dO o O) -> dQ kernel (CTA per query block, loops over key blocks)
//                                      dK/dV kernel (CTA per key block [x query split], loops over query blocks;
//                                      split mode reduces with red.global.add.f32 into fp32 scratch for cross-attention,
//                                      where Lq = F*H*W >> Lk = 77)
// q / k / v (and dq / dk / dv) are addressed as [batch][row][head*64 + d] with arbitrary row pitch and batch stride, so
// they can be column slices of the fused QKV / K|V projections.
#include "common.h"
#include "mma_sync.cuh"
#include "ptx.cuh"

#include <algorithm>
#include <cmath>
#include <cuda_bf16.h>

namespace t2v {
namespace fa {
using namespace wmma16;   // ldmatrix / mma.sync fragment helpers (mma_sync.cuh)

constexpr int D = 64;                    // head dim
constexpr int BM = 64, BN = 64;          // query rows / key rows per block
constexpr int kThreads = 128;            // 4 warps x 16 query rows
constexpr int PITCH = (D + 8) * 2;       // bytes per tile row: 144 -> conflict-free ldmatrix
constexpr int TILE = BM * PITCH;         // 9216 bytes
constexpr float kLog2e = 1.4426950408889634f;

struct Mat {                             // [batch][row][head * 64 + d]
    const __nv_bfloat16* p;
    int64_t ld, bs;                      // row pitch, batch stride (elements)
};
struct MatW {
    __nv_bfloat16* p;
    int64_t ld, bs;
};

// acc += A(registers: a 16 x K tile in accumulator layout, packed to bf16) * B(K x N in shared memory, stored [k][n])
template <int N, int K>
__device__ __forceinline__ void warp_mma_regA(float (&acc)[N / 8][4], const float (&a_acc)[K / 8][4], uint32_t sb, int pb, int lane) {
#pragma unroll
    for (int kt = 0; kt < K / 16; ++kt) {
        uint32_t a[4];
        a[0] = pack_bf16(a_acc[2 * kt][0], a_acc[2 * kt][1]);
        a[1] = pack_bf16(a_acc[2 * kt][2], a_acc[2 * kt][3]);
        a[2] = pack_bf16(a_acc[2 * kt + 1][0], a_acc[2 * kt + 1][1]);
        a[3] = pack_bf16(a_acc[2 * kt + 1][2], a_acc[2 * kt + 1][3]);
#pragma unroll
        for (int np = 0; np < N / 16; ++np) {
            uint32_t b[4];
            load_b2<true>(sb, pb, np * 16, kt * 16, lane, b);
            mma_bf16(acc[2 * np], a, b[0], b[1]);
            mma_bf16(acc[2 * np + 1], a, b[2], b[3]);
        }
    }
}

__device__ __forceinline__ void cp_async16(uint32_t saddr, const void* g, bool valid) {
    const int sz = valid ? 16 : 0;   // src-size 0: the 16 destination bytes are zero-filled
    asm volatile("cp.async.cg.shared.global [%0], [%1], 16, %2;" ::"r"(saddr), "l"(g), "r"(sz) : "memory");
}
__device__ __forceinline__ void cp_async_commit() { asm volatile("cp.async.commit_group;" ::: "memory"); }
template <int N>
__device__ __forceinline__ void cp_async_wait() { asm volatile("cp.async.wait_group %0;" ::"n"(N) : "memory"); }

// 64 rows x 64 columns (one head) of a [batch][row][...] matrix -> shared tile; rows >= nrows are zero-filled
__device__ __forceinline__ void load_tile_async(uint8_t* sm, const __nv_bfloat16* base, int64_t ld, int row0, int nrows) {
    const uint32_t s = smem_u32(sm);
#pragma unroll
    for (int i = 0; i < (BM * 8) / kThreads; ++i) {
        const int idx = threadIdx.x + i * kThreads;
        const int r = idx >> 3, c = idx & 7;
        const bool ok = row0 + r < nrows;
        const __nv_bfloat16* g = ok ? base + int64_t(row0 + r) * ld + c * 8 : base;
        cp_async16(s + r * PITCH + c * 16, g, ok);
    }
}
// accumulator rows (this warp's 16 rows at m0) -> bf16 tile in shared memory
__device__ __forceinline__ void stage_acc(const float (&acc)[D / 8][4], uint8_t* sm, int m0, int lane) {
    const int g = lane >> 2, t = lane & 3;
#pragma unroll
    for (int nt = 0; nt < D / 8; ++nt) {
        *reinterpret_cast<uint32_t*>(sm + (m0 + g) * PITCH + (nt * 8 + 2 * t) * 2) = pack_bf16(acc[nt][0], acc[nt][1]);
        *reinterpret_cast<uint32_t*>(sm + (m0 + g + 8) * PITCH + (nt * 8 + 2 * t) * 2) = pack_bf16(acc[nt][2], acc[nt][3]);
    }
}
// this warp's 16 staged rows -> global (coalesced 16-byte stores), rows >= nrows skipped
__device__ __forceinline__ void store_rows(const uint8_t* sm, __nv_bfloat16* base, int64_t ld, int row0, int m0, int nrows, int lane) {
#pragma unroll
    for (int i = 0; i < 4; ++i) {
        const int idx = lane + 32 * i;
        const int r = m0 + (idx >> 3), c = idx & 7;
        if (row0 + r < nrows)
            *(reinterpret_cast<uint4*>(base + int64_t(row0 + r) * ld) + c) = *reinterpret_cast<const uint4*>(sm + r * PITCH + c * 16);
    }
}

// ------------------------------------------------------------------------------------------------ forward
// grid (ceil(Lq / 64), heads, Nb); O [Nb][Lq][heads*64] (pitch o.ld), lse [Nb][heads][Lq] (natural log)
__global__ void __launch_bounds__(kThreads) flash_fwd_kernel(Mat q, Mat k, Mat v, MatW o, float* __restrict__ lse, int Lq, int Lk,
                                                             int heads, float scale) {
    pdl_sync();
    extern __shared__ __align__(16) uint8_t sm[];
    uint8_t* sQ = sm;
    uint8_t* sK = sm + TILE;        // [2]
    uint8_t* sV = sm + 3 * TILE;    // [2]
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int g = lane >> 2, t = lane & 3;
    const int q0 = blockIdx.x * BM, h = blockIdx.y, b = blockIdx.z;
    const __nv_bfloat16* qb = q.p + b * q.bs + h * D;
    const __nv_bfloat16* kb = k.p + b * k.bs + h * D;
    const __nv_bfloat16* vb = v.p + b * v.bs + h * D;
    const int nkb = (Lk + BN - 1) / BN;
    load_tile_async(sQ, qb, q.ld, q0, Lq);
    load_tile_async(sK, kb, k.ld, 0, Lk);
    load_tile_async(sV, vb, v.ld, 0, Lk);
    cp_async_commit();
    const uint32_t aQ = smem_u32(sQ);
    const int m0 = warp * 16;
    const float sc2 = scale * kLog2e;
    float mrow[2] = {-INFINITY, -INFINITY}, lrow[2] = {0.f, 0.f};
    float oacc[D / 8][4];
#pragma unroll
    for (int nt = 0; nt < D / 8; ++nt) oacc[nt][0] = oacc[nt][1] = oacc[nt][2] = oacc[nt][3] = 0.f;
    for (int j = 0; j < nkb; ++j) {
        const int st = j & 1;
        if (j + 1 < nkb) {   // prefetch the next K/V block into the other stage (free since the barrier ending iteration j-1)
            load_tile_async(sK + (st ^ 1) * TILE, kb, k.ld, (j + 1) * BN, Lk);
            load_tile_async(sV + (st ^ 1) * TILE, vb, v.ld, (j + 1) * BN, Lk);
            cp_async_commit();
            cp_async_wait<1>();
        } else {
            cp_async_wait<0>();
        }
        __syncthreads();
        float s[BN / 8][4];
#pragma unroll
        for (int nt = 0; nt < BN / 8; ++nt) s[nt][0] = s[nt][1] = s[nt][2] = s[nt][3] = 0.f;
        warp_mma<BN, D, false, false>(s, aQ, PITCH, m0, smem_u32(sK + st * TILE), PITCH, lane);
        // online softmax (base-2 domain)
        float mx[2] = {mrow[0], mrow[1]};
#pragma unroll
        for (int nt = 0; nt < BN / 8; ++nt)
#pragma unroll
            for (int e = 0; e < 4; ++e) {
                const int col = j * BN + nt * 8 + 2 * t + (e & 1);
                s[nt][e] = col < Lk ? s[nt][e] * sc2 : -INFINITY;
                mx[e >> 1] = fmaxf(mx[e >> 1], s[nt][e]);
            }
        float alpha[2];
#pragma unroll
        for (int hh = 0; hh < 2; ++hh) {
            mx[hh] = fmaxf(mx[hh], __shfl_xor_sync(0xffffffffu, mx[hh], 1));
            mx[hh] = fmaxf(mx[hh], __shfl_xor_sync(0xffffffffu, mx[hh], 2));
            alpha[hh] = exp2f(mrow[hh] - mx[hh]);   // first block: exp2(-inf) = 0
            mrow[hh] = mx[hh];
            lrow[hh] *= alpha[hh];
        }
#pragma unroll
        for (int nt = 0; nt < BN / 8; ++nt)
#pragma unroll
            for (int e = 0; e < 4; ++e) {
                s[nt][e] = exp2f(s[nt][e] - mx[e >> 1]);
                lrow[e >> 1] += s[nt][e];   // per-thread partial row sum; reduced over the quad once, at the end
            }
#pragma unroll
        for (int nt = 0; nt < D / 8; ++nt) {
            oacc[nt][0] *= alpha[0]; oacc[nt][1] *= alpha[0];
            oacc[nt][2] *= alpha[1]; oacc[nt][3] *= alpha[1];
        }
        warp_mma_regA<D, BN>(oacc, s, smem_u32(sV + st * TILE), PITCH, lane);   // O += P V
        __syncthreads();   // everyone is done with stage st before iteration j+1 prefetches into it
    }
#pragma unroll
    for (int hh = 0; hh < 2; ++hh) {
        lrow[hh] += __shfl_xor_sync(0xffffffffu, lrow[hh], 1);
        lrow[hh] += __shfl_xor_sync(0xffffffffu, lrow[hh], 2);
    }
    const float inv[2] = {1.0f / lrow[0], 1.0f / lrow[1]};
#pragma unroll
    for (int nt = 0; nt < D / 8; ++nt) {
        oacc[nt][0] *= inv[0]; oacc[nt][1] *= inv[0];
        oacc[nt][2] *= inv[1]; oacc[nt][3] *= inv[1];
    }
    // each warp reads only its own 16 rows of sQ, so it may overwrite them with its O rows without a block barrier
    __syncwarp();
    stage_acc(oacc, sQ, m0, lane);
    __syncwarp();
    store_rows(sQ, o.p + b * o.bs + h * D, o.ld, q0, m0, Lq, lane);
    if (t == 0) {
        float* l = lse + (int64_t(b) * heads + h) * Lq;
        if (q0 + m0 + g < Lq) l[q0 + m0 + g] = (mrow[0] + log2f(lrow[0])) / kLog2e;
        if (q0 + m0 + g + 8 < Lq) l[q0 + m0 + g + 8] = (mrow[1] + log2f(lrow[1])) / kLog2e;
    }
}

// ------------------------------------------------------------------------------------------------ backward
// delta[b][h][i] = sum_d dO[i][d] * O[i][d]; one warp per (b, h, row): lane owns 2 of the 64 head dims
__global__ void flash_delta_kernel(Mat o, Mat dout, float* __restrict__ delta, int Nb, int heads, int Lq) {
    pdl_sync();
    const int lane = threadIdx.x & 31;
    const int64_t warp = (blockIdx.x * int64_t(blockDim.x) + threadIdx.x) >> 5;
    const int64_t nwarps = (int64_t(gridDim.x) * blockDim.x) >> 5;
    const int64_t total = int64_t(Nb) * heads * Lq;
    for (int64_t w = warp; w < total; w += nwarps) {
        const int i = int(w % Lq);
        const int h = int((w / Lq) % heads);
        const int b = int(w / (int64_t(Lq) * heads));
        const uint32_t a = __ldg(reinterpret_cast<const uint32_t*>(o.p + b * o.bs + int64_t(i) * o.ld + h * D) + lane);
        const uint32_t c = __ldg(reinterpret_cast<const uint32_t*>(dout.p + b * dout.bs + int64_t(i) * dout.ld + h * D) + lane);
        float d = bf16_lo(a) * bf16_lo(c) + bf16_hi(a) * bf16_hi(c);
#pragma unroll
        for (int off = 16; off > 0; off >>= 1) d += __shfl_xor_sync(0xffffffffu, d, off);
        if (lane == 0) delta[w] = d;
    }
}

// P and dS of one 16 x 64 block in accumulator layout, from S (raw QK^T) and dP (dO V^T):
//   P = exp2(S * scale * log2e - lse2[row]),  dS = P * (dP - delta[row]) * scale;  columns >= Lk give P = dS = 0
__device__ __forceinline__ void p_and_ds(float (&s)[BN / 8][4], float (&dp)[BN / 8][4], const float (&lse2)[2], const float (&dl)[2],
                                         int col0, int Lk, float scale, int t) {
    const float sc2 = scale * kLog2e;
#pragma unroll
    for (int nt = 0; nt < BN / 8; ++nt)
#pragma unroll
        for (int e = 0; e < 4; ++e) {
            const int col = col0 + nt * 8 + 2 * t + (e & 1);
            const float p = col < Lk ? exp2f(s[nt][e] * sc2 - lse2[e >> 1]) : 0.f;
            s[nt][e] = p;
            dp[nt][e] = p * (dp[nt][e] - dl[e >> 1]) * scale;
        }
}

// dQ: grid (ceil(Lq / 64), heads, Nb); loops over key blocks
__global__ void __launch_bounds__(kThreads) flash_bwd_dq_kernel(Mat q, Mat k, Mat v, Mat dout, const float* __restrict__ lse,
                                                                const float* __restrict__ delta, MatW dq, int Lq, int Lk, int heads,
                                                                float scale) {
    pdl_sync();
    extern __shared__ __align__(16) uint8_t sm[];
    uint8_t* sQ = sm;
    uint8_t* sD = sm + TILE;
    uint8_t* sK = sm + 2 * TILE;    // [2]
    uint8_t* sV = sm + 4 * TILE;    // [2]
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int g = lane >> 2, t = lane & 3;
    const int q0 = blockIdx.x * BM, h = blockIdx.y, b = blockIdx.z;
    const __nv_bfloat16* kb = k.p + b * k.bs + h * D;
    const __nv_bfloat16* vb = v.p + b * v.bs + h * D;
    const int nkb = (Lk + BN - 1) / BN;
    load_tile_async(sQ, q.p + b * q.bs + h * D, q.ld, q0, Lq);
    load_tile_async(sD, dout.p + b * dout.bs + h * D, dout.ld, q0, Lq);
    load_tile_async(sK, kb, k.ld, 0, Lk);
    load_tile_async(sV, vb, v.ld, 0, Lk);
    cp_async_commit();
    const int m0 = warp * 16;
    const uint32_t aQ = smem_u32(sQ), aD = smem_u32(sD);
    float lse2[2], dl[2];
    {
        const int64_t base = (int64_t(b) * heads + h) * Lq;
        const int r0 = q0 + m0 + g, r1 = r0 + 8;
        lse2[0] = r0 < Lq ? lse[base + r0] * kLog2e : INFINITY;   // +inf: P = 0 for rows past the end
        lse2[1] = r1 < Lq ? lse[base + r1] * kLog2e : INFINITY;
        dl[0] = r0 < Lq ? delta[base + r0] : 0.f;
        dl[1] = r1 < Lq ? delta[base + r1] : 0.f;
    }
    float acc[D / 8][4];
#pragma unroll
    for (int nt = 0; nt < D / 8; ++nt) acc[nt][0] = acc[nt][1] = acc[nt][2] = acc[nt][3] = 0.f;
    for (int j = 0; j < nkb; ++j) {
        const int st = j & 1;
        if (j + 1 < nkb) {
            load_tile_async(sK + (st ^ 1) * TILE, kb, k.ld, (j + 1) * BN, Lk);
            load_tile_async(sV + (st ^ 1) * TILE, vb, v.ld, (j + 1) * BN, Lk);
            cp_async_commit();
            cp_async_wait<1>();
        } else {
            cp_async_wait<0>();
        }
        __syncthreads();
        float s[BN / 8][4], dp[BN / 8][4];
#pragma unroll
        for (int nt = 0; nt < BN / 8; ++nt)
#pragma unroll
            for (int e = 0; e < 4; ++e) s[nt][e] = dp[nt][e] = 0.f;
        const uint32_t aK = smem_u32(sK + st * TILE), aV = smem_u32(sV + st * TILE);
        warp_mma<BN, D, false, false>(s, aQ, PITCH, m0, aK, PITCH, lane);    // S  = Q K^T
        warp_mma<BN, D, false, false>(dp, aD, PITCH, m0, aV, PITCH, lane);   // dP = dO V^T
        p_and_ds(s, dp, lse2, dl, j * BN, Lk, scale, t);
        warp_mma_regA<D, BN>(acc, dp, aK, PITCH, lane);                      // dQ += dS K
        __syncthreads();
    }
    __syncwarp();
    stage_acc(acc, sQ, m0, lane);
    __syncwarp();
    store_rows(sQ, dq.p + b * dq.bs + h * D, dq.ld, q0, m0, Lq, lane);
}

// dK, dV: grid (ceil(Lk / 64) * qsplits, heads, Nb); each CTA owns one key block and a range of query blocks.
// qsplits == 1: bf16 results written to dk / dv.  qsplits > 1: fp32 partials reduced into dk32 / dv32 ([Nb][Lk][heads*64])
__global__ void __launch_bounds__(kThreads) flash_bwd_dkv_kernel(Mat q, Mat k, Mat v, Mat dout, const float* __restrict__ lse,
                                                                 const float* __restrict__ delta, MatW dk, MatW dv,
                                                                 float* __restrict__ dk32, float* __restrict__ dv32, int Lq, int Lk,
                                                                 int heads, float scale, int qsplits, int qblocks_per_split) {
    pdl_sync();
    extern __shared__ __align__(16) uint8_t sm[];
    uint8_t* sK = sm;
    uint8_t* sV = sm + TILE;
    uint8_t* sQ = sm + 2 * TILE;    // [2]
    uint8_t* sD = sm + 4 * TILE;    // [2]
    uint8_t* sP = sm + 6 * TILE;    // [64 queries][64 keys] bf16
    uint8_t* sS = sm + 7 * TILE;    // dS, same shape
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int g = lane >> 2, t = lane & 3;
    const int kblk = blockIdx.x / qsplits, split = blockIdx.x % qsplits;
    const int k0 = kblk * BN, h = blockIdx.y, b = blockIdx.z;
    const __nv_bfloat16* qb = q.p + b * q.bs + h * D;
    const __nv_bfloat16* db = dout.p + b * dout.bs + h * D;
    const int nqb_all = (Lq + BM - 1) / BM;
    const int i_begin = split * qblocks_per_split, i_end = min(nqb_all, i_begin + qblocks_per_split);
    if (i_begin >= i_end) return;
    load_tile_async(sK, k.p + b * k.bs + h * D, k.ld, k0, Lk);
    load_tile_async(sV, v.p + b * v.bs + h * D, v.ld, k0, Lk);
    load_tile_async(sQ, qb, q.ld, i_begin * BM, Lq);
    load_tile_async(sD, db, dout.ld, i_begin * BM, Lq);
    cp_async_commit();
    const int m0 = warp * 16;
    const uint32_t aK = smem_u32(sK), aV = smem_u32(sV), aP = smem_u32(sP), aS = smem_u32(sS);
    const int64_t stat_base = (int64_t(b) * heads + h) * Lq;
    float dkacc[D / 8][4], dvacc[D / 8][4];
#pragma unroll
    for (int nt = 0; nt < D / 8; ++nt)
#pragma unroll
        for (int e = 0; e < 4; ++e) dkacc[nt][e] = dvacc[nt][e] = 0.f;
    for (int i = i_begin; i < i_end; ++i) {
        const int st = (i - i_begin) & 1;
        if (i + 1 < i_end) {
            load_tile_async(sQ + (st ^ 1) * TILE, qb, q.ld, (i + 1) * BM, Lq);
            load_tile_async(sD + (st ^ 1) * TILE, db, dout.ld, (i + 1) * BM, Lq);
            cp_async_commit();
            cp_async_wait<1>();
        } else {
            cp_async_wait<0>();
        }
        __syncthreads();
        const uint32_t aQ = smem_u32(sQ + st * TILE), aD = smem_u32(sD + st * TILE);
        float lse2[2], dl[2];
        {
            const int r0 = i * BM + m0 + g, r1 = r0 + 8;
            lse2[0] = r0 < Lq ? lse[stat_base + r0] * kLog2e : INFINITY;
            lse2[1] = r1 < Lq ? lse[stat_base + r1] * kLog2e : INFINITY;
            dl[0] = r0 < Lq ? delta[stat_base + r0] : 0.f;
            dl[1] = r1 < Lq ? delta[stat_base + r1] : 0.f;
        }
        float s[BN / 8][4], dp[BN / 8][4];
#pragma unroll
        for (int nt = 0; nt < BN / 8; ++nt)
#pragma unroll
            for (int e = 0; e < 4; ++e) s[nt][e] = dp[nt][e] = 0.f;
        warp_mma<BN, D, false, false>(s, aQ, PITCH, m0, aK, PITCH, lane);    // S  (this warp's 16 queries x 64 keys)
        warp_mma<BN, D, false, false>(dp, aD, PITCH, m0, aV, PITCH, lane);   // dP
        p_and_ds(s, dp, lse2, dl, k0, Lk, scale, t);
#pragma unroll
        for (int nt = 0; nt < BN / 8; ++nt) {   // park P and dS: rows = queries, columns = keys
            const int col = (nt * 8 + 2 * t) * 2;
            *reinterpret_cast<uint32_t*>(sP + (m0 + g) * PITCH + col) = pack_bf16(s[nt][0], s[nt][1]);
            *reinterpret_cast<uint32_t*>(sP + (m0 + g + 8) * PITCH + col) = pack_bf16(s[nt][2], s[nt][3]);
            *reinterpret_cast<uint32_t*>(sS + (m0 + g) * PITCH + col) = pack_bf16(dp[nt][0], dp[nt][1]);
            *reinterpret_cast<uint32_t*>(sS + (m0 + g + 8) * PITCH + col) = pack_bf16(dp[nt][2], dp[nt][3]);
        }
        __syncthreads();
        // this warp's 16 KEY rows: dV += P^T dO, dK += dS^T Q  (contraction over the 64 queries of the block)
        warp_mma<D, BM, true, true>(dvacc, aP, PITCH, m0, aD, PITCH, lane);
        warp_mma<D, BM, true, true>(dkacc, aS, PITCH, m0, aQ, PITCH, lane);
        __syncthreads();   // sP / sS / stage st are free again
    }
    if (qsplits == 1) {
        __syncwarp();
        stage_acc(dkacc, sP, m0, lane);
        stage_acc(dvacc, sS, m0, lane);
        __syncwarp();
        store_rows(sP, dk.p + b * dk.bs + h * D, dk.ld, k0, m0, Lk, lane);
        store_rows(sS, dv.p + b * dv.bs + h * D, dv.ld, k0, m0, Lk, lane);
    } else {
        const int64_t C = int64_t(heads) * D;
#pragma unroll
        for (int nt = 0; nt < D / 8; ++nt)
#pragma unroll
            for (int e = 0; e < 4; ++e) {
                const int r = k0 + m0 + g + (e >> 1) * 8, c = nt * 8 + 2 * t + (e & 1);
                if (r < Lk) {
                    const int64_t off = (int64_t(b) * Lk + r) * C + h * D + c;
                    atomicAdd(dk32 + off, dkacc[nt][e]);
                    atomicAdd(dv32 + off, dvacc[nt][e]);
                }
            }
    }
}

}  // namespace fa
}  // namespace t2v

using namespace t2v;
using namespace t2v::fa;

namespace {
template <typename Kernel>
void set_smem(Kernel kernel, int bytes) {
    cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, bytes);
}
int check(int D_, int64_t ld) {
    if (D_ != D) return fail(-2, "flash_attn: head_dim %d unsupported (64)", D_);
    if (ld % 8) return fail(-2, "flash_attn: row pitches must be multiples of 8 elements");
    return 0;
}
}  // namespace

extern "C" {

int t2v_flash_attn_fwd(const void* q, const void* k, const void* v, void* o, float* lse, int32_t Nb, int32_t heads, int32_t Lq,
                       int32_t Lk, int32_t head_dim, int64_t q_ld, int64_t q_bs, int64_t k_ld, int64_t k_bs, int64_t v_ld,
                       int64_t v_bs, int64_t o_ld, int64_t o_bs, void* stream) {
    if (int r = check(head_dim, q_ld | k_ld | v_ld | o_ld)) return r;
    if (Nb <= 0 || heads <= 0 || Lq <= 0 || Lk <= 0) return fail(-2, "flash_attn_fwd: bad shape");
    static bool once = false;
    if (!once) { set_smem(flash_fwd_kernel, 5 * TILE); once = true; }
    const Mat Q{static_cast<const __nv_bfloat16*>(q), q_ld, q_bs}, K{static_cast<const __nv_bfloat16*>(k), k_ld, k_bs},
        V{static_cast<const __nv_bfloat16*>(v), v_ld, v_bs};
    const MatW O{static_cast<__nv_bfloat16*>(o), o_ld, o_bs};
    const dim3 grid((Lq + BM - 1) / BM, heads, Nb);
    const int rc = int(launch_pdl(flash_fwd_kernel, grid, dim3(kThreads), size_t(5 * TILE), static_cast<cudaStream_t>(stream), Q, K, V, O,
                                  lse, Lq, Lk, heads, 1.0f / sqrtf(float(D))));
    return launch_checked(rc, "flash_attn_fwd");
}

int32_t t2v_flash_attn_bwd_splits(int32_t Nb, int32_t heads, int32_t Lq, int32_t Lk) {
    const int64_t ctas = int64_t((Lk + BN - 1) / BN) * heads * Nb;
    const int nqb = (Lq + BM - 1) / BM;
    if (ctas >= 148 || nqb < 8) return 1;
    return int(std::min<int64_t>((2 * 148 + ctas - 1) / ctas, nqb / 4));
}

int t2v_flash_attn_bwd(const void* q, const void* k, const void* v, const void* o, const void* dout, const float* lse, void* dq,
                       void* dk, void* dv, float* delta_ws, float* dkv_ws, int32_t Nb, int32_t heads, int32_t Lq, int32_t Lk,
                       int32_t head_dim, int64_t q_ld, int64_t q_bs, int64_t k_ld, int64_t k_bs, int64_t v_ld, int64_t v_bs,
                       int64_t o_ld, int64_t o_bs, int64_t dq_ld, int64_t dq_bs, int64_t dk_ld, int64_t dk_bs, int64_t dv_ld,
                       int64_t dv_bs, void* stream) {
    if (int r = check(head_dim, q_ld | k_ld | v_ld | o_ld | dq_ld | dk_ld | dv_ld)) return r;
    if (Nb <= 0 || heads <= 0 || Lq <= 0 || Lk <= 0) return fail(-2, "flash_attn_bwd: bad shape");
    static bool once = false;
    if (!once) {
        set_smem(flash_bwd_dq_kernel, 6 * TILE);
        set_smem(flash_bwd_dkv_kernel, 8 * TILE);
        once = true;
    }
    cudaStream_t st = static_cast<cudaStream_t>(stream);
    auto B = [](const void* p) { return static_cast<const __nv_bfloat16*>(p); };
    const Mat Q{B(q), q_ld, q_bs}, K{B(k), k_ld, k_bs}, V{B(v), v_ld, v_bs}, O{B(o), o_ld, o_bs}, DO{B(dout), o_ld, o_bs};
    const MatW DQ{static_cast<__nv_bfloat16*>(dq), dq_ld, dq_bs}, DK{static_cast<__nv_bfloat16*>(dk), dk_ld, dk_bs},
        DV{static_cast<__nv_bfloat16*>(dv), dv_ld, dv_bs};
    const float scale = 1.0f / sqrtf(float(D));
    const int64_t rows = int64_t(Nb) * heads * Lq;
    int rc = int(launch_pdl(flash_delta_kernel, dim3(int(std::min<int64_t>((rows + 7) / 8, 148 * 16))), dim3(256), size_t(0), st, O, DO,
                            delta_ws, Nb, heads, Lq));
    if (rc) return launch_checked(rc, "flash_attn_bwd(delta)");
    rc = int(launch_pdl(flash_bwd_dq_kernel, dim3((Lq + BM - 1) / BM, heads, Nb), dim3(kThreads), size_t(6 * TILE), st, Q, K, V, DO, lse,
                        static_cast<const float*>(delta_ws), DQ, Lq, Lk, heads, scale));
    if (rc) return launch_checked(rc, "flash_attn_bwd(dq)");
    const int splits = t2v_flash_attn_bwd_splits(Nb, heads, Lq, Lk);
    if (splits > 1 && !dkv_ws) return fail(-3, "flash_attn_bwd: this shape splits the query range and needs dkv_ws");
    const int nqb = (Lq + BM - 1) / BM;
    const int per = (nqb + splits - 1) / splits;
    float* dk32 = splits > 1 ? dkv_ws : nullptr;
    float* dv32 = splits > 1 ? dkv_ws + int64_t(Nb) * Lk * heads * D : nullptr;
    rc = int(launch_pdl(flash_bwd_dkv_kernel, dim3(((Lk + BN - 1) / BN) * splits, heads, Nb), dim3(kThreads), size_t(8 * TILE), st, Q, K,
                        V, DO, lse, static_cast<const float*>(delta_ws), DK, DV, dk32, dv32, Lq, Lk, heads, scale, splits, per));
    count_launch(2);
    return launch_checked(rc, "flash_attn_bwd(dkv)");
}

}  // extern "C"
