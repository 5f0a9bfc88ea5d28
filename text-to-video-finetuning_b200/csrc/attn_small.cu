// Short-sequence self-attention (L <= 32) for TransformerTemporalModel: attention along the frame axis.
//
// The reference permutes (B,C,F,H,W) -> (B*H*W, F, C) before this attention (diffusers TransformerTemporalModel,
// wired at unet_3d_blocks.py:331-340,491-500 and unet_3d_condition.py:147-152).  Here activations stay in the
// frames-major token order [B][F][H*W][C]; a sequence is addressed with strides instead (SeqAddr), so both permute
// copies disappear, and q / k / v may be column slices of one fused [rows][3C] projection.
//
// One warp owns one (sequence, head).  The tiles are 16 x 64 (L x D): far below one tcgen05 instruction (M = 128), so
// the four small products run on warp-level tensor-core MMAs (mma.sync m16n8k16, bf16 in / fp32 accumulate) fed by
// ldmatrix from shared memory; softmax and its gradient stay in the accumulator registers.  Global traffic is 16-byte
// coalesced in both directions (q, k, v, dO read once; o / dq, dk, dv written once): the kernel is HBM-bound.
#include "common.h"
#include "mma_sync.cuh"
#include "ptx.cuh"

#include <algorithm>
#include <cuda_bf16.h>

namespace t2v {
using namespace wmma16;   // ldmatrix / mma.sync fragment helpers (mma_sync.cuh)

constexpr int kMaxL = 32;

// Sequence addressing in ROWS of a token matrix; q/k/v (and their gradients) have row pitch ld_in - which lets them be
// column slices of one fused [rows][3C] QKV projection - while o / dO have row pitch ld_out.
struct SeqAddr {
    int64_t outer_rows, inner_rows, seq_rows;
    int64_t ld_in, ld_out;
    int32_t inner;
};

__device__ __forceinline__ int64_t seq_row(const SeqAddr& a, int64_t z) {
    return (z / a.inner) * a.outer_rows + (z % a.inner) * a.inner_rows;
}


// global [L rows x D] (row stride `stride` elements) -> shared [LP][D + 8] bf16; rows >= L stay zero (filled once at start)
template <int D>
__device__ __forceinline__ void load_tile(const __nv_bfloat16* __restrict__ g, int64_t base, int64_t stride, int L, uint8_t* sm, int lane) {
    constexpr int CPT = D / 8, PITCH = (D + 8) * 2;
    const int total = L * CPT;
#pragma unroll 4
    for (int idx = lane; idx < total; idx += 32) {
        const int t = idx / CPT, c = idx % CPT;
        *reinterpret_cast<uint4*>(sm + t * PITCH + c * 16) = __ldg(reinterpret_cast<const uint4*>(g + base + t * stride) + c);
    }
}
template <int D>
__device__ __forceinline__ void store_tile(__nv_bfloat16* __restrict__ g, int64_t base, int64_t stride, int L, const uint8_t* sm, int lane) {
    constexpr int CPT = D / 8, PITCH = (D + 8) * 2;
    const int total = L * CPT;
#pragma unroll 4
    for (int idx = lane; idx < total; idx += 32) {
        const int t = idx / CPT, c = idx % CPT;
        *(reinterpret_cast<uint4*>(g + base + t * stride) + c) = *reinterpret_cast<const uint4*>(sm + t * PITCH + c * 16);
    }
}
// accumulator tile (rows m0.. of a [LP][N] result) -> bf16 staging
template <int N>
__device__ __forceinline__ void stage_acc(const float (&acc)[N / 8][4], uint8_t* sm, int pitch, int m0, int lane) {
    const int g = lane >> 2, t = lane & 3;
#pragma unroll
    for (int nt = 0; nt < N / 8; ++nt) {
        *reinterpret_cast<uint32_t*>(sm + (m0 + g) * pitch + (nt * 8 + 2 * t) * 2) = pack_bf16(acc[nt][0], acc[nt][1]);
        *reinterpret_cast<uint32_t*>(sm + (m0 + g + 8) * pitch + (nt * 8 + 2 * t) * 2) = pack_bf16(acc[nt][2], acc[nt][3]);
    }
}

// Row softmax of a 16 x LP score tile held in accumulator layout (columns >= L masked); returns P in place.
template <int LP>
__device__ __forceinline__ void softmax_rows(float (&s)[LP / 8][4], int L, float scale, int lane) {
    const int t = lane & 3;
    float mx[2] = {-INFINITY, -INFINITY};
#pragma unroll
    for (int nt = 0; nt < LP / 8; ++nt)
#pragma unroll
        for (int e = 0; e < 4; ++e) {
            const int j = nt * 8 + 2 * t + (e & 1);
            s[nt][e] = j < L ? s[nt][e] * scale : -INFINITY;
            mx[e >> 1] = fmaxf(mx[e >> 1], s[nt][e]);
        }
#pragma unroll
    for (int h = 0; h < 2; ++h) {
        mx[h] = fmaxf(mx[h], __shfl_xor_sync(0xffffffffu, mx[h], 1));
        mx[h] = fmaxf(mx[h], __shfl_xor_sync(0xffffffffu, mx[h], 2));
    }
    float sum[2] = {0.f, 0.f};
#pragma unroll
    for (int nt = 0; nt < LP / 8; ++nt)
#pragma unroll
        for (int e = 0; e < 4; ++e) {
            s[nt][e] = __expf(s[nt][e] - mx[e >> 1]);
            sum[e >> 1] += s[nt][e];
        }
#pragma unroll
    for (int h = 0; h < 2; ++h) {
        sum[h] += __shfl_xor_sync(0xffffffffu, sum[h], 1);
        sum[h] += __shfl_xor_sync(0xffffffffu, sum[h], 2);
        sum[h] = 1.0f / sum[h];
    }
#pragma unroll
    for (int nt = 0; nt < LP / 8; ++nt)
#pragma unroll
        for (int e = 0; e < 4; ++e) s[nt][e] *= sum[e >> 1];
}

template <int D, int LP>
struct AttnSmem {
    static constexpr int kTilePitch = (D + 8) * 2;            // bytes per row of a [LP][D] tile (conflict-free ldmatrix)
    static constexpr int kTileBytes = LP * kTilePitch;
    static constexpr int kProbPitch = (LP + 8) * 2;
    static constexpr int kProbBytes = LP * kProbPitch;
    static constexpr int kFwdBytes = 4 * kTileBytes;                    // q, k, v, staging
    static constexpr int kBwdBytes = 5 * kTileBytes + 2 * kProbBytes;   // q, k, v, dO, staging, P, dS
};

template <int D, int LP>
__global__ void attn_small_fwd_kernel(const __nv_bfloat16* __restrict__ q, const __nv_bfloat16* __restrict__ k,
                                      const __nv_bfloat16* __restrict__ v, __nv_bfloat16* __restrict__ o, SeqAddr a, int64_t nseq,
                                      int heads, int L, float scale) {
    pdl_sync();
    using S = AttnSmem<D, LP>;
    extern __shared__ __align__(16) uint8_t sm_all[];
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    uint8_t* sq = sm_all + warp * S::kFwdBytes;
    uint8_t* sk = sq + S::kTileBytes;
    uint8_t* sv = sk + S::kTileBytes;
    uint8_t* so = sv + S::kTileBytes;
    for (int i = lane; i < S::kFwdBytes / 16; i += 32) reinterpret_cast<uint4*>(sq)[i] = make_uint4(0, 0, 0, 0);
    const uint32_t aq = smem_u32(sq), ak = smem_u32(sk), av = smem_u32(sv);
    const int nwarps = blockDim.x >> 5;
    const int64_t total = nseq * heads;
    for (int64_t w = blockIdx.x * int64_t(nwarps) + warp; w < total; w += int64_t(gridDim.x) * nwarps) {
        const int64_t z = w / heads;
        const int h = int(w % heads);
        const int64_t row0 = seq_row(a, z);
        const int64_t base = row0 * a.ld_in + int64_t(h) * D, sstr = a.seq_rows * a.ld_in;
        const int64_t obase = row0 * a.ld_out + int64_t(h) * D, ostr = a.seq_rows * a.ld_out;
        __syncwarp();
        load_tile<D>(q, base, sstr, L, sq, lane);
        load_tile<D>(k, base, sstr, L, sk, lane);
        load_tile<D>(v, base, sstr, L, sv, lane);
        __syncwarp();
#pragma unroll
        for (int mt = 0; mt < LP / 16; ++mt) {
            if (mt * 16 >= L) break;
            float s[LP / 8][4];
#pragma unroll
            for (int nt = 0; nt < LP / 8; ++nt) s[nt][0] = s[nt][1] = s[nt][2] = s[nt][3] = 0.f;
            warp_mma<LP, D, false, false>(s, aq, S::kTilePitch, mt * 16, ak, S::kTilePitch, lane);  // S = Q K^T
            softmax_rows<LP>(s, L, scale, lane);
            float oacc[D / 8][4];
#pragma unroll
            for (int nt = 0; nt < D / 8; ++nt) oacc[nt][0] = oacc[nt][1] = oacc[nt][2] = oacc[nt][3] = 0.f;
#pragma unroll
            for (int kt = 0; kt < LP / 16; ++kt) {  // O = P V: the score accumulators ARE the A fragments of P
                uint32_t pa[4];
                pa[0] = pack_bf16(s[2 * kt][0], s[2 * kt][1]);
                pa[1] = pack_bf16(s[2 * kt][2], s[2 * kt][3]);
                pa[2] = pack_bf16(s[2 * kt + 1][0], s[2 * kt + 1][1]);
                pa[3] = pack_bf16(s[2 * kt + 1][2], s[2 * kt + 1][3]);
#pragma unroll
                for (int np = 0; np < D / 16; ++np) {
                    uint32_t b[4];
                    load_b2<true>(av, S::kTilePitch, np * 16, kt * 16, lane, b);
                    mma_bf16(oacc[2 * np], pa, b[0], b[1]);
                    mma_bf16(oacc[2 * np + 1], pa, b[2], b[3]);
                }
            }
            stage_acc<D>(oacc, so, S::kTilePitch, mt * 16, lane);
        }
        __syncwarp();
        store_tile<D>(o, obase, ostr, L, so, lane);
    }
}

// Backward: per 16-row block recompute P, form dP = dO V^T and dS = P o (dP - rowsum(P o dP)) * scale in registers, park
// P and dS (bf16) in shared memory, then dQ = dS K, dK = dS^T Q, dV = P^T dO (transposed operands via ldmatrix.trans).
template <int D, int LP>
__global__ void attn_small_bwd_kernel(const __nv_bfloat16* __restrict__ q, const __nv_bfloat16* __restrict__ k,
                                      const __nv_bfloat16* __restrict__ v, const __nv_bfloat16* __restrict__ dout,
                                      __nv_bfloat16* __restrict__ dq, __nv_bfloat16* __restrict__ dk, __nv_bfloat16* __restrict__ dv,
                                      SeqAddr a, int64_t nseq, int heads, int L, float scale) {
    pdl_sync();
    using S = AttnSmem<D, LP>;
    extern __shared__ __align__(16) uint8_t sm_all[];
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    uint8_t* sq = sm_all + warp * S::kBwdBytes;
    uint8_t* sk = sq + S::kTileBytes;
    uint8_t* sv = sk + S::kTileBytes;
    uint8_t* sd = sv + S::kTileBytes;
    uint8_t* so = sd + S::kTileBytes;
    uint8_t* sp = so + S::kTileBytes;
    uint8_t* ss = sp + S::kProbBytes;
    for (int i = lane; i < S::kBwdBytes / 16; i += 32) reinterpret_cast<uint4*>(sq)[i] = make_uint4(0, 0, 0, 0);
    const uint32_t aq = smem_u32(sq), ak = smem_u32(sk), av = smem_u32(sv), ad = smem_u32(sd), ap = smem_u32(sp), as = smem_u32(ss);
    const int nwarps = blockDim.x >> 5;
    const int64_t total = nseq * heads;
    const int g = lane >> 2, t = lane & 3;
    for (int64_t w = blockIdx.x * int64_t(nwarps) + warp; w < total; w += int64_t(gridDim.x) * nwarps) {
        const int64_t z = w / heads;
        const int h = int(w % heads);
        const int64_t row0 = seq_row(a, z);
        const int64_t base = row0 * a.ld_in + int64_t(h) * D, sstr = a.seq_rows * a.ld_in;
        const int64_t obase = row0 * a.ld_out + int64_t(h) * D, ostr = a.seq_rows * a.ld_out;
        __syncwarp();
        load_tile<D>(q, base, sstr, L, sq, lane);
        load_tile<D>(k, base, sstr, L, sk, lane);
        load_tile<D>(v, base, sstr, L, sv, lane);
        load_tile<D>(dout, obase, ostr, L, sd, lane);
        __syncwarp();
#pragma unroll
        for (int mt = 0; mt < LP / 16; ++mt) {
            float s[LP / 8][4], dp[LP / 8][4];
#pragma unroll
            for (int nt = 0; nt < LP / 8; ++nt)
#pragma unroll
                for (int e = 0; e < 4; ++e) s[nt][e] = dp[nt][e] = 0.f;
            warp_mma<LP, D, false, false>(s, aq, S::kTilePitch, mt * 16, ak, S::kTilePitch, lane);   // S  = Q K^T
            warp_mma<LP, D, false, false>(dp, ad, S::kTilePitch, mt * 16, av, S::kTilePitch, lane);  // dP = dO V^T
            softmax_rows<LP>(s, L, scale, lane);
            float dot[2] = {0.f, 0.f};
#pragma unroll
            for (int nt = 0; nt < LP / 8; ++nt)
#pragma unroll
                for (int e = 0; e < 4; ++e) dot[e >> 1] += s[nt][e] * dp[nt][e];
#pragma unroll
            for (int hh = 0; hh < 2; ++hh) {
                dot[hh] += __shfl_xor_sync(0xffffffffu, dot[hh], 1);
                dot[hh] += __shfl_xor_sync(0xffffffffu, dot[hh], 2);
            }
#pragma unroll
            for (int nt = 0; nt < LP / 8; ++nt) {
                const int col = (nt * 8 + 2 * t) * 2;
                const int r0 = (mt * 16 + g) * S::kProbPitch, r1 = (mt * 16 + g + 8) * S::kProbPitch;
                *reinterpret_cast<uint32_t*>(sp + r0 + col) = pack_bf16(s[nt][0], s[nt][1]);
                *reinterpret_cast<uint32_t*>(sp + r1 + col) = pack_bf16(s[nt][2], s[nt][3]);
                *reinterpret_cast<uint32_t*>(ss + r0 + col) =
                    pack_bf16(s[nt][0] * (dp[nt][0] - dot[0]) * scale, s[nt][1] * (dp[nt][1] - dot[0]) * scale);
                *reinterpret_cast<uint32_t*>(ss + r1 + col) =
                    pack_bf16(s[nt][2] * (dp[nt][2] - dot[1]) * scale, s[nt][3] * (dp[nt][3] - dot[1]) * scale);
            }
        }
        __syncwarp();
        // three [LP x D] results, each staged to shared memory and written with coalesced 16-byte stores
#pragma unroll
        for (int which = 0; which < 3; ++which) {
#pragma unroll
            for (int mt = 0; mt < LP / 16; ++mt) {
                if (mt * 16 >= L) break;
                float acc[D / 8][4];
#pragma unroll
                for (int nt = 0; nt < D / 8; ++nt) acc[nt][0] = acc[nt][1] = acc[nt][2] = acc[nt][3] = 0.f;
                if (which == 0) warp_mma<D, LP, false, true>(acc, as, S::kProbPitch, mt * 16, ak, S::kTilePitch, lane);      // dQ = dS K
                else if (which == 1) warp_mma<D, LP, true, true>(acc, as, S::kProbPitch, mt * 16, aq, S::kTilePitch, lane);  // dK = dS^T Q
                else warp_mma<D, LP, true, true>(acc, ap, S::kProbPitch, mt * 16, ad, S::kTilePitch, lane);                  // dV = P^T dO
                stage_acc<D>(acc, so, S::kTilePitch, mt * 16, lane);
            }
            __syncwarp();
            store_tile<D>(which == 0 ? dq : (which == 1 ? dk : dv), base, sstr, L, so, lane);
            __syncwarp();
        }
    }
}

}  // namespace t2v

using namespace t2v;

namespace {

int check_args(int L, int D, int64_t ld_in, int64_t ld_out) {
    if (L < 1 || L > kMaxL) return fail(-2, "attn_small: L=%d out of range (1..%d)", L, kMaxL);
    if (D != 64 && D != 32) return fail(-2, "attn_small: head_dim %d unsupported (32 or 64)", D);
    if (ld_in % 8 || ld_out % 8) return fail(-2, "attn_small: row pitches must be multiples of 8 elements");
    return 0;
}

// warps per block so that a block stays near 64 KB of shared memory (3 blocks per SM)
int warps_for(int per_warp_bytes) { return std::max(1, std::min(8, (64 * 1024) / per_warp_bytes)); }

template <typename Kernel>
void set_smem_once(Kernel kernel, bool& done) {
    if (done) return;
    cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024);
    done = true;
}

template <int D, int LP>
int launch_fwd(const void* q, const void* k, const void* v, void* o, const SeqAddr& a, int64_t nseq, int heads, int L, cudaStream_t st) {
    static bool done = false;
    set_smem_once(attn_small_fwd_kernel<D, LP>, done);
    const int per_warp = AttnSmem<D, LP>::kFwdBytes;
    const int warps = warps_for(per_warp);
    const int64_t total = nseq * heads;
    const int grid = int(std::min<int64_t>((total + warps - 1) / warps, 148 * 8));
    return int(launch_pdl(attn_small_fwd_kernel<D, LP>, dim3(grid), dim3(warps * 32), size_t(per_warp) * warps, st,
                          static_cast<const __nv_bfloat16*>(q), static_cast<const __nv_bfloat16*>(k), static_cast<const __nv_bfloat16*>(v),
                          static_cast<__nv_bfloat16*>(o), a, nseq, heads, L, 1.0f / sqrtf(float(D))));
}

template <int D, int LP>
int launch_bwd(const void* q, const void* k, const void* v, const void* dout, void* dq, void* dk, void* dv, const SeqAddr& a,
               int64_t nseq, int heads, int L, cudaStream_t st) {
    static bool done = false;
    set_smem_once(attn_small_bwd_kernel<D, LP>, done);
    const int per_warp = AttnSmem<D, LP>::kBwdBytes;
    const int warps = warps_for(per_warp);
    const int64_t total = nseq * heads;
    const int grid = int(std::min<int64_t>((total + warps - 1) / warps, 148 * 8));
    auto B = [](const void* p) { return static_cast<const __nv_bfloat16*>(p); };
    auto W = [](void* p) { return static_cast<__nv_bfloat16*>(p); };
    return int(launch_pdl(attn_small_bwd_kernel<D, LP>, dim3(grid), dim3(warps * 32), size_t(per_warp) * warps, st, B(q), B(k), B(v),
                          B(dout), W(dq), W(dk), W(dv), a, nseq, heads, L, 1.0f / sqrtf(float(D))));
}

}  // namespace

extern "C" {

int t2v_attn_small_fwd(const void* q, const void* k, const void* v, void* o, int64_t nseq, int32_t inner, int64_t outer_rows,
                       int64_t inner_rows, int64_t seq_rows, int64_t ld_in, int64_t ld_out, int32_t heads, int32_t L, int32_t D,
                       void* stream) {
    if (int r = check_args(L, D, ld_in, ld_out)) return r;
    const SeqAddr a{outer_rows, inner_rows, seq_rows, ld_in, ld_out, inner};
    cudaStream_t st = static_cast<cudaStream_t>(stream);
    int rc;
    if (D == 64) rc = L <= 16 ? launch_fwd<64, 16>(q, k, v, o, a, nseq, heads, L, st) : launch_fwd<64, 32>(q, k, v, o, a, nseq, heads, L, st);
    else rc = L <= 16 ? launch_fwd<32, 16>(q, k, v, o, a, nseq, heads, L, st) : launch_fwd<32, 32>(q, k, v, o, a, nseq, heads, L, st);
    return launch_checked(rc, "attn_small_fwd");
}

int t2v_attn_small_bwd(const void* q, const void* k, const void* v, const void* dout, void* dq, void* dk, void* dv, int64_t nseq,
                       int32_t inner, int64_t outer_rows, int64_t inner_rows, int64_t seq_rows, int64_t ld_in, int64_t ld_out,
                       int32_t heads, int32_t L, int32_t D, void* stream) {
    if (int r = check_args(L, D, ld_in, ld_out)) return r;
    const SeqAddr a{outer_rows, inner_rows, seq_rows, ld_in, ld_out, inner};
    cudaStream_t st = static_cast<cudaStream_t>(stream);
    int rc;
    if (D == 64)
        rc = L <= 16 ? launch_bwd<64, 16>(q, k, v, dout, dq, dk, dv, a, nseq, heads, L, st)
                     : launch_bwd<64, 32>(q, k, v, dout, dq, dk, dv, a, nseq, heads, L, st);
    else
        rc = L <= 16 ? launch_bwd<32, 16>(q, k, v, dout, dq, dk, dv, a, nseq, heads, L, st)
                     : launch_bwd<32, 32>(q, k, v, dout, dq, dk, dv, a, nseq, heads, L, st);
    return launch_checked(rc, "attn_small_bwd");
}

}  // extern "C"
