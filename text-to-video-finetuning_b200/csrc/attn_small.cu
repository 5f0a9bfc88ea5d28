// Short-sequence self-attention (L <= 32) for TransformerTemporalModel: attention along the frame axis.
//
// The reference permutes (B,C,F,H,W) -> (B*H*W, F, C) before this attention (diffusers TransformerTemporalModel,
// wired at unet_3d_blocks.py:331-340,491-500 and unet_3d_condition.py:147-152).  Here activations stay in the
// frames-major token order [B][F][H*W][C]; a sequence is addressed with strides instead:
//   token t of sequence z, head h  ->  base + (z / inner) * outer_stride + (z % inner) * inner_stride + t * seq_stride + h * D
// so both permute copies disappear.  One warp owns one (sequence, head): q/k/v (L x D) are staged in shared memory
// as fp32, scores/softmax use warp shuffles, everything else is registers.  The kernel is HBM-bound (reads q,k,v
// once, writes o once).
#include "common.h"
#include "ptx.cuh"

#include <algorithm>
#include <cuda_bf16.h>

namespace t2v {

constexpr int kWarpsPerBlock = 4;
constexpr int kMaxL = 32;

struct SeqAddr {
    int64_t outer_stride, inner_stride, seq_stride;
    int32_t inner;
};

__device__ __forceinline__ int64_t seq_base(const SeqAddr& a, int64_t z, int h, int D) {
    return (z / a.inner) * a.outer_stride + (z % a.inner) * a.inner_stride + int64_t(h) * D;
}

template <int D>
__device__ __forceinline__ void load_tile(const __nv_bfloat16* __restrict__ g, int64_t base, int64_t stride, int L, float* sm,
                                          int lane) {
    // D/2 bf16 pairs per token; lane covers pairs lane, lane+32, ...
    for (int t = 0; t < L; ++t) {
        const __nv_bfloat162* row = reinterpret_cast<const __nv_bfloat162*>(g + base + t * stride);
#pragma unroll
        for (int i = lane; i < D / 2; i += 32) {
            const float2 f = __bfloat1622float2(row[i]);
            sm[t * (D + 1) + 2 * i] = f.x;
            sm[t * (D + 1) + 2 * i + 1] = f.y;
        }
    }
}

template <int D>
__global__ void attn_small_fwd_kernel(const __nv_bfloat16* __restrict__ q, const __nv_bfloat16* __restrict__ k,
                                      const __nv_bfloat16* __restrict__ v, __nv_bfloat16* __restrict__ o, SeqAddr a, int64_t nseq,
                                      int heads, int L, float scale) {
    extern __shared__ float sm_all[];
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    float* sq = sm_all + warp * 3 * kMaxL * (D + 1);
    float* sk = sq + kMaxL * (D + 1);
    float* sv = sk + kMaxL * (D + 1);
    const int64_t total = nseq * heads;
    for (int64_t w = blockIdx.x * int64_t(kWarpsPerBlock) + warp; w < total; w += int64_t(gridDim.x) * kWarpsPerBlock) {
        const int64_t z = w / heads;
        const int h = int(w % heads);
        const int64_t base = seq_base(a, z, h, D);
        __syncwarp();
        load_tile<D>(q, base, a.seq_stride, L, sq, lane);
        load_tile<D>(k, base, a.seq_stride, L, sk, lane);
        load_tile<D>(v, base, a.seq_stride, L, sv, lane);
        __syncwarp();
        for (int i = 0; i < L; ++i) {
            float s = -INFINITY;
            if (lane < L) {
                float acc = 0.f;
#pragma unroll 8
                for (int d = 0; d < D; ++d) acc += sq[i * (D + 1) + d] * sk[lane * (D + 1) + d];
                s = acc * scale;
            }
            float mx = s;
#pragma unroll
            for (int off = 16; off > 0; off >>= 1) mx = fmaxf(mx, __shfl_xor_sync(0xffffffffu, mx, off));
            float p = lane < L ? __expf(s - mx) : 0.f;
            float sum = p;
#pragma unroll
            for (int off = 16; off > 0; off >>= 1) sum += __shfl_xor_sync(0xffffffffu, sum, off);
            p /= sum;
            float acc[D / 32];
#pragma unroll
            for (int r = 0; r < D / 32; ++r) acc[r] = 0.f;
            for (int j = 0; j < L; ++j) {
                const float pj = __shfl_sync(0xffffffffu, p, j);
#pragma unroll
                for (int r = 0; r < D / 32; ++r) acc[r] += pj * sv[j * (D + 1) + lane + 32 * r];
            }
            __nv_bfloat16* orow = o + base + i * a.seq_stride;
#pragma unroll
            for (int r = 0; r < D / 32; ++r) orow[lane + 32 * r] = __float2bfloat16_rn(acc[r]);
        }
    }
}

template <int D>
__global__ void attn_small_bwd_kernel(const __nv_bfloat16* __restrict__ q, const __nv_bfloat16* __restrict__ k,
                                      const __nv_bfloat16* __restrict__ v, const __nv_bfloat16* __restrict__ dout,
                                      __nv_bfloat16* __restrict__ dq, __nv_bfloat16* __restrict__ dk, __nv_bfloat16* __restrict__ dv,
                                      SeqAddr a, int64_t nseq, int heads, int L, float scale) {
    extern __shared__ float sm_all[];
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    float* sq = sm_all + warp * 4 * kMaxL * (D + 1);
    float* sk = sq + kMaxL * (D + 1);
    float* sv = sk + kMaxL * (D + 1);
    float* sd = sv + kMaxL * (D + 1);
    const int64_t total = nseq * heads;
    for (int64_t w = blockIdx.x * int64_t(kWarpsPerBlock) + warp; w < total; w += int64_t(gridDim.x) * kWarpsPerBlock) {
        const int64_t z = w / heads;
        const int h = int(w % heads);
        const int64_t base = seq_base(a, z, h, D);
        __syncwarp();
        load_tile<D>(q, base, a.seq_stride, L, sq, lane);
        load_tile<D>(k, base, a.seq_stride, L, sk, lane);
        load_tile<D>(v, base, a.seq_stride, L, sv, lane);
        load_tile<D>(dout, base, a.seq_stride, L, sd, lane);
        __syncwarp();
        // lane j accumulates dK_j and dV_j rows in registers (D values each)
        float dkj[D], dvj[D];
#pragma unroll
        for (int d = 0; d < D; ++d) dkj[d] = dvj[d] = 0.f;
        for (int i = 0; i < L; ++i) {
            float s = -INFINITY, dp = 0.f;
            if (lane < L) {
                float acc = 0.f, accp = 0.f;
#pragma unroll 8
                for (int d = 0; d < D; ++d) {
                    acc += sq[i * (D + 1) + d] * sk[lane * (D + 1) + d];
                    accp += sd[i * (D + 1) + d] * sv[lane * (D + 1) + d];
                }
                s = acc * scale;
                dp = accp;
            }
            float mx = s;
#pragma unroll
            for (int off = 16; off > 0; off >>= 1) mx = fmaxf(mx, __shfl_xor_sync(0xffffffffu, mx, off));
            float p = lane < L ? __expf(s - mx) : 0.f;
            float sum = p;
#pragma unroll
            for (int off = 16; off > 0; off >>= 1) sum += __shfl_xor_sync(0xffffffffu, sum, off);
            p /= sum;
            float dot = p * dp;
#pragma unroll
            for (int off = 16; off > 0; off >>= 1) dot += __shfl_xor_sync(0xffffffffu, dot, off);
            const float ds = p * (dp - dot) * scale;  // dS_ij (already times the softmax scale)
            if (lane < L) {
#pragma unroll
                for (int d = 0; d < D; ++d) {
                    dkj[d] += ds * sq[i * (D + 1) + d];
                    dvj[d] += p * sd[i * (D + 1) + d];
                }
            }
            // dQ_i = sum_j dS_ij k_j : lane owns dims lane, lane+32
            float acc[D / 32];
#pragma unroll
            for (int r = 0; r < D / 32; ++r) acc[r] = 0.f;
            for (int j = 0; j < L; ++j) {
                const float dsj = __shfl_sync(0xffffffffu, ds, j);
#pragma unroll
                for (int r = 0; r < D / 32; ++r) acc[r] += dsj * sk[j * (D + 1) + lane + 32 * r];
            }
            __nv_bfloat16* qrow = dq + base + i * a.seq_stride;
#pragma unroll
            for (int r = 0; r < D / 32; ++r) qrow[lane + 32 * r] = __float2bfloat16_rn(acc[r]);
        }
        if (lane < L) {
            __nv_bfloat162* krow = reinterpret_cast<__nv_bfloat162*>(dk + base + lane * a.seq_stride);
            __nv_bfloat162* vrow = reinterpret_cast<__nv_bfloat162*>(dv + base + lane * a.seq_stride);
#pragma unroll
            for (int d = 0; d < D / 2; ++d) {
                krow[d] = __floats2bfloat162_rn(dkj[2 * d], dkj[2 * d + 1]);
                vrow[d] = __floats2bfloat162_rn(dvj[2 * d], dvj[2 * d + 1]);
            }
        }
    }
}

}  // namespace t2v

using namespace t2v;

extern "C" {

int t2v_attn_small_fwd(const void* q, const void* k, const void* v, void* o, int64_t nseq, int32_t inner, int64_t outer_stride,
                       int64_t inner_stride, int64_t seq_stride, int32_t heads, int32_t L, int32_t D, void* stream) {
    if (L < 1 || L > kMaxL) return fail(-2, "attn_small: L=%d out of range (1..%d)", L, kMaxL);
    if (D != 64 && D != 32) return fail(-2, "attn_small: head_dim %d unsupported (32 or 64)", D);
    SeqAddr a{outer_stride, inner_stride, seq_stride, inner};
    const int64_t total = nseq * heads;
    const int grid = int(std::min<int64_t>((total + kWarpsPerBlock - 1) / kWarpsPerBlock, 148 * 8));
    const size_t smem = size_t(kWarpsPerBlock) * 3 * kMaxL * (D + 1) * sizeof(float);
    const float scale = 1.0f / sqrtf(float(D));
    auto Q = static_cast<const __nv_bfloat16*>(q);
    auto K = static_cast<const __nv_bfloat16*>(k);
    auto V = static_cast<const __nv_bfloat16*>(v);
    auto O = static_cast<__nv_bfloat16*>(o);
    cudaStream_t st = static_cast<cudaStream_t>(stream);
    static bool attr_done = false;
    if (!attr_done) {
        cudaFuncSetAttribute(attn_small_fwd_kernel<64>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                             int(size_t(kWarpsPerBlock) * 3 * kMaxL * 65 * sizeof(float)));
        cudaFuncSetAttribute(attn_small_fwd_kernel<32>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                             int(size_t(kWarpsPerBlock) * 3 * kMaxL * 33 * sizeof(float)));
        attr_done = true;
    }
    if (D == 64) {
        attn_small_fwd_kernel<64><<<grid, kWarpsPerBlock * 32, smem, st>>>(Q, K, V, O, a, nseq, heads, L, scale);
    } else {
        attn_small_fwd_kernel<32><<<grid, kWarpsPerBlock * 32, smem, st>>>(Q, K, V, O, a, nseq, heads, L, scale);
    }
    return launch_checked(int(cudaGetLastError()), "attn_small_fwd");
}

int t2v_attn_small_bwd(const void* q, const void* k, const void* v, const void* dout, void* dq, void* dk, void* dv, int64_t nseq,
                       int32_t inner, int64_t outer_stride, int64_t inner_stride, int64_t seq_stride, int32_t heads, int32_t L,
                       int32_t D, void* stream) {
    if (L < 1 || L > kMaxL) return fail(-2, "attn_small: L=%d out of range (1..%d)", L, kMaxL);
    if (D != 64 && D != 32) return fail(-2, "attn_small: head_dim %d unsupported (32 or 64)", D);
    SeqAddr a{outer_stride, inner_stride, seq_stride, inner};
    const int64_t total = nseq * heads;
    const int grid = int(std::min<int64_t>((total + kWarpsPerBlock - 1) / kWarpsPerBlock, 148 * 8));
    const size_t smem = size_t(kWarpsPerBlock) * 4 * kMaxL * (D + 1) * sizeof(float);
    const float scale = 1.0f / sqrtf(float(D));
    auto B = [](const void* p) { return static_cast<const __nv_bfloat16*>(p); };
    auto W = [](void* p) { return static_cast<__nv_bfloat16*>(p); };
    cudaStream_t st = static_cast<cudaStream_t>(stream);
    static bool attr_done = false;
    if (!attr_done) {
        cudaFuncSetAttribute(attn_small_bwd_kernel<64>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                             int(size_t(kWarpsPerBlock) * 4 * kMaxL * 65 * sizeof(float)));
        cudaFuncSetAttribute(attn_small_bwd_kernel<32>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                             int(size_t(kWarpsPerBlock) * 4 * kMaxL * 33 * sizeof(float)));
        attr_done = true;
    }
    if (D == 64) {
        attn_small_bwd_kernel<64><<<grid, kWarpsPerBlock * 32, smem, st>>>(B(q), B(k), B(v), B(dout), W(dq), W(dk), W(dv), a, nseq, heads, L, scale);
    } else {
        attn_small_bwd_kernel<32><<<grid, kWarpsPerBlock * 32, smem, st>>>(B(q), B(k), B(v), B(dout), W(dq), W(dk), W(dv), a, nseq, heads, L, scale);
    }
    return launch_checked(int(cudaGetLastError()), "attn_small_bwd");
}

}  // extern "C"
