// Fused AdamW + global-norm clipping over the flat parameter arena (the reference builds
// torch.optim.AdamW at train.py:616-623 and clips with accelerator.clip_grad_norm_ at :868-876).
//
// Three kernels per optimizer step, all free of host-side scalars so that the step can sit inside the CUDA graph of the
// training step (learning rate, step count and clip factor live in device memory):
//   sqnorm   one read of the trainable gradient ranges -> sum of squares (fp64 accumulation)             4 B / param
//   prepare  one thread: step += 1, bias corrections, clip factor min(1, max_norm / (norm + 1e-6))       -
//   update   reads p, g, m, v, writes p, m, v, the bf16 compute shadow of p (kernel layout == master layout, so the
//            per-step cast kernel disappears) and zero into g (so the per-step gradient memset disappears)  34 B / param
// HBM-bound; algorithmic bytes = 38 per trainable parameter with clipping, 34 without.
//
// The trainable set is described by a CHUNK TABLE (int64 pairs: offset, length; offsets and lengths are multiples of 64
// elements, a chunk never straddles the matrix/vector boundary of the arena): one launch covers every parameter of a
// hyper-parameter set, however fragmented (LoRA: 1,148 small matrices between frozen base weights).
//
// Semantics = torch.optim.AdamW (decoupled weight decay, no amsgrad):
//   p *= 1 - lr * wd;  m = b1 m + (1 - b1) g;  v = b2 v + (1 - b2) g^2;  p -= (lr / bc1) * m / (sqrt(v) / sqrt(bc2) + eps)
// with g pre-multiplied by the clip factor.
#include "common.h"
#include "gemm_tc.cuh"
#include "ptx.cuh"

#include <algorithm>
#include <cmath>
#include <cuda_bf16.h>

namespace t2v {

// hp layout (floats): [0] lr [1] beta1 [2] beta2 [3] eps [4] weight_decay [5] bias_c1 [6] sqrt(bias_c2) [7] grad_scale
constexpr int kHp = 8;

struct AdamWArgs {
    float lr, beta1, beta2, eps, weight_decay, bias_c1, bias_c2_sqrt, grad_scale;
};

__device__ __forceinline__ void adamw_one(float& p, float g, float& m, float& v, const AdamWArgs& a) {
    g *= a.grad_scale;
    p *= 1.0f - a.lr * a.weight_decay;
    m = a.beta1 * m + (1.0f - a.beta1) * g;
    v = a.beta2 * v + (1.0f - a.beta2) * g * g;
    const float denom = sqrtf(v) / a.bias_c2_sqrt + a.eps;
    p -= (a.lr / a.bias_c1) * (m / denom);
}

// g16 != NULL: the gradient comes from the bf16 communication buffer (the all-reduced, averaged gradient of a data-parallel
// step); the fp32 accumulation buffer g is then only zeroed.
__global__ void __launch_bounds__(256) adamw_chunks_kernel(float* __restrict__ p, float* __restrict__ g, float* __restrict__ m,
                                                           float* __restrict__ v, __nv_bfloat16* __restrict__ shadow, int64_t n_shadow,
                                                           const int64_t* __restrict__ chunks, int n_chunks, const float* __restrict__ hp,
                                                           int zero_grad, const __nv_bfloat16* __restrict__ g16) {
    pdl_sync();
    AdamWArgs a;
    a.lr = hp[0]; a.beta1 = hp[1]; a.beta2 = hp[2]; a.eps = hp[3]; a.weight_decay = hp[4];
    a.bias_c1 = hp[5]; a.bias_c2_sqrt = hp[6]; a.grad_scale = hp[7];
    for (int c = blockIdx.x; c < n_chunks; c += gridDim.x) {
        const int64_t off = chunks[2 * c], len = chunks[2 * c + 1];
        const bool sh = shadow != nullptr && off < n_shadow;
        float4* p4 = reinterpret_cast<float4*>(p + off);
        float4* g4 = reinterpret_cast<float4*>(g + off);
        float4* m4 = reinterpret_cast<float4*>(m + off);
        float4* v4 = reinterpret_cast<float4*>(v + off);
        uint2* s2 = reinterpret_cast<uint2*>(shadow + off);
        const uint2* h2 = reinterpret_cast<const uint2*>(g16 + off);
        const int nv = int(len >> 2);
        for (int i = threadIdx.x; i < nv; i += blockDim.x) {
            float4 pp = p4[i];
            float4 gg;
            if (g16) {
                const uint2 h = __ldg(h2 + i);
                gg = make_float4(bf16_lo(h.x), bf16_hi(h.x), bf16_lo(h.y), bf16_hi(h.y));
            } else {
                gg = g4[i];
            }
            float4 mm = m4[i];
            float4 vv = v4[i];
            adamw_one(pp.x, gg.x, mm.x, vv.x, a);
            adamw_one(pp.y, gg.y, mm.y, vv.y, a);
            adamw_one(pp.z, gg.z, mm.z, vv.z, a);
            adamw_one(pp.w, gg.w, mm.w, vv.w, a);
            p4[i] = pp;
            m4[i] = mm;
            v4[i] = vv;
            if (sh) {
                uint2 q;
                q.x = pack_bf16(pp.x, pp.y);
                q.y = pack_bf16(pp.z, pp.w);
                s2[i] = q;
            }
            if (zero_grad) g4[i] = make_float4(0.f, 0.f, 0.f, 0.f);
        }
    }
}

__global__ void __launch_bounds__(256) sqnorm_chunks_kernel(const float* __restrict__ g, const int64_t* __restrict__ chunks, int n_chunks,
                                                            double* __restrict__ out, const __nv_bfloat16* __restrict__ g16) {
    pdl_sync();
    double acc = 0.0;
    for (int c = blockIdx.x; c < n_chunks; c += gridDim.x) {
        const int64_t off = chunks[2 * c], len = chunks[2 * c + 1];
        const float4* g4 = reinterpret_cast<const float4*>(g + off);
        const uint2* h2 = reinterpret_cast<const uint2*>(g16 + off);
        const int nv = int(len >> 2);
        float part = 0.f;  // one chunk is at most 64 K elements: 64 fp32 terms per thread, then fp64
        for (int i = threadIdx.x; i < nv; i += blockDim.x) {
            float4 q;
            if (g16) {
                const uint2 h = __ldg(h2 + i);
                q = make_float4(bf16_lo(h.x), bf16_hi(h.x), bf16_lo(h.y), bf16_hi(h.y));
            } else {
                q = __ldg(g4 + i);
            }
            part += q.x * q.x + q.y * q.y + q.z * q.z + q.w * q.w;
        }
        acc += double(part);
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) acc += __shfl_xor_sync(0xffffffffu, acc, o);
    __shared__ double wsum[8];
    if ((threadIdx.x & 31) == 0) wsum[threadIdx.x >> 5] = acc;
    __syncthreads();
    if (threadIdx.x == 0) {
        double t = 0.0;
        for (int w = 0; w < 8; ++w) t += wsum[w];
        atomicAdd(out, t);
    }
}

// hp_in: n_sets x 5 floats (lr, beta1, beta2, eps, weight_decay) written by the host before the step; hp: n_sets x kHp.
// state: [0] optimizer step count (int64)  ;  sq: [0] sum of squared gradients of this step (consumed and reset here),
// [1] the gradient norm of the last step (kept for logging)
__global__ void adamw_prepare_kernel(const float* __restrict__ hp_in, float* __restrict__ hp, int n_sets, int64_t* __restrict__ state,
                                     double* __restrict__ sq, float max_norm) {
    pdl_sync();
    if (threadIdx.x != 0 || blockIdx.x != 0) return;
    const int64_t step = state[0] + 1;
    state[0] = step;
    const double norm = sqrt(sq[0]);
    sq[1] = norm;
    sq[0] = 0.0;
    float scale = 1.0f;
    if (max_norm > 0.f) scale = fminf(1.0f, max_norm / (float(norm) + 1e-6f));
    for (int s = 0; s < n_sets; ++s) {
        const float* in = hp_in + 5 * s;
        float* o = hp + kHp * s;
        o[0] = in[0]; o[1] = in[1]; o[2] = in[2]; o[3] = in[3]; o[4] = in[4];
        o[5] = float(1.0 - pow(double(in[1]), double(step)));
        o[6] = float(sqrt(1.0 - pow(double(in[2]), double(step))));
        o[7] = scale;
    }
}

__global__ void counter_add_kernel(int64_t* p, int64_t v) {
    pdl_sync();
    if (threadIdx.x == 0 && blockIdx.x == 0) *p += v;
}

}  // namespace t2v

using namespace t2v;

extern "C" {

static int chunk_grid(int n_chunks) { return std::max(1, std::min(n_chunks, device_sm_count() * 8)); }

int t2v_sqnorm_chunks(const float* g, const void* g_bf16, const int64_t* chunks, int32_t n_chunks, double* out, void* stream) {
    if (n_chunks <= 0) return 0;
    const int rc = int(launch_pdl(sqnorm_chunks_kernel, dim3(chunk_grid(n_chunks)), dim3(256), size_t(0), static_cast<cudaStream_t>(stream), g,
                                  chunks, int(n_chunks), out, static_cast<const __nv_bfloat16*>(g_bf16)));
    return launch_checked(rc, "sqnorm_chunks");
}

int t2v_adamw_prepare(const float* hp_in, float* hp, int32_t n_sets, int64_t* state, double* sq, float max_norm, void* stream) {
    if (n_sets <= 0) return fail(-2, "adamw_prepare: no hyper-parameter sets");
    const int rc = int(launch_pdl(adamw_prepare_kernel, dim3(1), dim3(32), size_t(0), static_cast<cudaStream_t>(stream), hp_in, hp, int(n_sets),
                                  state, sq, max_norm));
    return launch_checked(rc, "adamw_prepare");
}

int t2v_adamw_chunks(float* p, float* g, const void* g_bf16, float* m, float* v, void* shadow_bf16, int64_t n_shadow, const int64_t* chunks,
                     int32_t n_chunks, const float* hp, int32_t zero_grad, void* stream) {
    if (n_chunks <= 0) return 0;
    if ((reinterpret_cast<uintptr_t>(p) | reinterpret_cast<uintptr_t>(g) | reinterpret_cast<uintptr_t>(m) | reinterpret_cast<uintptr_t>(v)) & 15u)
        return fail(-2, "adamw_chunks: p, g, m, v must be 16-byte aligned");
    if (shadow_bf16 && (reinterpret_cast<uintptr_t>(shadow_bf16) & 7u)) return fail(-2, "adamw_chunks: shadow must be 8-byte aligned");
    const int rc = int(launch_pdl(adamw_chunks_kernel, dim3(chunk_grid(n_chunks)), dim3(256), size_t(0), static_cast<cudaStream_t>(stream), p, g, m,
                                  v, static_cast<__nv_bfloat16*>(shadow_bf16), n_shadow, chunks, int(n_chunks), hp, int(zero_grad),
                                  static_cast<const __nv_bfloat16*>(g_bf16)));
    return launch_checked(rc, "adamw_chunks");
}

int t2v_counter_add(int64_t* counter, int64_t value, void* stream) {
    const int rc = int(launch_pdl(counter_add_kernel, dim3(1), dim3(32), size_t(0), static_cast<cudaStream_t>(stream), counter, value));
    return launch_checked(rc, "counter_add");
}

}  // extern "C"
