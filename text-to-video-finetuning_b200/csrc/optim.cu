// Fused AdamW over a contiguous range of the flat parameter arena ("next" after the hot path).
// One pass reads p, g, m, v (fp32) and writes p, m, v, the bf16 compute shadow of p (kernel layout == master layout, so the
// per-step cast kernel disappears) and, optionally, zero into g (so the per-step gradient memset disappears): 34 bytes per
// parameter instead of 28 + 6 + 4 in three passes.  HBM-bound; algorithmic bytes = 34 n.
// Semantics = torch.optim.AdamW (decoupled weight decay, no amsgrad) as used by the reference (train.py:616-623):
//   p *= 1 - lr * wd;  m = b1 m + (1 - b1) g;  v = b2 v + (1 - b2) g^2;  p -= (lr / bc1) * m / (sqrt(v) / sqrt(bc2) + eps)
// with g pre-multiplied by grad_scale (gradient clipping folded in: scale = min(1, max_norm / (norm + 1e-6))).
// STATUS: restated in oracle/ops_ref.py and checked on CPU against torch.optim.AdamW; the CUDA kernel itself has not
// run on a GPU yet (opt-in: train.main(fused_adamw=True), tests/test_fused_adamw.py with T2V_TEST_OPTIN=1).
#include "common.h"
#include "ptx.cuh"

#include <algorithm>
#include <cmath>
#include <cuda_bf16.h>

namespace t2v {

struct AdamWArgs {
    float lr, beta1, beta2, eps, weight_decay, bias_c1, bias_c2_sqrt, grad_scale;
    int zero_grad;
};

__device__ __forceinline__ float adamw_one(float& p, float g, float& m, float& v, const AdamWArgs& a) {
    g *= a.grad_scale;
    p *= 1.0f - a.lr * a.weight_decay;
    m = a.beta1 * m + (1.0f - a.beta1) * g;
    v = a.beta2 * v + (1.0f - a.beta2) * g * g;
    const float denom = sqrtf(v) / a.bias_c2_sqrt + a.eps;
    p -= (a.lr / a.bias_c1) * (m / denom);
    return p;
}

__global__ void adamw_kernel(float* __restrict__ p, float* __restrict__ g, float* __restrict__ m, float* __restrict__ v,
                             __nv_bfloat16* __restrict__ shadow, int64_t n, AdamWArgs a) {
    pdl_sync();
    const int64_t nv = n >> 2;
    for (int64_t i = blockIdx.x * int64_t(blockDim.x) + threadIdx.x; i < nv; i += int64_t(gridDim.x) * blockDim.x) {
        float4 pp = reinterpret_cast<float4*>(p)[i];
        const float4 gg = reinterpret_cast<const float4*>(g)[i];
        float4 mm = reinterpret_cast<float4*>(m)[i];
        float4 vv = reinterpret_cast<float4*>(v)[i];
        adamw_one(pp.x, gg.x, mm.x, vv.x, a);
        adamw_one(pp.y, gg.y, mm.y, vv.y, a);
        adamw_one(pp.z, gg.z, mm.z, vv.z, a);
        adamw_one(pp.w, gg.w, mm.w, vv.w, a);
        reinterpret_cast<float4*>(p)[i] = pp;
        reinterpret_cast<float4*>(m)[i] = mm;
        reinterpret_cast<float4*>(v)[i] = vv;
        if (shadow) {
            uint2 q;
            q.x = pack_bf16(pp.x, pp.y);
            q.y = pack_bf16(pp.z, pp.w);
            reinterpret_cast<uint2*>(shadow)[i] = q;
        }
        if (a.zero_grad) reinterpret_cast<float4*>(g)[i] = make_float4(0.f, 0.f, 0.f, 0.f);
    }
    // tail (n not a multiple of 4): first block, scalar
    if (blockIdx.x == 0 && threadIdx.x < (n & 3)) {
        const int64_t i = (nv << 2) + threadIdx.x;
        float pp = p[i], mm = m[i], vv = v[i];
        adamw_one(pp, g[i], mm, vv, a);
        p[i] = pp; m[i] = mm; v[i] = vv;
        if (shadow) shadow[i] = __float2bfloat16_rn(pp);
        if (a.zero_grad) g[i] = 0.f;
    }
}

}  // namespace t2v

using namespace t2v;

extern "C" {

int t2v_adamw_step(float* p, float* g, float* m, float* v, void* shadow_bf16, int64_t n, float lr, float beta1, float beta2, float eps,
                   float weight_decay, int32_t step, float grad_scale, int32_t zero_grad, void* stream) {
    if (n <= 0) return 0;
    if (step < 1) return fail(-2, "adamw_step: step counts from 1");
    if ((reinterpret_cast<uintptr_t>(p) | reinterpret_cast<uintptr_t>(g) | reinterpret_cast<uintptr_t>(m) | reinterpret_cast<uintptr_t>(v)) & 15u)
        return fail(-2, "adamw_step: p, g, m, v must be 16-byte aligned");
    if (shadow_bf16 && (reinterpret_cast<uintptr_t>(shadow_bf16) & 7u)) return fail(-2, "adamw_step: shadow must be 8-byte aligned");
    AdamWArgs a;
    a.lr = lr; a.beta1 = beta1; a.beta2 = beta2; a.eps = eps; a.weight_decay = weight_decay;
    a.bias_c1 = float(1.0 - std::pow(double(beta1), double(step)));
    a.bias_c2_sqrt = float(std::sqrt(1.0 - std::pow(double(beta2), double(step))));
    a.grad_scale = grad_scale;
    a.zero_grad = zero_grad;
    const int grid = int(std::min<int64_t>(((n >> 2) + 255) / 256 + 1, 148 * 16));
    const int rc = int(launch_pdl(adamw_kernel, dim3(grid), dim3(256), size_t(0), static_cast<cudaStream_t>(stream), p, g, m, v,
                                  static_cast<__nv_bfloat16*>(shadow_bf16), n, a));
    return launch_checked(rc, "adamw_step");
}

}  // extern "C"
