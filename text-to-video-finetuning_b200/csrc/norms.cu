// GroupNorm (+SiLU) and LayerNorm, forward and backward, for channels-last bf16 activations.
// HBM-bound kernels: 128-bit loads, fp32 statistics, warp-shuffle / shared-memory reductions.
//
// GroupNorm works on x [S][P][C]: S normalisation samples (frames for the per-frame norms of ResnetBlock2D /
// Transformer2DModel, clips for the per-clip norms of TemporalConvLayer / TransformerTemporalModel), P pixels per
// sample, C channels in G groups.  Statistics are reduced in two levels (pixel chunks -> sample) so the grid fills
// the GPU even when S == 1.
#include "common.h"
#include "ptx.cuh"

#include <algorithm>
#include <cuda_bf16.h>

namespace t2v {

__device__ __forceinline__ void unpack8(const uint4& q, float* v) {
    v[0] = bf16_lo(q.x); v[1] = bf16_hi(q.x); v[2] = bf16_lo(q.y); v[3] = bf16_hi(q.y);
    v[4] = bf16_lo(q.z); v[5] = bf16_hi(q.z); v[6] = bf16_lo(q.w); v[7] = bf16_hi(q.w);
}
__device__ __forceinline__ uint4 pack8(const float* v) {
    uint4 q;
    q.x = pack_bf16(v[0], v[1]); q.y = pack_bf16(v[2], v[3]); q.z = pack_bf16(v[4], v[5]); q.w = pack_bf16(v[6], v[7]);
    return q;
}
__device__ __forceinline__ float sigmoidf_(float z) { return 1.0f / (1.0f + __expf(-z)); }

// ------------------------------------------------------------------------------------------------ GroupNorm
// Per-channel partial sums over a chunk of pixels.  MODE 0: (sum x, sum x^2).  MODE 1 (backward): (sum dz, sum dz*xhat)
// where dz = dy * silu'(a x + b) (or dy) and xhat = (x - mean) rstd.
// Thread layout: V = C/8 channel vectors; thread owns vector tid % V and pixel lane tid / V.
template <int MODE>
__global__ void gn_partial_kernel(const __nv_bfloat16* __restrict__ x, const __nv_bfloat16* __restrict__ dy,
                                  const float* __restrict__ ab,   // [S][C][2] (a,b) for z = a x + b   (MODE 1)
                                  const float* __restrict__ stat, // [S][G][2] (mean, rstd)            (MODE 1)
                                  float* __restrict__ partial,    // [S][chunks][C][2]
                                  int64_t P, int C, int G, int chunk_pixels, int silu) {
    extern __shared__ float sh[];  // [2][C]
    const int s = blockIdx.y, chunk = blockIdx.x, chunks = gridDim.x;
    const int V = C >> 3;
    const int lanes = blockDim.x / V;
    const int cv = threadIdx.x % V, pl = threadIdx.x / V;
    for (int i = threadIdx.x; i < 2 * C; i += blockDim.x) sh[i] = 0.f;
    __syncthreads();
    float acc0[8], acc1[8];
#pragma unroll
    for (int j = 0; j < 8; ++j) acc0[j] = acc1[j] = 0.f;
    const int64_t p0 = int64_t(chunk) * chunk_pixels;
    const int64_t p1 = min(P, p0 + chunk_pixels);
    if (pl < lanes) {
        float a[8], b[8], mean[8], rstd[8];
        if (MODE == 1) {
            const int cpg = C / G;
#pragma unroll
            for (int j = 0; j < 8; ++j) {
                const int c = cv * 8 + j;
                a[j] = ab[(int64_t(s) * C + c) * 2];
                b[j] = ab[(int64_t(s) * C + c) * 2 + 1];
                mean[j] = stat[(int64_t(s) * G + c / cpg) * 2];
                rstd[j] = stat[(int64_t(s) * G + c / cpg) * 2 + 1];
            }
        }
        const uint4* xs = reinterpret_cast<const uint4*>(x + (int64_t(s) * P) * C) + cv;
        const uint4* ds = MODE == 1 ? reinterpret_cast<const uint4*>(dy + (int64_t(s) * P) * C) + cv : nullptr;
        for (int64_t p = p0 + pl; p < p1; p += lanes) {
            float v[8];
            unpack8(__ldg(xs + p * V), v);
            if (MODE == 0) {
#pragma unroll
                for (int j = 0; j < 8; ++j) {
                    acc0[j] += v[j];
                    acc1[j] += v[j] * v[j];
                }
            } else {
                float d[8];
                unpack8(__ldg(ds + p * V), d);
#pragma unroll
                for (int j = 0; j < 8; ++j) {
                    float dz = d[j];
                    if (silu) {
                        const float z = a[j] * v[j] + b[j];
                        const float sg = sigmoidf_(z);
                        dz *= sg * (1.f + z * (1.f - sg));
                    }
                    acc0[j] += dz;
                    acc1[j] += dz * (v[j] - mean[j]) * rstd[j];
                }
            }
        }
#pragma unroll
        for (int j = 0; j < 8; ++j) {
            atomicAdd(&sh[cv * 8 + j], acc0[j]);
            atomicAdd(&sh[C + cv * 8 + j], acc1[j]);
        }
    }
    __syncthreads();
    float* out = partial + (int64_t(s) * chunks + chunk) * C * 2;
    for (int c = threadIdx.x; c < C; c += blockDim.x) {
        out[2 * c] = sh[c];
        out[2 * c + 1] = sh[C + c];
    }
}

// Forward finalize: one block per sample.  Reduces chunk partials -> (mean, rstd) per group and the per-channel
// affine (a, b) with y = act(a x + b).
__global__ void gn_fwd_finalize_kernel(const float* __restrict__ partial, const float* __restrict__ gamma,
                                       const float* __restrict__ beta, float* __restrict__ stat, float* __restrict__ ab,
                                       int64_t P, int C, int G, int chunks, float eps) {
    extern __shared__ float sh[];  // [2][C] then [2][G]
    const int s = blockIdx.x;
    float* gs = sh + 2 * C;
    for (int c = threadIdx.x; c < C; c += blockDim.x) {
        double a0 = 0, a1 = 0;
        for (int k = 0; k < chunks; ++k) {
            const float* pp = partial + ((int64_t(s) * chunks + k) * C + c) * 2;
            a0 += pp[0];
            a1 += pp[1];
        }
        sh[c] = float(a0);
        sh[C + c] = float(a1);
    }
    __syncthreads();
    const int cpg = C / G;
    for (int g = threadIdx.x; g < G; g += blockDim.x) {
        double a0 = 0, a1 = 0;
        for (int j = 0; j < cpg; ++j) {
            a0 += sh[g * cpg + j];
            a1 += sh[C + g * cpg + j];
        }
        const double n = double(P) * cpg;
        const double mean = a0 / n;
        double var = a1 / n - mean * mean;
        if (var < 0) var = 0;
        const float rstd = float(1.0 / sqrt(var + double(eps)));
        gs[g] = float(mean);
        gs[G + g] = rstd;
        stat[(int64_t(s) * G + g) * 2] = float(mean);
        stat[(int64_t(s) * G + g) * 2 + 1] = rstd;
    }
    __syncthreads();
    for (int c = threadIdx.x; c < C; c += blockDim.x) {
        const int g = c / cpg;
        const float a = gs[G + g] * gamma[c];
        ab[(int64_t(s) * C + c) * 2] = a;
        ab[(int64_t(s) * C + c) * 2 + 1] = beta[c] - gs[g] * a;
    }
}

// y = act(a x + b), elementwise over [S][P][C].
__global__ void gn_apply_kernel(const __nv_bfloat16* __restrict__ x, const float* __restrict__ ab,
                                __nv_bfloat16* __restrict__ y, int64_t P, int C, int64_t total_vec, int silu) {
    const int V = C >> 3;
    for (int64_t i = blockIdx.x * int64_t(blockDim.x) + threadIdx.x; i < total_vec; i += int64_t(gridDim.x) * blockDim.x) {
        const int cv = int(i % V);
        const int64_t s = (i / V) / P;
        float v[8];
        unpack8(__ldg(reinterpret_cast<const uint4*>(x) + i), v);
        const float4* abp = reinterpret_cast<const float4*>(ab + (s * C + cv * 8) * 2);
#pragma unroll
        for (int j = 0; j < 4; ++j) {
            const float4 q = __ldg(abp + j);
            float z0 = q.x * v[2 * j] + q.y, z1 = q.z * v[2 * j + 1] + q.w;
            if (silu) {
                z0 *= sigmoidf_(z0);
                z1 *= sigmoidf_(z1);
            }
            v[2 * j] = z0;
            v[2 * j + 1] = z1;
        }
        reinterpret_cast<uint4*>(y)[i] = pack8(v);
    }
}

// Backward finalize: one block per sample.  Produces the per-channel coefficients (pc, qc, rc) with
//   dx = pc * dz + qc * x + rc, and accumulates dgamma / dbeta.
__global__ void gn_bwd_finalize_kernel(const float* __restrict__ partial, const float* __restrict__ gamma,
                                       const float* __restrict__ stat, float* __restrict__ coef /*[S][C][4]*/,
                                       float* __restrict__ dgamma, float* __restrict__ dbeta, int64_t P, int C, int G,
                                       int chunks) {
    extern __shared__ float sh[];  // [2][C] then [2][G]
    const int s = blockIdx.x;
    float* gs = sh + 2 * C;
    for (int c = threadIdx.x; c < C; c += blockDim.x) {
        double a0 = 0, a1 = 0;
        for (int k = 0; k < chunks; ++k) {
            const float* pp = partial + ((int64_t(s) * chunks + k) * C + c) * 2;
            a0 += pp[0];
            a1 += pp[1];
        }
        sh[c] = float(a0);      // sum dz
        sh[C + c] = float(a1);  // sum dz * xhat
        if (dbeta) atomicAdd(dbeta + c, float(a0));
        if (dgamma) atomicAdd(dgamma + c, float(a1));
    }
    __syncthreads();
    const int cpg = C / G;
    for (int g = threadIdx.x; g < G; g += blockDim.x) {
        double s1 = 0, s2 = 0;
        for (int j = 0; j < cpg; ++j) {
            const int c = g * cpg + j;
            s1 += double(gamma[c]) * sh[c];
            s2 += double(gamma[c]) * sh[C + c];
        }
        gs[g] = float(s1);
        gs[G + g] = float(s2);
    }
    __syncthreads();
    for (int c = threadIdx.x; c < C; c += blockDim.x) {
        const int g = c / cpg;
        const float mean = stat[(int64_t(s) * G + g) * 2], rstd = stat[(int64_t(s) * G + g) * 2 + 1];
        const float invn = 1.0f / (float(P) * cpg);
        const float q = -rstd * rstd * gs[G + g] * invn;
        float* o = coef + (int64_t(s) * C + c) * 4;
        o[0] = rstd * gamma[c];
        o[1] = q;
        o[2] = -rstd * gs[g] * invn - q * mean;
        o[3] = 0.f;
    }
}

// dx = pc * dz + qc * x + rc (+ add), dz = dy * silu'(a x + b).
__global__ void gn_bwd_apply_kernel(const __nv_bfloat16* __restrict__ x, const __nv_bfloat16* __restrict__ dy,
                                    const float* __restrict__ ab, const float* __restrict__ coef,
                                    const __nv_bfloat16* __restrict__ add, __nv_bfloat16* __restrict__ dx, int64_t P, int C,
                                    int64_t total_vec, int silu) {
    const int V = C >> 3;
    for (int64_t i = blockIdx.x * int64_t(blockDim.x) + threadIdx.x; i < total_vec; i += int64_t(gridDim.x) * blockDim.x) {
        const int cv = int(i % V);
        const int64_t s = (i / V) / P;
        float v[8], d[8], r[8];
        unpack8(__ldg(reinterpret_cast<const uint4*>(x) + i), v);
        unpack8(__ldg(reinterpret_cast<const uint4*>(dy) + i), d);
        if (add) unpack8(__ldg(reinterpret_cast<const uint4*>(add) + i), r);
        const float2* abp = reinterpret_cast<const float2*>(ab + (s * C + cv * 8) * 2);
        const float4* cp = reinterpret_cast<const float4*>(coef + (s * C + cv * 8) * 4);
#pragma unroll
        for (int j = 0; j < 8; ++j) {
            float dz = d[j];
            if (silu) {
                const float2 q = __ldg(abp + j);
                const float z = q.x * v[j] + q.y;
                const float sg = sigmoidf_(z);
                dz *= sg * (1.f + z * (1.f - sg));
            }
            const float4 c4 = __ldg(cp + j);
            float o = c4.x * dz + c4.y * v[j] + c4.z;
            if (add) o += r[j];
            v[j] = o;
        }
        reinterpret_cast<uint4*>(dx)[i] = pack8(v);
    }
}

// ------------------------------------------------------------------------------------------------ LayerNorm
// One warp per row; the row lives in registers (VPL 16-byte vectors per lane), exact two-pass statistics.
template <int VPL>
__global__ void ln_fwd_kernel(const __nv_bfloat16* __restrict__ x, const float* __restrict__ gamma,
                              const float* __restrict__ beta, __nv_bfloat16* __restrict__ y, float* __restrict__ stat,
                              int64_t rows, int C, float eps) {
    const int V = C >> 3;
    const int lane = threadIdx.x & 31;
    const int64_t warp = (blockIdx.x * int64_t(blockDim.x) + threadIdx.x) >> 5;
    const int64_t nwarps = (int64_t(gridDim.x) * blockDim.x) >> 5;
    for (int64_t row = warp; row < rows; row += nwarps) {
        float v[VPL][8];
        float sum = 0.f;
#pragma unroll
        for (int k = 0; k < VPL; ++k) {
            const int cv = lane + 32 * k;
            if (cv < V) {
                unpack8(__ldg(reinterpret_cast<const uint4*>(x + row * C) + cv), v[k]);
#pragma unroll
                for (int j = 0; j < 8; ++j) sum += v[k][j];
            }
        }
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) sum += __shfl_xor_sync(0xffffffffu, sum, o);
        const float mean = sum / C;
        float sq = 0.f;
#pragma unroll
        for (int k = 0; k < VPL; ++k) {
            if (lane + 32 * k < V) {
#pragma unroll
                for (int j = 0; j < 8; ++j) {
                    const float d = v[k][j] - mean;
                    sq += d * d;
                }
            }
        }
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) sq += __shfl_xor_sync(0xffffffffu, sq, o);
        const float rstd = rsqrtf(sq / C + eps);
        if (lane == 0 && stat) {
            stat[row * 2] = mean;
            stat[row * 2 + 1] = rstd;
        }
#pragma unroll
        for (int k = 0; k < VPL; ++k) {
            const int cv = lane + 32 * k;
            if (cv < V) {
                float o[8];
#pragma unroll
                for (int j = 0; j < 8; ++j) o[j] = (v[k][j] - mean) * rstd * __ldg(gamma + cv * 8 + j) + __ldg(beta + cv * 8 + j);
                reinterpret_cast<uint4*>(y + row * C)[cv] = pack8(o);
            }
        }
    }
}

// dx = rstd (dy g - mean_c(dy g) - xhat mean_c(dy g xhat)) (+ add); dgamma += sum_rows dy xhat; dbeta += sum_rows dy.
template <int VPL>
__global__ void ln_bwd_kernel(const __nv_bfloat16* __restrict__ x, const __nv_bfloat16* __restrict__ dy,
                              const float* __restrict__ gamma, const float* __restrict__ stat,
                              const __nv_bfloat16* __restrict__ add, __nv_bfloat16* __restrict__ dx,
                              float* __restrict__ dgamma, float* __restrict__ dbeta, int64_t rows, int C) {
    const int V = C >> 3;
    const int lane = threadIdx.x & 31;
    const int64_t warp = (blockIdx.x * int64_t(blockDim.x) + threadIdx.x) >> 5;
    const int64_t nwarps = (int64_t(gridDim.x) * blockDim.x) >> 5;
    float gacc[VPL][8], bacc[VPL][8], gam[VPL][8];
#pragma unroll
    for (int k = 0; k < VPL; ++k)
#pragma unroll
        for (int j = 0; j < 8; ++j) {
            gacc[k][j] = bacc[k][j] = 0.f;
            const int cv = lane + 32 * k;
            gam[k][j] = cv < V ? __ldg(gamma + cv * 8 + j) : 0.f;
        }
    for (int64_t row = warp; row < rows; row += nwarps) {
        const float mean = stat[row * 2], rstd = stat[row * 2 + 1];
        float xh[VPL][8], dg[VPL][8];
        float s1 = 0.f, s2 = 0.f;
#pragma unroll
        for (int k = 0; k < VPL; ++k) {
            const int cv = lane + 32 * k;
            if (cv < V) {
                float v[8], d[8];
                unpack8(__ldg(reinterpret_cast<const uint4*>(x + row * C) + cv), v);
                unpack8(__ldg(reinterpret_cast<const uint4*>(dy + row * C) + cv), d);
#pragma unroll
                for (int j = 0; j < 8; ++j) {
                    xh[k][j] = (v[j] - mean) * rstd;
                    dg[k][j] = d[j] * gam[k][j];
                    s1 += dg[k][j];
                    s2 += dg[k][j] * xh[k][j];
                    gacc[k][j] += d[j] * xh[k][j];
                    bacc[k][j] += d[j];
                }
            }
        }
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) {
            s1 += __shfl_xor_sync(0xffffffffu, s1, o);
            s2 += __shfl_xor_sync(0xffffffffu, s2, o);
        }
        s1 /= C;
        s2 /= C;
#pragma unroll
        for (int k = 0; k < VPL; ++k) {
            const int cv = lane + 32 * k;
            if (cv < V) {
                float o[8], r[8];
                if (add) unpack8(__ldg(reinterpret_cast<const uint4*>(add + row * C) + cv), r);
#pragma unroll
                for (int j = 0; j < 8; ++j) {
                    o[j] = rstd * (dg[k][j] - s1 - xh[k][j] * s2);
                    if (add) o[j] += r[j];
                }
                reinterpret_cast<uint4*>(dx + row * C)[cv] = pack8(o);
            }
        }
    }
    // block-level reduction of the parameter gradients, then one atomic per channel per block
    extern __shared__ float sh[];  // [2][C]
    for (int i = threadIdx.x; i < 2 * C; i += blockDim.x) sh[i] = 0.f;
    __syncthreads();
#pragma unroll
    for (int k = 0; k < VPL; ++k) {
        const int cv = lane + 32 * k;
        if (cv < V) {
#pragma unroll
            for (int j = 0; j < 8; ++j) {
                atomicAdd(&sh[cv * 8 + j], gacc[k][j]);
                atomicAdd(&sh[C + cv * 8 + j], bacc[k][j]);
            }
        }
    }
    __syncthreads();
    for (int c = threadIdx.x; c < C; c += blockDim.x) {
        if (dgamma) atomicAdd(dgamma + c, sh[c]);
        if (dbeta) atomicAdd(dbeta + c, sh[C + c]);
    }
}

static int gn_block(int C) {
    const int V = C / 8;
    int b = 256;
    while (b < V) b += 32;
    return b;
}

static void gn_chunks(int S, int64_t P, int& chunk_pixels, int& chunks) {
    // aim for >= ~4 blocks per SM overall, at least 16 pixels per chunk
    const int64_t want = std::max<int64_t>(1, (4 * 148 + S - 1) / S);
    int64_t cp = std::max<int64_t>(16, (P + want - 1) / want);
    chunk_pixels = int(std::min<int64_t>(cp, P));
    chunks = int((P + chunk_pixels - 1) / chunk_pixels);
}

}  // namespace t2v

using namespace t2v;

extern "C" {

int64_t t2v_groupnorm_workspace_bytes(int32_t S, int64_t P, int32_t C) {
    int cp, ch;
    gn_chunks(S, P, cp, ch);
    return int64_t(S) * ch * C * 2 * sizeof(float) + int64_t(S) * C * 4 * sizeof(float) + 256;
}

int t2v_groupnorm_fwd(const void* x, const float* gamma, const float* beta, void* y, float* stat, float* ab, void* workspace,
                      int32_t S, int64_t P, int32_t C, int32_t G, float eps, int32_t silu, void* stream_) {
    if (C % 8 || C % G || C / 8 > 1024) return fail(-2, "groupnorm: C=%d G=%d unsupported", C, G);
    cudaStream_t st = static_cast<cudaStream_t>(stream_);
    int cp, chunks;
    gn_chunks(S, P, cp, chunks);
    float* partial = static_cast<float*>(workspace);
    const int bs = gn_block(C);
    gn_partial_kernel<0><<<dim3(chunks, S), bs, 2 * C * sizeof(float), st>>>(
        static_cast<const __nv_bfloat16*>(x), nullptr, nullptr, nullptr, partial, P, C, G, cp, 0);
    gn_fwd_finalize_kernel<<<S, 256, (2 * C + 2 * G) * sizeof(float), st>>>(partial, gamma, beta, stat, ab, P, C, G, chunks, eps);
    const int64_t tv = int64_t(S) * P * (C / 8);
    const int grid = int(std::min<int64_t>((tv + 255) / 256, 148 * 16));
    gn_apply_kernel<<<grid, 256, 0, st>>>(static_cast<const __nv_bfloat16*>(x), ab, static_cast<__nv_bfloat16*>(y), P, C, tv, silu);
    count_launch(2);
    return launch_checked(int(cudaGetLastError()), "groupnorm_fwd");
}

int t2v_groupnorm_bwd(const void* dy, const void* x, const float* gamma, const float* stat, const float* ab, const void* add,
                      void* dx, float* dgamma, float* dbeta, void* workspace, int32_t S, int64_t P, int32_t C, int32_t G,
                      int32_t silu, void* stream_) {
    if (C % 8 || C % G || C / 8 > 1024) return fail(-2, "groupnorm: C=%d G=%d unsupported", C, G);
    cudaStream_t st = static_cast<cudaStream_t>(stream_);
    int cp, chunks;
    gn_chunks(S, P, cp, chunks);
    float* partial = static_cast<float*>(workspace);
    float* coef = partial + int64_t(S) * chunks * C * 2;
    const int bs = gn_block(C);
    gn_partial_kernel<1><<<dim3(chunks, S), bs, 2 * C * sizeof(float), st>>>(
        static_cast<const __nv_bfloat16*>(x), static_cast<const __nv_bfloat16*>(dy), ab, stat, partial, P, C, G, cp, silu);
    gn_bwd_finalize_kernel<<<S, 256, (2 * C + 2 * G) * sizeof(float), st>>>(partial, gamma, stat, coef, dgamma, dbeta, P, C, G, chunks);
    const int64_t tv = int64_t(S) * P * (C / 8);
    const int grid = int(std::min<int64_t>((tv + 255) / 256, 148 * 16));
    gn_bwd_apply_kernel<<<grid, 256, 0, st>>>(static_cast<const __nv_bfloat16*>(x), static_cast<const __nv_bfloat16*>(dy), ab, coef,
                                              static_cast<const __nv_bfloat16*>(add), static_cast<__nv_bfloat16*>(dx), P, C, tv, silu);
    count_launch(2);
    return launch_checked(int(cudaGetLastError()), "groupnorm_bwd");
}

#define LN_DISPATCH(KERNEL, ...)                                                       \
    switch (vpl) {                                                                     \
        case 1: KERNEL<1> __VA_ARGS__; break;                                          \
        case 2: KERNEL<2> __VA_ARGS__; break;                                          \
        case 3: KERNEL<3> __VA_ARGS__; break;                                          \
        case 4: KERNEL<4> __VA_ARGS__; break;                                          \
        case 5: KERNEL<5> __VA_ARGS__; break;                                          \
        case 6: KERNEL<6> __VA_ARGS__; break;                                          \
        case 7: KERNEL<7> __VA_ARGS__; break;                                          \
        default: KERNEL<8> __VA_ARGS__; break;                                         \
    }

int t2v_layernorm_fwd(const void* x, const float* gamma, const float* beta, void* y, float* stat, int64_t rows, int32_t C,
                      float eps, void* stream_) {
    if (C % 8 || C > 2048) return fail(-2, "layernorm: C=%d unsupported (multiple of 8, <= 2048)", C);
    cudaStream_t st = static_cast<cudaStream_t>(stream_);
    const int vpl = (C / 8 + 31) / 32;
    const int grid = int(std::min<int64_t>((rows + 7) / 8, 148 * 8));
    LN_DISPATCH(ln_fwd_kernel, <<<grid, 256, 0, st>>>(static_cast<const __nv_bfloat16*>(x), gamma, beta,
                                                     static_cast<__nv_bfloat16*>(y), stat, rows, C, eps));
    return launch_checked(int(cudaGetLastError()), "layernorm_fwd");
}

int t2v_layernorm_bwd(const void* dy, const void* x, const float* gamma, const float* stat, const void* add, void* dx,
                      float* dgamma, float* dbeta, int64_t rows, int32_t C, void* stream_) {
    if (C % 8 || C > 2048) return fail(-2, "layernorm: C=%d unsupported (multiple of 8, <= 2048)", C);
    cudaStream_t st = static_cast<cudaStream_t>(stream_);
    const int vpl = (C / 8 + 31) / 32;
    const int grid = int(std::min<int64_t>((rows + 7) / 8, 148 * 2));
    LN_DISPATCH(ln_bwd_kernel, <<<grid, 256, 2 * C * sizeof(float), st>>>(
                    static_cast<const __nv_bfloat16*>(x), static_cast<const __nv_bfloat16*>(dy), gamma, stat,
                    static_cast<const __nv_bfloat16*>(add), static_cast<__nv_bfloat16*>(dx), dgamma, dbeta, rows, C));
    return launch_checked(int(cudaGetLastError()), "layernorm_bwd");
}

}  // extern "C"
