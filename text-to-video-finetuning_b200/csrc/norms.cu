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
// Two kernels per direction (plus one memset of the [S][C][2] accumulator):
//   gn_partial_kernel : per-channel sums over a chunk of pixels, reduced in shared memory, then ONE atomicAdd per
//                       channel per block into accum[S][C][2].   MODE 0: (sum x, sum x^2); MODE 1: (sum dz, sum dz*xhat)
//   gn_*_apply_kernel : every block first finalises its sample's statistics / coefficients from accum in shared memory
//                       (2C..5C floats, fp64 for the group combine), then streams its own chunk of pixels.
// Thread layout of the streaming loops: V = C/8 channel vectors; thread owns vector tid % V and pixel lane tid / V.
template <int MODE>
__global__ void gn_partial_kernel(const __nv_bfloat16* __restrict__ x, const __nv_bfloat16* __restrict__ dy,
                                  const float* __restrict__ ab,   // [S][C][2] (a,b) for z = a x + b   (MODE 1)
                                  const float* __restrict__ stat, // [S][G][2] (mean, rstd)            (MODE 1)
                                  float* __restrict__ accum,      // [S][C][2]
                                  int64_t P, int C, int G, int chunk_pixels, int silu) {
    extern __shared__ float sh[];  // [2][C]
    const int s = blockIdx.y, chunk = blockIdx.x;
    const int V = C >> 3;
    const int lanes = blockDim.x / V;
    const int cv = threadIdx.x % V, pl = threadIdx.x / V;
    for (int i = threadIdx.x; i < 2 * C; i += blockDim.x) sh[i] = 0.f;
    __syncthreads();
    const int64_t p0 = int64_t(chunk) * chunk_pixels;
    const int64_t p1 = min(P, p0 + chunk_pixels);
    if (pl < lanes) {
        float acc0[8], acc1[8];
#pragma unroll
        for (int j = 0; j < 8; ++j) acc0[j] = acc1[j] = 0.f;
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
    float* out = accum + int64_t(s) * C * 2;
    for (int c = threadIdx.x; c < C; c += blockDim.x) {
        atomicAdd(out + 2 * c, sh[c]);
        atomicAdd(out + 2 * c + 1, sh[C + c]);
    }
}

// Forward apply: finalise (mean, rstd) per group and the per-channel affine (a, b) in shared memory, then y = act(a x + b).
__global__ void gn_fwd_apply_kernel(const __nv_bfloat16* __restrict__ x, const float* __restrict__ accum,
                                    const float* __restrict__ gamma, const float* __restrict__ beta, __nv_bfloat16* __restrict__ y,
                                    float* __restrict__ stat, float* __restrict__ ab, int64_t P, int C, int G, int chunk_pixels,
                                    float eps, int silu) {
    extern __shared__ float sh[];  // a[C], b[C], gmean[G], grstd[G]
    float* sa = sh;
    float* sb = sh + C;
    float* gm = sh + 2 * C;
    float* gr = gm + G;
    const int s = blockIdx.y;
    const int cpg = C / G;
    const float* acc = accum + int64_t(s) * C * 2;
    for (int g = threadIdx.x; g < G; g += blockDim.x) {
        double a0 = 0, a1 = 0;
        for (int j = 0; j < cpg; ++j) {
            a0 += acc[2 * (g * cpg + j)];
            a1 += acc[2 * (g * cpg + j) + 1];
        }
        const double n = double(P) * cpg;
        const double mean = a0 / n;
        double var = a1 / n - mean * mean;
        if (var < 0) var = 0;
        const float rstd = float(1.0 / sqrt(var + double(eps)));
        gm[g] = float(mean);
        gr[g] = rstd;
        if (blockIdx.x == 0) {
            stat[(int64_t(s) * G + g) * 2] = float(mean);
            stat[(int64_t(s) * G + g) * 2 + 1] = rstd;
        }
    }
    __syncthreads();
    for (int c = threadIdx.x; c < C; c += blockDim.x) {
        const int g = c / cpg;
        const float a = gr[g] * gamma[c];
        const float b = beta[c] - gm[g] * a;
        sa[c] = a;
        sb[c] = b;
        if (blockIdx.x == 0) {
            ab[(int64_t(s) * C + c) * 2] = a;
            ab[(int64_t(s) * C + c) * 2 + 1] = b;
        }
    }
    __syncthreads();
    const int V = C >> 3;
    const int64_t p0 = int64_t(blockIdx.x) * chunk_pixels, p1 = min(P, p0 + chunk_pixels);
    const int64_t nvec = (p1 - p0) * V;
    const uint4* xs = reinterpret_cast<const uint4*>(x + (int64_t(s) * P + p0) * C);
    uint4* ys = reinterpret_cast<uint4*>(y + (int64_t(s) * P + p0) * C);
    for (int64_t i = threadIdx.x; i < nvec; i += blockDim.x) {
        const int cv = int(i % V);
        float v[8];
        unpack8(__ldg(xs + i), v);
#pragma unroll
        for (int j = 0; j < 8; ++j) {
            float z = sa[cv * 8 + j] * v[j] + sb[cv * 8 + j];
            if (silu) z *= sigmoidf_(z);
            v[j] = z;
        }
        ys[i] = pack8(v);
    }
}

// Backward apply: dx = pc * dz + qc * x + rc (+ add), dz = dy * silu'(a x + b); block 0 of each sample also
// accumulates dgamma / dbeta.
__global__ void gn_bwd_apply_kernel(const __nv_bfloat16* __restrict__ x, const __nv_bfloat16* __restrict__ dy,
                                    const float* __restrict__ accum, const float* __restrict__ gamma, const float* __restrict__ stat,
                                    const float* __restrict__ ab, const __nv_bfloat16* __restrict__ add, __nv_bfloat16* __restrict__ dx,
                                    float* __restrict__ dgamma, float* __restrict__ dbeta, int64_t P, int C, int G, int chunk_pixels,
                                    int silu) {
    extern __shared__ float sh[];  // pc[C], qc[C], rc[C], a[C], b[C], s1[G], s2[G]
    float* pc = sh;
    float* qc = sh + C;
    float* rc = sh + 2 * C;
    float* sa = sh + 3 * C;
    float* sb = sh + 4 * C;
    float* g1 = sh + 5 * C;
    float* g2 = g1 + G;
    const int s = blockIdx.y;
    const int cpg = C / G;
    const float* acc = accum + int64_t(s) * C * 2;  // (sum dz, sum dz*xhat) per channel
    for (int g = threadIdx.x; g < G; g += blockDim.x) {
        double s1 = 0, s2 = 0;
        for (int j = 0; j < cpg; ++j) {
            const int c = g * cpg + j;
            s1 += double(gamma[c]) * acc[2 * c];
            s2 += double(gamma[c]) * acc[2 * c + 1];
        }
        g1[g] = float(s1);
        g2[g] = float(s2);
    }
    __syncthreads();
    for (int c = threadIdx.x; c < C; c += blockDim.x) {
        const int g = c / cpg;
        const float mean = stat[(int64_t(s) * G + g) * 2], rstd = stat[(int64_t(s) * G + g) * 2 + 1];
        const float invn = 1.0f / (float(P) * cpg);
        const float q = -rstd * rstd * g2[g] * invn;
        pc[c] = rstd * gamma[c];
        qc[c] = q;
        rc[c] = -rstd * g1[g] * invn - q * mean;
        sa[c] = ab[(int64_t(s) * C + c) * 2];
        sb[c] = ab[(int64_t(s) * C + c) * 2 + 1];
        if (blockIdx.x == 0) {
            if (dbeta) atomicAdd(dbeta + c, acc[2 * c]);
            if (dgamma) atomicAdd(dgamma + c, acc[2 * c + 1]);
        }
    }
    __syncthreads();
    const int V = C >> 3;
    const int64_t p0 = int64_t(blockIdx.x) * chunk_pixels, p1 = min(P, p0 + chunk_pixels);
    const int64_t nvec = (p1 - p0) * V;
    const int64_t base = (int64_t(s) * P + p0) * V;
    for (int64_t i = threadIdx.x; i < nvec; i += blockDim.x) {
        const int cv = int(i % V);
        float v[8], d[8], r[8];
        unpack8(__ldg(reinterpret_cast<const uint4*>(x) + base + i), v);
        unpack8(__ldg(reinterpret_cast<const uint4*>(dy) + base + i), d);
        if (add) unpack8(__ldg(reinterpret_cast<const uint4*>(add) + base + i), r);
#pragma unroll
        for (int j = 0; j < 8; ++j) {
            const int c = cv * 8 + j;
            float dz = d[j];
            if (silu) {
                const float z = sa[c] * v[j] + sb[c];
                const float sg = sigmoidf_(z);
                dz *= sg * (1.f + z * (1.f - sg));
            }
            float o = pc[c] * dz + qc[c] * v[j] + rc[c];
            if (add) o += r[j];
            v[j] = o;
        }
        reinterpret_cast<uint4*>(dx)[base + i] = pack8(v);
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
    // aim for ~4 blocks per SM overall; small problems get small chunks (>= 4 pixels) so they still spread over SMs
    const int64_t want = std::max<int64_t>(1, (4 * 148 + S - 1) / S);
    int64_t cp = std::max<int64_t>(4, (P + want - 1) / want);
    chunk_pixels = int(std::min<int64_t>(cp, P));
    chunks = int((P + chunk_pixels - 1) / chunk_pixels);
}

static void gn_set_attrs() {
    static bool done = false;
    if (done) return;
    cudaFuncSetAttribute(gn_bwd_apply_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 160 * 1024);
    cudaFuncSetAttribute(gn_fwd_apply_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 96 * 1024);
    cudaFuncSetAttribute(gn_partial_kernel<0>, cudaFuncAttributeMaxDynamicSharedMemorySize, 96 * 1024);
    cudaFuncSetAttribute(gn_partial_kernel<1>, cudaFuncAttributeMaxDynamicSharedMemorySize, 96 * 1024);
    done = true;
}

}  // namespace t2v

using namespace t2v;

extern "C" {

int64_t t2v_groupnorm_workspace_bytes(int32_t S, int64_t P, int32_t C) {
    (void)P;
    return int64_t(S) * C * 2 * sizeof(float) + 256;
}

int t2v_groupnorm_fwd(const void* x, const float* gamma, const float* beta, void* y, float* stat, float* ab, void* workspace,
                      int32_t S, int64_t P, int32_t C, int32_t G, float eps, int32_t silu, void* stream_) {
    if (C % 8 || C % G || C / 8 > 1024) return fail(-2, "groupnorm: C=%d G=%d unsupported", C, G);
    cudaStream_t st = static_cast<cudaStream_t>(stream_);
    gn_set_attrs();
    int cp, chunks;
    gn_chunks(S, P, cp, chunks);
    float* accum = static_cast<float*>(workspace);
    cudaMemsetAsync(accum, 0, size_t(S) * C * 2 * sizeof(float), st);
    const int bs = gn_block(C);
    gn_partial_kernel<0><<<dim3(chunks, S), bs, 2 * C * sizeof(float), st>>>(
        static_cast<const __nv_bfloat16*>(x), nullptr, nullptr, nullptr, accum, P, C, G, cp, 0);
    gn_fwd_apply_kernel<<<dim3(chunks, S), 256, (2 * C + 2 * G) * sizeof(float), st>>>(
        static_cast<const __nv_bfloat16*>(x), accum, gamma, beta, static_cast<__nv_bfloat16*>(y), stat, ab, P, C, G, cp, eps, silu);
    count_launch(1);
    return launch_checked(int(cudaGetLastError()), "groupnorm_fwd");
}

int t2v_groupnorm_bwd(const void* dy, const void* x, const float* gamma, const float* stat, const float* ab, const void* add,
                      void* dx, float* dgamma, float* dbeta, void* workspace, int32_t S, int64_t P, int32_t C, int32_t G,
                      int32_t silu, void* stream_) {
    if (C % 8 || C % G || C / 8 > 1024) return fail(-2, "groupnorm: C=%d G=%d unsupported", C, G);
    cudaStream_t st = static_cast<cudaStream_t>(stream_);
    gn_set_attrs();
    int cp, chunks;
    gn_chunks(S, P, cp, chunks);
    float* accum = static_cast<float*>(workspace);
    cudaMemsetAsync(accum, 0, size_t(S) * C * 2 * sizeof(float), st);
    const int bs = gn_block(C);
    gn_partial_kernel<1><<<dim3(chunks, S), bs, 2 * C * sizeof(float), st>>>(
        static_cast<const __nv_bfloat16*>(x), static_cast<const __nv_bfloat16*>(dy), ab, stat, accum, P, C, G, cp, silu);
    gn_bwd_apply_kernel<<<dim3(chunks, S), 256, (5 * C + 2 * G) * sizeof(float), st>>>(
        static_cast<const __nv_bfloat16*>(x), static_cast<const __nv_bfloat16*>(dy), accum, gamma, stat, ab,
        static_cast<const __nv_bfloat16*>(add), static_cast<__nv_bfloat16*>(dx), dgamma, dbeta, P, C, G, cp, silu);
    count_launch(1);
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
