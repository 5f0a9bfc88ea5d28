// GroupNorm (+SiLU) and LayerNorm, forward and backward, for channels-last bf16 activations.
// HBM-bound kernels: 128-bit loads, fp32 statistics, warp-shuffle / shared-memory reductions.
//
// GroupNorm works on x [S][P][C]: S normalisation samples (frames for the per-frame norms of ResnetBlock2D /
// Transformer2DModel, clips for the per-clip norms of TemporalConvLayer / TransformerTemporalModel), P pixels per
// sample, C channels in G groups.  Statistics are reduced in two levels (pixel chunks -> sample) so the grid fills
// the GPU even when S == 1.
#include "common.h"
#include "gemm_tc.cuh"
#include "ptx.cuh"

#include <algorithm>
#include <cuda_bf16.h>
#include <mutex>

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
// Two kernels per direction, chained with programmatic dependent launch:
//   sums     per-channel sums over a block's chunk of pixels, reduced in shared memory, one red.global.add.v2 per channel
//            and block into accum[S][C][2]          fwd: (sum x, sum x^2)      bwd: (sum dz, sum dz*xhat)
//            The forward sums are usually NOT run: the GEMM that produced x has already accumulated them per frame and
//            channel in its epilogue (gemm_tc.cu, EPI_STATS), so GroupNorm forward is one read + one write of the tensor.
//   apply    every block finalises its sample's group statistics / coefficients from the sums (fp64 group combine, one
//            L2 round trip) and streams its chunk of pixels once.
// Thread layout: V = C/8 channel vectors; thread owns vector tid % V (coefficients live in registers) and pixel lane
// tid / V; loads are 16 bytes, 4 pixels in flight per thread.
constexpr int kGnThreads = 512;

struct GnArgs {
    const __nv_bfloat16* x;
    const __nv_bfloat16* dy;
    const __nv_bfloat16* add;
    __nv_bfloat16* out;       // y (fwd) / dx (bwd)
    const float* gamma;
    const float* beta;
    float* stat;              // [S][G][2] (mean, rstd): written by fwd, read by bwd
    float* ab;                // [S][C][2] (a, b) with z = a x + b: written by fwd, read by bwd
    float* accum;             // sums kernel output / bwd apply input: [S][C][2]
    const float* stats0;      // fwd apply input: per-frame sums of channels [0, C0), row pitch ld0 channels
    const float* stats1;      // ... of channels [C0, C), row pitch ld1 (NULL when C0 == C)
    int64_t ld0, ld1;
    float* dgamma;
    float* dbeta;
    int64_t P;
    int C, C0, G, fps, chunk_pixels, chunks, silu;
    float eps;
};

// ---- sums: MODE 0 forward (x), MODE 1 backward (dy, x, saved coefficients)
template <int MODE>
__global__ void __launch_bounds__(kGnThreads, 1) gn_sums_kernel(const GnArgs g) {
    pdl_sync();
    extern __shared__ float sh[];  // [2][C] partial sums
    const int C = g.C, G = g.G, cpg = C / G;
    const int s = blockIdx.x / g.chunks, chunk = blockIdx.x % g.chunks;
    const int V = C >> 3;
    const int lanes = kGnThreads / V;
    const int cv = threadIdx.x % V, pl = threadIdx.x / V;
    const bool active = pl < lanes;
    const int64_t p0 = int64_t(chunk) * g.chunk_pixels;
    const int64_t p1 = min(g.P, p0 + g.chunk_pixels);
    const uint4* xs = reinterpret_cast<const uint4*>(g.x + int64_t(s) * g.P * C) + cv;
    const uint4* ds = MODE == 1 ? reinterpret_cast<const uint4*>(g.dy + int64_t(s) * g.P * C) + cv : nullptr;
    for (int i = threadIdx.x; i < 2 * C; i += kGnThreads) sh[i] = 0.f;
    __syncthreads();
    if (active) {
        float a[8], b[8], mean[8], rstd[8];
        if (MODE == 1) {
#pragma unroll
            for (int j = 0; j < 8; ++j) {
                const int c = cv * 8 + j;
                a[j] = g.ab[(int64_t(s) * C + c) * 2];
                b[j] = g.ab[(int64_t(s) * C + c) * 2 + 1];
                mean[j] = g.stat[(int64_t(s) * G + c / cpg) * 2];
                rstd[j] = g.stat[(int64_t(s) * G + c / cpg) * 2 + 1];
            }
        }
        float acc0[8], acc1[8];
#pragma unroll
        for (int j = 0; j < 8; ++j) acc0[j] = acc1[j] = 0.f;
        auto accumulate = [&](const uint4& qx, const uint4& qd) {
            float v[8];
            unpack8(qx, v);
            if (MODE == 0) {
#pragma unroll
                for (int j = 0; j < 8; ++j) {
                    acc0[j] += v[j];
                    acc1[j] += v[j] * v[j];
                }
            } else {
                float d[8];
                unpack8(qd, d);
#pragma unroll
                for (int j = 0; j < 8; ++j) {
                    float dz = d[j];
                    if (g.silu) {
                        const float z = a[j] * v[j] + b[j];
                        const float sg = sigmoidf_(z);
                        dz *= sg * (1.f + z * (1.f - sg));
                    }
                    acc0[j] += dz;
                    acc1[j] += dz * (v[j] - mean[j]) * rstd[j];
                }
            }
        };
        int64_t p = p0 + pl;
        for (; p + 3 * lanes < p1; p += 4 * lanes) {
            uint4 qx[4], qd[4];
#pragma unroll
            for (int u = 0; u < 4; ++u) {
                qx[u] = __ldg(xs + (p + u * lanes) * V);
                if (MODE == 1) qd[u] = __ldg(ds + (p + u * lanes) * V);
            }
#pragma unroll
            for (int u = 0; u < 4; ++u) accumulate(qx[u], qd[u]);
        }
        for (; p < p1; p += lanes) {
            uint4 qd = make_uint4(0, 0, 0, 0);
            if (MODE == 1) qd = __ldg(ds + p * V);
            accumulate(__ldg(xs + p * V), qd);
        }
        if (lanes == 1) {   // one pixel lane per channel vector: no contention, plain stores
#pragma unroll
            for (int j = 0; j < 8; ++j) {
                sh[cv * 8 + j] = acc0[j];
                sh[C + cv * 8 + j] = acc1[j];
            }
        } else {
#pragma unroll
            for (int j = 0; j < 8; ++j) {
                atomicAdd(&sh[cv * 8 + j], acc0[j]);
                atomicAdd(&sh[C + cv * 8 + j], acc1[j]);
            }
        }
    }
    __syncthreads();
    float* acc = g.accum + int64_t(s) * C * 2;
    for (int c = threadIdx.x; c < C; c += kGnThreads) red_add_f32x2(acc + 2 * c, sh[c], sh[C + c]);
}

// ---- apply: MODE 0 forward, MODE 1 backward
template <int MODE>
__global__ void __launch_bounds__(kGnThreads, 1) gn_apply_kernel(const GnArgs g) {
    pdl_sync();
    extern __shared__ float sh[];
    const int C = g.C, G = g.G, cpg = C / G;
    const int s = blockIdx.x / g.chunks, chunk = blockIdx.x % g.chunks;
    const int V = C >> 3;
    const int lanes = kGnThreads / V;
    const int cv = threadIdx.x % V, pl = threadIdx.x / V;
    const bool active = pl < lanes;
    const int64_t p0 = int64_t(chunk) * g.chunk_pixels;
    const int64_t p1 = min(g.P, p0 + g.chunk_pixels);
    const uint4* xs = reinterpret_cast<const uint4*>(g.x + int64_t(s) * g.P * C) + cv;
    const uint4* ds = MODE == 1 ? reinterpret_cast<const uint4*>(g.dy + int64_t(s) * g.P * C) + cv : nullptr;
    // ---- per-channel sums of this sample -> shared memory (fwd: summed over the sample's frames)
    float* cs = sh;              // [2][C]
    float* t0 = sh + 2 * C;      // [G]  fwd: group mean   bwd: sum_c gamma * sum dz
    float* t1 = t0 + G;          // [G]  fwd: group rstd   bwd: sum_c gamma * sum dz*xhat
    if (MODE == 0) {
        for (int c = threadIdx.x; c < C; c += kGnThreads) {
            const bool first = c < g.C0;
            const float2* src = reinterpret_cast<const float2*>(first ? g.stats0 : g.stats1);
            const int64_t ld = first ? g.ld0 : g.ld1;
            const int cc = first ? c : c - g.C0;
            float a0 = 0.f, a1 = 0.f;
            for (int f = 0; f < g.fps; ++f) {
                const float2 v = __ldcg(src + (int64_t(s) * g.fps + f) * ld + cc);
                a0 += v.x;
                a1 += v.y;
            }
            cs[c] = a0;
            cs[C + c] = a1;
        }
    } else {
        const float2* acc = reinterpret_cast<const float2*>(g.accum + int64_t(s) * C * 2);
        for (int c = threadIdx.x; c < C; c += kGnThreads) {
            const float2 v = __ldcg(acc + c);
            cs[c] = v.x;
            cs[C + c] = v.y;
            if (chunk == 0) {
                if (g.dbeta) atomicAdd(g.dbeta + c, v.x);
                if (g.dgamma) atomicAdd(g.dgamma + c, v.y);
            }
        }
    }
    __syncthreads();
    {   // one warp per group: fp64 combine of the group's channels through shuffles
        const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
        for (int gi = warp; gi < G; gi += kGnThreads / 32) {
            double a0 = 0, a1 = 0;
            for (int j = lane; j < cpg; j += 32) {
                const int c = gi * cpg + j;
                const double w = MODE == 0 ? 1.0 : double(g.gamma[c]);
                a0 += w * cs[c];
                a1 += w * cs[C + c];
            }
#pragma unroll
            for (int o = 16; o > 0; o >>= 1) {
                a0 += __shfl_xor_sync(0xffffffffu, a0, o);
                a1 += __shfl_xor_sync(0xffffffffu, a1, o);
            }
            if (lane == 0) {
                if (MODE == 0) {
                    const double n = double(g.P) * cpg;
                    const double m = a0 / n;
                    double var = a1 / n - m * m;
                    if (var < 0) var = 0;
                    const float r = float(1.0 / sqrt(var + double(g.eps)));
                    t0[gi] = float(m);
                    t1[gi] = r;
                    if (chunk == 0) {
                        g.stat[(int64_t(s) * G + gi) * 2] = float(m);
                        g.stat[(int64_t(s) * G + gi) * 2 + 1] = r;
                    }
                } else {
                    t0[gi] = float(a0);
                    t1[gi] = float(a1);
                }
            }
        }
    }
    __syncthreads();
    if (MODE == 0 && chunk == 0) {
        for (int c = threadIdx.x; c < C; c += kGnThreads) {
            const float aa = t1[c / cpg] * g.gamma[c];
            g.ab[(int64_t(s) * C + c) * 2] = aa;
            g.ab[(int64_t(s) * C + c) * 2 + 1] = g.beta[c] - t0[c / cpg] * aa;
        }
    }
    if (!active) return;
    uint4* os = reinterpret_cast<uint4*>(g.out + int64_t(s) * g.P * C) + cv;
    float a[8], b[8];
    if (MODE == 0) {
#pragma unroll
        for (int j = 0; j < 8; ++j) {
            const int c = cv * 8 + j;
            a[j] = t1[c / cpg] * g.gamma[c];
            b[j] = g.beta[c] - t0[c / cpg] * a[j];
        }
        auto apply = [&](const uint4& qx) {
            float v[8];
            unpack8(qx, v);
#pragma unroll
            for (int j = 0; j < 8; ++j) {
                float z = a[j] * v[j] + b[j];
                if (g.silu) z *= sigmoidf_(z);
                v[j] = z;
            }
            return pack8(v);
        };
        int64_t p = p0 + pl;
        for (; p + 3 * lanes < p1; p += 4 * lanes) {
            uint4 qx[4];
#pragma unroll
            for (int u = 0; u < 4; ++u) qx[u] = __ldg(xs + (p + u * lanes) * V);
#pragma unroll
            for (int u = 0; u < 4; ++u) os[(p + u * lanes) * V] = apply(qx[u]);
        }
        for (; p < p1; p += lanes) os[p * V] = apply(__ldg(xs + p * V));
    } else {
        // dx = pc * dz + qc * x + rc (+ add)
        float pc[8], qc[8], rc[8];
        const float invn = 1.0f / (float(g.P) * cpg);
#pragma unroll
        for (int j = 0; j < 8; ++j) {
            const int c = cv * 8 + j;
            a[j] = g.ab[(int64_t(s) * C + c) * 2];
            b[j] = g.ab[(int64_t(s) * C + c) * 2 + 1];
            const float mean = g.stat[(int64_t(s) * G + c / cpg) * 2];
            const float rstd = g.stat[(int64_t(s) * G + c / cpg) * 2 + 1];
            const float q = -rstd * rstd * t1[c / cpg] * invn;
            pc[j] = rstd * g.gamma[c];
            qc[j] = q;
            rc[j] = -rstd * t0[c / cpg] * invn - q * mean;
        }
        const uint4* as = g.add ? reinterpret_cast<const uint4*>(g.add + int64_t(s) * g.P * C) + cv : nullptr;
        auto apply = [&](const uint4& qx, const uint4& qd, const uint4& qa) {
            float v[8], d[8], r[8];
            unpack8(qx, v);
            unpack8(qd, d);
            unpack8(qa, r);
#pragma unroll
            for (int j = 0; j < 8; ++j) {
                float dz = d[j];
                if (g.silu) {
                    const float z = a[j] * v[j] + b[j];
                    const float sg = sigmoidf_(z);
                    dz *= sg * (1.f + z * (1.f - sg));
                }
                v[j] = pc[j] * dz + qc[j] * v[j] + rc[j] + r[j];
            }
            return pack8(v);
        };
        const uint4 zero = make_uint4(0, 0, 0, 0);
        int64_t p = p0 + pl;
        for (; p + lanes < p1; p += 2 * lanes) {
            uint4 qx[2], qd[2], qa[2];
#pragma unroll
            for (int u = 0; u < 2; ++u) {
                qx[u] = __ldg(xs + (p + u * lanes) * V);
                qd[u] = __ldg(ds + (p + u * lanes) * V);
                qa[u] = as ? __ldg(as + (p + u * lanes) * V) : zero;
            }
#pragma unroll
            for (int u = 0; u < 2; ++u) os[(p + u * lanes) * V] = apply(qx[u], qd[u], qa[u]);
        }
        for (; p < p1; p += lanes) os[p * V] = apply(__ldg(xs + p * V), __ldg(ds + p * V), as ? __ldg(as + p * V) : zero);
    }
}

// ------------------------------------------------------------------------------------------------ LayerNorm
// One warp per row; the row lives in registers (VPL 16-byte vectors per lane), exact two-pass statistics.
template <int VPL>
__global__ void ln_fwd_kernel(const __nv_bfloat16* __restrict__ x, const float* __restrict__ gamma,
                              const float* __restrict__ beta, __nv_bfloat16* __restrict__ y, float* __restrict__ stat,
                              int64_t rows, int C, float eps) {
    pdl_sync();
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
    pdl_sync();
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

static size_t gn_smem(int C, int G) { return size_t(2 * C + 2 * G) * sizeof(float); }

static void gn_set_attrs() {
    static std::once_flag once;
    std::call_once(once, [] {
        cudaFuncSetAttribute(gn_sums_kernel<0>, cudaFuncAttributeMaxDynamicSharedMemorySize, 64 * 1024);
        cudaFuncSetAttribute(gn_sums_kernel<1>, cudaFuncAttributeMaxDynamicSharedMemorySize, 64 * 1024);
        cudaFuncSetAttribute(gn_apply_kernel<0>, cudaFuncAttributeMaxDynamicSharedMemorySize, 64 * 1024);
        cudaFuncSetAttribute(gn_apply_kernel<1>, cudaFuncAttributeMaxDynamicSharedMemorySize, 64 * 1024);
    });
}

// Pixel chunks per sample: about two blocks per SM overall, at least one pixel per lane (and >= 2 pixels) per block.
static void gn_plan(int S, int64_t P, int C, int& chunk_pixels, int& chunks) {
    const int lanes = kGnThreads / (C / 8);
    const int64_t want = std::max<int64_t>(1, (2 * device_sm_count() + S - 1) / S);
    const int64_t min_px = std::max<int64_t>(2, lanes);
    const int64_t cp = std::max<int64_t>(min_px, (P + want - 1) / want);
    chunk_pixels = int(std::min<int64_t>(cp, P));
    chunks = int((P + chunk_pixels - 1) / chunk_pixels);
}

static int gn_check(const GnArgs& g) {
    if (g.C % 8 || g.C % g.G || g.C / 8 > kGnThreads || gn_smem(g.C, g.G) > 64 * 1024)
        return fail(-2, "groupnorm: C=%d G=%d unsupported", g.C, g.G);
    return 0;
}

// Standalone per-sample channel sums of x [S][P][C] into stats (+=), row pitch ld channels.  Used by t2v_channel_stats and
// by conv_fwd when a problem's tiling cannot produce the statistics in the GEMM epilogue.
int launch_channel_stats(const void* x, float* stats, int S, int64_t P, int C, int64_t ld, cudaStream_t st) {
    if (ld != C) return fail(-2, "channel_stats: row pitch %lld != C=%d is not supported by the standalone pass", (long long)ld, C);
    GnArgs g{};
    g.x = static_cast<const __nv_bfloat16*>(x);
    g.accum = stats;
    g.P = P; g.C = C; g.G = 1;
    if (C % 8 || C / 8 > kGnThreads || gn_smem(C, 1) > 64 * 1024) return fail(-2, "channel_stats: C=%d unsupported", C);
    gn_set_attrs();
    gn_plan(S, P, C, g.chunk_pixels, g.chunks);
    return int(launch_pdl(gn_sums_kernel<0>, dim3(S * g.chunks), dim3(kGnThreads), gn_smem(C, 1), st, g));
}

}  // namespace t2v

using namespace t2v;

extern "C" {

int64_t t2v_groupnorm_workspace_bytes(int32_t S, int64_t P, int32_t C) {
    (void)P;
    return int64_t(S) * C * 2 * sizeof(float);
}

int t2v_channel_stats(const void* x, float* stats, int32_t S, int64_t P, int32_t C, int64_t ld, void* stream_) {
    return launch_checked(launch_channel_stats(x, stats, S, P, C, ld, static_cast<cudaStream_t>(stream_)), "channel_stats");
}

int t2v_groupnorm_fwd(const void* x, const float* gamma, const float* beta, void* y, float* stat, float* ab, const float* stats0,
                      int32_t C0, int64_t ld0, const float* stats1, int64_t ld1, int32_t fps, void* workspace, int32_t S, int64_t P,
                      int32_t C, int32_t G, float eps, int32_t silu, void* stream_) {
    cudaStream_t st = static_cast<cudaStream_t>(stream_);
    GnArgs g{};
    g.x = static_cast<const __nv_bfloat16*>(x);
    g.out = static_cast<__nv_bfloat16*>(y);
    g.gamma = gamma; g.beta = beta; g.stat = stat; g.ab = ab;
    g.P = P; g.C = C; g.G = G; g.silu = silu; g.eps = eps;
    if (int r = gn_check(g)) return r;
    gn_set_attrs();
    gn_plan(S, P, C, g.chunk_pixels, g.chunks);
    if (stats0) {
        if (fps < 1 || C0 <= 0 || C0 > C || (C0 < C && !stats1)) return fail(-2, "groupnorm_fwd: bad statistics arguments");
        g.stats0 = stats0; g.stats1 = stats1; g.C0 = C0; g.ld0 = ld0; g.ld1 = ld1; g.fps = fps;
    } else {
        if (!workspace) return fail(-3, "groupnorm_fwd: needs producer statistics or a zeroed workspace");
        g.accum = static_cast<float*>(workspace);
        if (int rc = int(launch_pdl(gn_sums_kernel<0>, dim3(S * g.chunks), dim3(kGnThreads), gn_smem(C, G), st, g)))
            return launch_checked(rc, "groupnorm_fwd(sums)");
        count_launch(1);
        g.stats0 = g.accum; g.stats1 = nullptr; g.C0 = C; g.ld0 = C; g.ld1 = 0; g.fps = 1;
    }
    return launch_checked(int(launch_pdl(gn_apply_kernel<0>, dim3(S * g.chunks), dim3(kGnThreads), gn_smem(C, G), st, g)), "groupnorm_fwd");
}

int t2v_groupnorm_bwd(const void* dy, const void* x, const float* gamma, const float* stat, const float* ab, const void* add,
                      void* dx, float* dgamma, float* dbeta, void* workspace, int32_t S, int64_t P, int32_t C, int32_t G,
                      int32_t silu, void* stream_) {
    cudaStream_t st = static_cast<cudaStream_t>(stream_);
    GnArgs g{};
    g.x = static_cast<const __nv_bfloat16*>(x);
    g.dy = static_cast<const __nv_bfloat16*>(dy);
    g.add = static_cast<const __nv_bfloat16*>(add);
    g.out = static_cast<__nv_bfloat16*>(dx);
    g.gamma = gamma; g.stat = const_cast<float*>(stat); g.ab = const_cast<float*>(ab);
    g.dgamma = dgamma; g.dbeta = dbeta;
    g.P = P; g.C = C; g.G = G; g.silu = silu;
    if (int r = gn_check(g)) return r;
    if (!workspace) return fail(-3, "groupnorm_bwd: needs a zeroed workspace");
    gn_set_attrs();
    gn_plan(S, P, C, g.chunk_pixels, g.chunks);
    g.accum = static_cast<float*>(workspace);
    if (int rc = int(launch_pdl(gn_sums_kernel<1>, dim3(S * g.chunks), dim3(kGnThreads), gn_smem(C, G), st, g)))
        return launch_checked(rc, "groupnorm_bwd(sums)");
    count_launch(1);
    return launch_checked(int(launch_pdl(gn_apply_kernel<1>, dim3(S * g.chunks), dim3(kGnThreads), gn_smem(C, G), st, g)), "groupnorm_bwd");
}

#define LN_DISPATCH(KERNEL, GRID, SMEM, ST, ...)                                                       \
    switch (vpl) {                                                                                         \
        case 1: launch_pdl(KERNEL<1>, dim3(GRID), dim3(256), size_t(SMEM), ST, __VA_ARGS__); break;       \
        case 2: launch_pdl(KERNEL<2>, dim3(GRID), dim3(256), size_t(SMEM), ST, __VA_ARGS__); break;       \
        case 3: launch_pdl(KERNEL<3>, dim3(GRID), dim3(256), size_t(SMEM), ST, __VA_ARGS__); break;       \
        case 4: launch_pdl(KERNEL<4>, dim3(GRID), dim3(256), size_t(SMEM), ST, __VA_ARGS__); break;       \
        case 5: launch_pdl(KERNEL<5>, dim3(GRID), dim3(256), size_t(SMEM), ST, __VA_ARGS__); break;       \
        case 6: launch_pdl(KERNEL<6>, dim3(GRID), dim3(256), size_t(SMEM), ST, __VA_ARGS__); break;       \
        case 7: launch_pdl(KERNEL<7>, dim3(GRID), dim3(256), size_t(SMEM), ST, __VA_ARGS__); break;       \
        default: launch_pdl(KERNEL<8>, dim3(GRID), dim3(256), size_t(SMEM), ST, __VA_ARGS__); break;      \
    }

int t2v_layernorm_fwd(const void* x, const float* gamma, const float* beta, void* y, float* stat, int64_t rows, int32_t C,
                      float eps, void* stream_) {
    if (C % 8 || C > 2048) return fail(-2, "layernorm: C=%d unsupported (multiple of 8, <= 2048)", C);
    cudaStream_t st = static_cast<cudaStream_t>(stream_);
    const int vpl = (C / 8 + 31) / 32;
    const int grid = int(std::min<int64_t>((rows + 7) / 8, 148 * 8));
    LN_DISPATCH(ln_fwd_kernel, grid, 0, st, static_cast<const __nv_bfloat16*>(x), gamma, beta, static_cast<__nv_bfloat16*>(y), stat,
                rows, C, eps);
    return launch_checked(int(cudaGetLastError()), "layernorm_fwd");
}

int t2v_layernorm_bwd(const void* dy, const void* x, const float* gamma, const float* stat, const void* add, void* dx,
                      float* dgamma, float* dbeta, int64_t rows, int32_t C, void* stream_) {
    if (C % 8 || C > 2048) return fail(-2, "layernorm: C=%d unsupported (multiple of 8, <= 2048)", C);
    cudaStream_t st = static_cast<cudaStream_t>(stream_);
    const int vpl = (C / 8 + 31) / 32;
    const int grid = int(std::min<int64_t>((rows + 7) / 8, 148 * 2));
    LN_DISPATCH(ln_bwd_kernel, grid, 2 * C * sizeof(float), st, static_cast<const __nv_bfloat16*>(x),
                static_cast<const __nv_bfloat16*>(dy), gamma, stat, static_cast<const __nv_bfloat16*>(add),
                static_cast<__nv_bfloat16*>(dx), dgamma, dbeta, rows, C);
    return launch_checked(int(cudaGetLastError()), "layernorm_bwd");
}

}  // extern "C"
