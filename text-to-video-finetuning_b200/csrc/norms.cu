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
// ONE kernel per direction: every block is resident at once (grid <= SMs x occupancy),
// so the blocks of a sample can meet at a spin barrier between the two passes over their own pixels:
//   pass 1  per-channel sums over the block's chunk of pixels, reduced in shared memory, one red.global.add per channel
//           and block into accum[S][C][2]           fwd: (sum x, sum x^2)      bwd: (sum dz, sum dz*xhat)
//   barrier per-sample arrival counter (workspace), bounded spin
//   pass 2  every block finalises its sample's group statistics / coefficients from accum (fp64 group combine) and
//           streams its chunk again - the second read hits L2 (the first pass just pulled it in)
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
    float* accum;             // [S][C][2] workspace (zero on entry, zero again on exit)
    unsigned* arrive;         // [S] workspace: blocks of the sample that finished pass 1
    unsigned* done;           // [S] workspace: blocks of the sample that finished reading accum
    float* dgamma;
    float* dbeta;
    int64_t P;
    int C, G, chunk_pixels, chunks, silu;
    float eps;
};

__device__ __forceinline__ uint64_t globaltimer_ns() {
    uint64_t t;
    asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t));
    return t;
}

__device__ __forceinline__ void gn_sample_barrier(unsigned* counter, unsigned target) {
    __syncthreads();
    if (threadIdx.x == 0) {
        __threadfence();
        atomicAdd(counter, 1u);
        const uint64_t t0 = globaltimer_ns();
        while (true) {
            unsigned v;
            asm volatile("ld.acquire.gpu.global.u32 %0, [%1];" : "=r"(v) : "l"(counter) : "memory");
            if (v >= target) break;
            __nanosleep(64);
            if (globaltimer_ns() - t0 > 3000000000ull) __trap();  // a resident-grid assumption was violated
        }
        __threadfence();
    }
    __syncthreads();
}

template <int MODE>
__global__ void __launch_bounds__(kGnThreads, 1) gn_fused_kernel(const GnArgs g) {
    pdl_sync();
    extern __shared__ float sh[];  // [2][C] partial sums, then [2][G] group terms
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

    float a[8], b[8];  // z = a x + b (bwd: read back from the forward pass; fwd: computed after the barrier)
    float mean[8], rstd[8];
    if (MODE == 1 && active) {
#pragma unroll
        for (int j = 0; j < 8; ++j) {
            const int c = cv * 8 + j;
            a[j] = g.ab[(int64_t(s) * C + c) * 2];
            b[j] = g.ab[(int64_t(s) * C + c) * 2 + 1];
            mean[j] = g.stat[(int64_t(s) * G + c / cpg) * 2];
            rstd[j] = g.stat[(int64_t(s) * G + c / cpg) * 2 + 1];
        }
    }
    // ---- pass 1
    if (active) {
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
#pragma unroll
        for (int j = 0; j < 8; ++j) {
            atomicAdd(&sh[cv * 8 + j], acc0[j]);
            atomicAdd(&sh[C + cv * 8 + j], acc1[j]);
        }
    }
    __syncthreads();
    float* acc = g.accum + int64_t(s) * C * 2;
    for (int c = threadIdx.x; c < C; c += kGnThreads) {
        atomicAdd(acc + 2 * c, sh[c]);
        atomicAdd(acc + 2 * c + 1, sh[C + c]);
    }
    gn_sample_barrier(g.arrive + s, unsigned(g.chunks));

    // ---- finalise (every block, redundantly): one L2 read per channel, fp64 group combine
    float* t0 = sh;        // fwd: group mean   bwd: sum_c gamma * sum dz
    float* t1 = sh + G;    // fwd: group rstd   bwd: sum_c gamma * sum dz*xhat
    // one warp per group: lanes stride over the group's channels (one L2 read each), fp64 combine through shuffles
    __shared__ int is_last;
    {
        const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
        for (int gi = warp; gi < G; gi += kGnThreads / 32) {
            double a0 = 0, a1 = 0;
            for (int j = lane; j < cpg; j += 32) {
                const int c = gi * cpg + j;
                const float2 v = __ldcg(reinterpret_cast<const float2*>(acc) + c);
                const double w = MODE == 0 ? 1.0 : double(g.gamma[c]);
                a0 += w * v.x;
                a1 += w * v.y;
                if (MODE == 1 && chunk == 0) {
                    if (g.dbeta) atomicAdd(g.dbeta + c, v.x);
                    if (g.dgamma) atomicAdd(g.dgamma + c, v.y);
                }
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
    // every read of accum by this block is done: the last block of the sample to get here puts the workspace back to zero
    // (the contract of the workspace: zero on entry, zero on exit - no memset node per call)
    if (threadIdx.x == 0) {
        __threadfence();
        is_last = atomicAdd(g.done + s, 1u) == unsigned(g.chunks) - 1u;
    }
    __syncthreads();
    if (is_last) {
        for (int i = threadIdx.x; i < 2 * C; i += kGnThreads) acc[i] = 0.f;
        if (threadIdx.x == 0) {
            g.arrive[s] = 0u;
            g.done[s] = 0u;
        }
    }
    if (MODE == 0 && chunk == 0) {
        for (int c = threadIdx.x; c < C; c += kGnThreads) {
            const float aa = t1[c / cpg] * g.gamma[c];
            g.ab[(int64_t(s) * C + c) * 2] = aa;
            g.ab[(int64_t(s) * C + c) * 2 + 1] = g.beta[c] - t0[c / cpg] * aa;
        }
    }
    if (!active) return;
    // ---- pass 2
    uint4* os = reinterpret_cast<uint4*>(g.out + int64_t(s) * g.P * C) + cv;
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
            const float q = -rstd[j] * rstd[j] * t1[c / cpg] * invn;
            pc[j] = rstd[j] * g.gamma[c];
            qc[j] = q;
            rc[j] = -rstd[j] * t0[c / cpg] * invn - q * mean[j];
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

// Blocks that can be resident at once (the spin barrier needs the whole grid on the chip).
// Why the barrier cannot deadlock although other kernels may share the SMs: the grid never exceeds what fits on the chip by
// itself; kernels launched BEFORE this one on any stream terminate without waiting for it, so their SM resources free up and
// the remaining blocks become resident; kernels launched AFTER it on the same stream (programmatic dependent launch) can
// only start once every block of this grid has executed griddepcontrol.launch_dependents, i.e. is already resident; and a
// concurrent kernel on the side stream (weight gradients) never waits on this one either.  The spin is bounded (3 s, then
// trap) so that a violated assumption fails loudly instead of hanging the GPU.
static int gn_resident_blocks() {
    static int cached = 0;
    if (cached) return cached;
    cudaFuncSetAttribute(gn_fused_kernel<0>, cudaFuncAttributeMaxDynamicSharedMemorySize, 64 * 1024);
    cudaFuncSetAttribute(gn_fused_kernel<1>, cudaFuncAttributeMaxDynamicSharedMemorySize, 64 * 1024);
    int per_sm0 = 0, per_sm1 = 0;
    cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm0, gn_fused_kernel<0>, kGnThreads, 64 * 1024);
    cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm1, gn_fused_kernel<1>, kGnThreads, 64 * 1024);
    cached = std::max(1, std::min(per_sm0, per_sm1)) * device_sm_count();
    return cached;
}

static int gn_plan(int S, int64_t P, int C, int& chunk_pixels, int& chunks) {
    const int resident = gn_resident_blocks();
    if (S > resident) return -1;
    const int lanes = kGnThreads / (C / 8);
    // one block per SM overall, each with at least one pixel per lane (and >= 2 pixels) so tiny maps still spread out
    const int64_t want = std::max<int64_t>(1, resident / S);
    const int64_t min_px = std::max<int64_t>(2, lanes);
    const int64_t cp = std::max<int64_t>(min_px, (P + want - 1) / want);
    chunk_pixels = int(std::min<int64_t>(cp, P));
    chunks = int((P + chunk_pixels - 1) / chunk_pixels);
    return 0;
}

static size_t gn_accum_bytes(int S, int C) { return (size_t(S) * C * 2 * sizeof(float) + 255) / 256 * 256; }

template <int MODE>
static int gn_launch(GnArgs& g, void* workspace, int S, cudaStream_t st) {
    if (g.C % 8 || g.C % g.G || g.C / 8 > kGnThreads || 2 * g.C * sizeof(float) > 64 * 1024)
        return fail(-2, "groupnorm: C=%d G=%d unsupported", g.C, g.G);
    if (gn_plan(S, g.P, g.C, g.chunk_pixels, g.chunks)) return fail(-2, "groupnorm: %d samples exceed the resident grid", S);
    const size_t ab = gn_accum_bytes(S, g.C);
    g.accum = static_cast<float*>(workspace);
    g.arrive = reinterpret_cast<unsigned*>(static_cast<char*>(workspace) + ab);
    g.done = g.arrive + S;
    launch_pdl(gn_fused_kernel<MODE>, dim3(S * g.chunks), dim3(kGnThreads), std::max<size_t>(2 * g.C, 2 * g.G) * sizeof(float), st, g);
    count_launch(1);
    return 0;
}

}  // namespace t2v

using namespace t2v;

extern "C" {

int64_t t2v_groupnorm_workspace_bytes(int32_t S, int64_t P, int32_t C) {
    (void)P;
    return int64_t(gn_accum_bytes(S, C)) + 2 * int64_t(S) * sizeof(unsigned) + 256;
}

int t2v_groupnorm_fwd(const void* x, const float* gamma, const float* beta, void* y, float* stat, float* ab, void* workspace,
                      int32_t S, int64_t P, int32_t C, int32_t G, float eps, int32_t silu, void* stream_) {
    GnArgs g{};
    g.x = static_cast<const __nv_bfloat16*>(x);
    g.out = static_cast<__nv_bfloat16*>(y);
    g.gamma = gamma; g.beta = beta; g.stat = stat; g.ab = ab;
    g.P = P; g.C = C; g.G = G; g.silu = silu; g.eps = eps;
    if (int r = gn_launch<0>(g, workspace, S, static_cast<cudaStream_t>(stream_))) return r;
    return launch_checked(int(cudaGetLastError()), "groupnorm_fwd");
}

int t2v_groupnorm_bwd(const void* dy, const void* x, const float* gamma, const float* stat, const float* ab, const void* add,
                      void* dx, float* dgamma, float* dbeta, void* workspace, int32_t S, int64_t P, int32_t C, int32_t G,
                      int32_t silu, void* stream_) {
    GnArgs g{};
    g.x = static_cast<const __nv_bfloat16*>(x);
    g.dy = static_cast<const __nv_bfloat16*>(dy);
    g.add = static_cast<const __nv_bfloat16*>(add);
    g.out = static_cast<__nv_bfloat16*>(dx);
    g.gamma = gamma; g.stat = const_cast<float*>(stat); g.ab = const_cast<float*>(ab);
    g.dgamma = dgamma; g.dbeta = dbeta;
    g.P = P; g.C = C; g.G = G; g.silu = silu;
    if (int r = gn_launch<1>(g, workspace, S, static_cast<cudaStream_t>(stream_))) return r;
    return launch_checked(int(cudaGetLastError()), "groupnorm_bwd");
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
