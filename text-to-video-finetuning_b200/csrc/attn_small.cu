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

template <int D>
__device__ __forceinline__ void load_tile(const __nv_bfloat16* __restrict__ g, int64_t base, int64_t stride, int L, float* sm,
                                          int lane) {
    // 16-byte loads: D/8 chunks per token, independent iterations (all loads of a tile are in flight together)
    constexpr int CPT = D / 8;
    const int total = L * CPT;
#pragma unroll 4
    for (int idx = lane; idx < total; idx += 32) {
        const int t = idx / CPT, c = idx % CPT;
        const uint4 q = __ldg(reinterpret_cast<const uint4*>(g + base + t * stride) + c);
        float* dst = sm + t * (D + 1) + c * 8;
        dst[0] = bf16_lo(q.x); dst[1] = bf16_hi(q.x); dst[2] = bf16_lo(q.y); dst[3] = bf16_hi(q.y);
        dst[4] = bf16_lo(q.z); dst[5] = bf16_hi(q.z); dst[6] = bf16_lo(q.w); dst[7] = bf16_hi(q.w);
    }
}

template <int D>
__global__ void attn_small_fwd_kernel(const __nv_bfloat16* __restrict__ q, const __nv_bfloat16* __restrict__ k,
                                      const __nv_bfloat16* __restrict__ v, __nv_bfloat16* __restrict__ o, SeqAddr a, int64_t nseq,
                                      int heads, int L, float scale) {
    pdl_sync();
    extern __shared__ float sm_all[];
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int tile = L * (D + 1);  // shared memory is sized by the actual sequence length (occupancy)
    float* sq = sm_all + warp * 3 * tile;
    float* sk = sq + tile;
    float* sv = sk + tile;
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
            if (D == 64) {  // lane owns dims (2*lane, 2*lane+1): one coalesced 128-byte row store per query
                float a0 = 0.f, a1 = 0.f;
                for (int j = 0; j < L; ++j) {
                    const float pj = __shfl_sync(0xffffffffu, p, j);
                    a0 += pj * sv[j * (D + 1) + 2 * lane];
                    a1 += pj * sv[j * (D + 1) + 2 * lane + 1];
                }
                reinterpret_cast<__nv_bfloat162*>(o + obase + i * ostr)[lane] = __floats2bfloat162_rn(a0, a1);
            } else {
                float a0 = 0.f;
                for (int j = 0; j < L; ++j) a0 += __shfl_sync(0xffffffffu, p, j) * sv[j * (D + 1) + lane];
                o[obase + i * ostr + lane] = __float2bfloat16_rn(a0);
            }
        }
    }
}

// Backward.  Phase 1 (lane j <-> key j): recompute P and dS row by row into shared memory.  Phase 2 (lane <-> head
// dims lane, lane+32): dQ = dS K, dK = dS^T Q, dV = P^T dO as L x L register-light loops over broadcast scalars.
template <int D>
__global__ void attn_small_bwd_kernel(const __nv_bfloat16* __restrict__ q, const __nv_bfloat16* __restrict__ k,
                                      const __nv_bfloat16* __restrict__ v, const __nv_bfloat16* __restrict__ dout,
                                      __nv_bfloat16* __restrict__ dq, __nv_bfloat16* __restrict__ dk, __nv_bfloat16* __restrict__ dv,
                                      SeqAddr a, int64_t nseq, int heads, int L, float scale) {
    pdl_sync();
    extern __shared__ float sm_all[];
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int tile = L * (D + 1);
    float* sq = sm_all + warp * (4 * tile + 2 * L * L);
    float* sk = sq + tile;
    float* sv = sk + tile;
    float* sd = sv + tile;
    float* sp = sd + tile;      // P  [L][L]
    float* ss = sp + L * L;     // dS [L][L] (already multiplied by the softmax scale)
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
        load_tile<D>(dout, obase, ostr, L, sd, lane);
        __syncwarp();
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
            if (lane < L) {
                sp[i * L + lane] = p;
                ss[i * L + lane] = p * (dp - dot) * scale;
            }
        }
        __syncwarp();
        // Phase 2: lane owns NP adjacent head dims of every row.  Rows are cached in registers in slabs of 8 keys so the
        // inner loops are broadcast-scalar x register FMAs (no bank-conflicted shared-memory reads).
        constexpr int NP = D == 64 ? 2 : 1;
        constexpr int SLAB = 8;
        for (int i0 = 0; i0 < L; i0 += SLAB) {
            float aq[SLAB][NP], ak[SLAB][NP], av[SLAB][NP];
#pragma unroll
            for (int ii = 0; ii < SLAB; ++ii)
#pragma unroll
                for (int r = 0; r < NP; ++r) aq[ii][r] = ak[ii][r] = av[ii][r] = 0.f;
            for (int j = 0; j < L; ++j) {
                float kj[NP], qj[NP], dj[NP];
#pragma unroll
                for (int r = 0; r < NP; ++r) {
                    const int d = NP * lane + r;
                    kj[r] = sk[j * (D + 1) + d];
                    qj[r] = sq[j * (D + 1) + d];
                    dj[r] = sd[j * (D + 1) + d];
                }
#pragma unroll
                for (int ii = 0; ii < SLAB; ++ii) {
                    const int i = i0 + ii;
                    if (i < L) {
                        const float ds_ij = ss[i * L + j], ds_ji = ss[j * L + i], p_ji = sp[j * L + i];
#pragma unroll
                        for (int r = 0; r < NP; ++r) {
                            aq[ii][r] += ds_ij * kj[r];   // dQ_i = sum_j dS_ij K_j
                            ak[ii][r] += ds_ji * qj[r];   // dK_i = sum_j dS_ji Q_j
                            av[ii][r] += p_ji * dj[r];    // dV_i = sum_j P_ji dO_j
                        }
                    }
                }
            }
#pragma unroll
            for (int ii = 0; ii < SLAB; ++ii) {
                const int i = i0 + ii;
                if (i >= L) break;
                const int64_t off = base + i * sstr;
                if (D == 64) {
                    reinterpret_cast<__nv_bfloat162*>(dq + off)[lane] = __floats2bfloat162_rn(aq[ii][0], aq[ii][NP - 1]);
                    reinterpret_cast<__nv_bfloat162*>(dk + off)[lane] = __floats2bfloat162_rn(ak[ii][0], ak[ii][NP - 1]);
                    reinterpret_cast<__nv_bfloat162*>(dv + off)[lane] = __floats2bfloat162_rn(av[ii][0], av[ii][NP - 1]);
                } else {
                    dq[off + lane] = __float2bfloat16_rn(aq[ii][0]);
                    dk[off + lane] = __float2bfloat16_rn(ak[ii][0]);
                    dv[off + lane] = __float2bfloat16_rn(av[ii][0]);
                }
            }
        }
    }
}

}  // namespace t2v

using namespace t2v;

static int attn_small_config(int L, int D, int tiles, int extra_floats, int& warps, size_t& smem) {
    // as many warps per block as fit ~100 KB, so that two blocks share an SM
    const size_t per_warp = (size_t(tiles) * L * (D + 1) + extra_floats) * sizeof(float);
    warps = int(std::min<size_t>(8, std::max<size_t>(1, (100 * 1024) / per_warp)));
    smem = per_warp * warps;
    return 0;
}

extern "C" {

int t2v_attn_small_fwd(const void* q, const void* k, const void* v, void* o, int64_t nseq, int32_t inner, int64_t outer_rows,
                       int64_t inner_rows, int64_t seq_rows, int64_t ld_in, int64_t ld_out, int32_t heads, int32_t L, int32_t D,
                       void* stream) {
    if (L < 1 || L > kMaxL) return fail(-2, "attn_small: L=%d out of range (1..%d)", L, kMaxL);
    if (D != 64 && D != 32) return fail(-2, "attn_small: head_dim %d unsupported (32 or 64)", D);
    static bool attr_done = false;
    if (!attr_done) {
        cudaFuncSetAttribute(attn_small_fwd_kernel<64>, cudaFuncAttributeMaxDynamicSharedMemorySize, 110 * 1024);
        cudaFuncSetAttribute(attn_small_fwd_kernel<32>, cudaFuncAttributeMaxDynamicSharedMemorySize, 110 * 1024);
        attr_done = true;
    }
    if (ld_in % 8 || ld_out % 8) return fail(-2, "attn_small: row pitches must be multiples of 8 elements");
    SeqAddr a{outer_rows, inner_rows, seq_rows, ld_in, ld_out, inner};
    int warps;
    size_t smem;
    attn_small_config(L, D, 3, 0, warps, smem);
    const int64_t total = nseq * heads;
    const int grid = int(std::min<int64_t>((total + warps - 1) / warps, 148 * 8));
    const float scale = 1.0f / sqrtf(float(D));
    auto Q = static_cast<const __nv_bfloat16*>(q);
    auto K = static_cast<const __nv_bfloat16*>(k);
    auto V = static_cast<const __nv_bfloat16*>(v);
    auto O = static_cast<__nv_bfloat16*>(o);
    cudaStream_t st = static_cast<cudaStream_t>(stream);
    if (D == 64) launch_pdl(attn_small_fwd_kernel<64>, dim3(grid), dim3(warps * 32), size_t(smem), st, Q, K, V, O, a, nseq, heads, L, scale);
    else launch_pdl(attn_small_fwd_kernel<32>, dim3(grid), dim3(warps * 32), size_t(smem), st, Q, K, V, O, a, nseq, heads, L, scale);
    return launch_checked(int(cudaGetLastError()), "attn_small_fwd");
}

int t2v_attn_small_bwd(const void* q, const void* k, const void* v, const void* dout, void* dq, void* dk, void* dv, int64_t nseq,
                       int32_t inner, int64_t outer_rows, int64_t inner_rows, int64_t seq_rows, int64_t ld_in, int64_t ld_out,
                       int32_t heads, int32_t L, int32_t D, void* stream) {
    if (L < 1 || L > kMaxL) return fail(-2, "attn_small: L=%d out of range (1..%d)", L, kMaxL);
    if (D != 64 && D != 32) return fail(-2, "attn_small: head_dim %d unsupported (32 or 64)", D);
    static bool attr_done = false;
    if (!attr_done) {
        cudaFuncSetAttribute(attn_small_bwd_kernel<64>, cudaFuncAttributeMaxDynamicSharedMemorySize, 110 * 1024);
        cudaFuncSetAttribute(attn_small_bwd_kernel<32>, cudaFuncAttributeMaxDynamicSharedMemorySize, 110 * 1024);
        attr_done = true;
    }
    if (ld_in % 8 || ld_out % 8) return fail(-2, "attn_small: row pitches must be multiples of 8 elements");
    SeqAddr a{outer_rows, inner_rows, seq_rows, ld_in, ld_out, inner};
    int warps;
    size_t smem;
    attn_small_config(L, D, 4, 2 * L * L, warps, smem);
    const int64_t total = nseq * heads;
    const int grid = int(std::min<int64_t>((total + warps - 1) / warps, 148 * 8));
    const float scale = 1.0f / sqrtf(float(D));
    auto B = [](const void* p) { return static_cast<const __nv_bfloat16*>(p); };
    auto W = [](void* p) { return static_cast<__nv_bfloat16*>(p); };
    cudaStream_t st = static_cast<cudaStream_t>(stream);
    if (D == 64) launch_pdl(attn_small_bwd_kernel<64>, dim3(grid), dim3(warps * 32), size_t(smem), st, B(q), B(k), B(v), B(dout), W(dq), W(dk), W(dv), a, nseq, heads, L, scale);
    else launch_pdl(attn_small_bwd_kernel<32>, dim3(grid), dim3(warps * 32), size_t(smem), st, B(q), B(k), B(v), B(dout), W(dq), W(dk), W(dv), a, nseq, heads, L, scale);
    return launch_checked(int(cudaGetLastError()), "attn_small_bwd");
}

}  // extern "C"
