// Warp-level tensor-core helpers shared by attn_small.cu (temporal attention) and flash_attn.cu (spatial / cross attention):
// ldmatrix fragment loaders and mma.sync.m16n8k16 (bf16 in, fp32 accumulate) over shared-memory tiles with a byte pitch.
#pragma once
#include <cstdint>

namespace t2v {
namespace wmma16 {

__device__ __forceinline__ void ldsm_x4(uint32_t addr, uint32_t (&r)[4]) {
    asm volatile("ldmatrix.sync.aligned.m8n8.x4.shared.b16 {%0,%1,%2,%3}, [%4];"
                 : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]) : "r"(addr));
}
__device__ __forceinline__ void ldsm_x4_trans(uint32_t addr, uint32_t (&r)[4]) {
    asm volatile("ldmatrix.sync.aligned.m8n8.x4.trans.shared.b16 {%0,%1,%2,%3}, [%4];"
                 : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]) : "r"(addr));
}
__device__ __forceinline__ void mma_bf16(float (&c)[4], const uint32_t (&a)[4], uint32_t b0, uint32_t b1) {
    asm volatile("mma.sync.aligned.m16n8k16.row.col.f32.bf16.bf16.f32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};"
                 : "+f"(c[0]), "+f"(c[1]), "+f"(c[2]), "+f"(c[3])
                 : "r"(a[0]), "r"(a[1]), "r"(a[2]), "r"(a[3]), "r"(b0), "r"(b1));
}

// A fragment (16 x 16 at rows m0.., k-columns k0..) from shared memory.  A_T = false: stored [m][k]; true: stored [k][m].
template <bool A_T>
__device__ __forceinline__ void load_a(uint32_t base, int pitch, int m0, int k0, int lane, uint32_t (&a)[4]) {
    const int q = lane >> 3, r = lane & 7;
    if (!A_T) ldsm_x4(base + (m0 + r + (q & 1) * 8) * pitch + (k0 + (q >> 1) * 8) * 2, a);
    else ldsm_x4_trans(base + (k0 + r + (q >> 1) * 8) * pitch + (m0 + (q & 1) * 8) * 2, a);
}
// B fragments of two adjacent 8-wide n-tiles (n0.., k0..).  B_T = false: stored [n][k] (k contiguous); true: stored [k][n].
template <bool B_T>
__device__ __forceinline__ void load_b2(uint32_t base, int pitch, int n0, int k0, int lane, uint32_t (&b)[4]) {
    const int q = lane >> 3, r = lane & 7;
    if (!B_T) ldsm_x4(base + (n0 + r + (q >> 1) * 8) * pitch + (k0 + (q & 1) * 8) * 2, b);
    else ldsm_x4_trans(base + (k0 + r + (q & 1) * 8) * pitch + (n0 + (q >> 1) * 8) * 2, b);
}

// acc[N/8][4] += A(16 x K at rows m0) * B(K x N), operands in shared memory.
template <int N, int K, bool A_T, bool B_T>
__device__ __forceinline__ void warp_mma(float (&acc)[N / 8][4], uint32_t sa, int pa, int m0, uint32_t sb, int pb, int lane) {
#pragma unroll
    for (int kt = 0; kt < K / 16; ++kt) {
        uint32_t a[4];
        load_a<A_T>(sa, pa, m0, kt * 16, lane, a);
#pragma unroll
        for (int np = 0; np < N / 16; ++np) {
            uint32_t b[4];
            load_b2<B_T>(sb, pb, np * 16, kt * 16, lane, b);
            mma_bf16(acc[2 * np], a, b[0], b[1]);
            mma_bf16(acc[2 * np + 1], a, b[2], b[3]);
        }
    }
}

}  // namespace wmma16
}  // namespace t2v
