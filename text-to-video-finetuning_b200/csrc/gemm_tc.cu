// Persistent warp-specialised tcgen05 GEMM for sm_100a.  See gemm_tc.cuh for the operand model.
//
//   warp 0 : TMA producer   (one elected lane) - fills the smem ring, one mbarrier pair per stage
//   warp 1 : MMA issuer     (one elected lane) - tcgen05.mma into one of two TMEM accumulator stages
//   warps 2-9 : epilogue    (2 x 128 threads, thread <-> accumulator row, the two groups split the columns) -
//               tcgen05.ld, fused epilogue, smem-staged coalesced global stores / red.add
//
// The accumulator is double-buffered in TMEM so the epilogue of tile i overlaps the main loop of tile i+1.
#include "common.h"
#include "gemm_tc.cuh"
#include "ptx.cuh"

#include <cstdio>
#include <mutex>

namespace t2v {

struct TileVars {
    int32_t t[6];
};

// tile id -> six mixed-radix digits, with host-precomputed magic numbers (q = umulhi(n, mul) >> shr; n < 2^31)
__device__ __forceinline__ void decompose_tile(const GemmParams& p, int32_t tile, TileVars& tv) {
    uint32_t n = static_cast<uint32_t>(tile);
#pragma unroll
    for (int i = 0; i < 6; ++i) {
        const uint32_t d = static_cast<uint32_t>(p.tdim[i]);
        const uint32_t q = d == 1u ? n : (__umulhi(n, p.tdiv_mul[i]) >> p.tdiv_shr[i]);
        tv.t[i] = static_cast<int32_t>(n - q * d);
        n = q;
    }
    if (p.mh == 2) {  // 256-row tiles: t[pair_var] counts PAIRS of row tiles; half h of the tile is row tile 2*t + h
#pragma unroll
        for (int i = 0; i < 6; ++i) tv.t[i] = (i == p.pair_var) ? tv.t[i] * 2 : tv.t[i];
    }
}

__device__ __forceinline__ void k_range(const GemmParams& p, const TileVars& tv, int32_t& kb0, int32_t& kb1) {
    if (p.ksplit_var >= 0) {
        int32_t sv = 0;
#pragma unroll
        for (int i = 0; i < 6; ++i) sv = (i == p.ksplit_var) ? tv.t[i] : sv;
        kb0 = sv * p.kb_per_split;
        kb1 = min(p.kb_total, kb0 + p.kb_per_split);
    } else {
        kb0 = 0;
        kb1 = p.kb_total;
    }
}

// TMA coordinates of one operand at (tile, k-loop digits kv)
__device__ __forceinline__ void operand_coords(const TmaOperand& op, const TileVars& tv, const int32_t* kv, int32_t* c) {
#pragma unroll
    for (int d = 0; d < 5; ++d) {
        int32_t v = op.base[d];
#pragma unroll
        for (int i = 0; i < 6; ++i) v += tv.t[i] * op.tcoef[d][i];
        v += kv[0] * op.kcoef[d][0] + kv[1] * op.kcoef[d][1] + kv[2] * op.kcoef[d][2];
        c[d] = v;
    }
}

__device__ __forceinline__ void issue_boxes(const TmaOperand& op, const int32_t* c_in, uint8_t* smem_dst, uint64_t* bar) {
    int32_t c[5] = {c_in[0], c_in[1], c_in[2], c_in[3], c_in[4]};
    for (int b = 0; b < op.nbox; ++b) {  // boxes of an MN-major operand always advance along dimension 0
        tma_load(op.rank, smem_dst + b * op.box_bytes, &op.map, bar, c);
        c[0] += op.box_step;
    }
}

// Epilogue of one CTA (warps 2..9).  mh == 1: the two groups of four warps split the accumulator columns; mh == 2: group g
// drains row half g.  Within a group warp w owns TMEM lanes 32*(w%4)..+31, i.e. thread <-> output row.  Output rows are
// strided in global memory, so every 32-column chunk goes through a per-warp swizzled staging tile and is moved with
// 16 bytes per lane covering whole rows (full 64/128-byte segments); the residual comes in the same way.  The per-column
// bias is fetched with one coalesced load per chunk and broadcast through shared memory.
// flags of the epilogue fast path: known at compile time (folded away) or only at run time (the catch-all instantiation)
template <bool V>
struct ConstFlag {
    __device__ constexpr operator bool() const { return V; }
};
struct RuntimeFlag {
    bool v;
    __device__ constexpr operator bool() const { return v; }
};

template <int ESZ>
__device__ __forceinline__ void run_epilogue(const GemmParams& p, uint32_t warp, uint32_t lane, uint32_t tmem_base,
                                             uint64_t* acc_full, uint64_t* acc_empty, uint8_t* stg_base) {
    constexpr int CPR = ESZ * 2;        // 16-byte chunks per 32-column row: 4 (bf16) or 8 (fp32)
    constexpr int EPC = 16 / ESZ;       // elements per 16-byte chunk
    constexpr int RPI = 32 / CPR;       // rows covered by one warp-wide 16-byte access
    const uint32_t ew = warp - 2u;
    const uint32_t quad = warp & 3u;
    const uint32_t grp = ew >> 2;
    const uint32_t row = quad * 32u + lane;
    const int32_t rw = static_cast<int32_t>(row) % p.bw;
    const int32_t rh = (static_cast<int32_t>(row) / p.bw) % p.bh;
    const int32_t rn = static_cast<int32_t>(row) / (p.bw * p.bh);
    const bool has_bias = p.flags & EPI_BIAS, has_rb = p.flags & EPI_ROWBIAS, has_res = p.flags & EPI_RESIDUAL;
    const bool vec = p.flags & EPI_VEC;
    const bool has_stats = (p.flags & EPI_STATS) && vec;
    const int nchunks = (p.block_n + 31) >> 5;
    const int c_begin = (p.mh == 2 || grp == 0) ? 0 : (nchunks + 1) >> 1;
    const int c_end = p.mh == 2 ? nchunks : (grp == 0 ? (nchunks + 1) >> 1 : nchunks);
    uint8_t* stg = stg_base + ew * 4096u;                                    // 32 rows x <= 128 B
    float* sbias = reinterpret_cast<float*>(stg_base + 8 * 4096u) + ew * 32;  // 32 floats per warp
    // swizzled byte offset of logical 16-byte chunk j of row r in a dense [32][n] chunk array
    auto phys = [](int r, int j, int n) { return (r * n + (j ^ (((r * n) >> 3) & (n - 1)))) * 16; };
    const int sr = lane / CPR, sj = lane % CPR;  // (row-in-group, chunk) this lane moves in the coalesced phases
    // tile-invariant shared-memory addresses of the staging tile (hoisted: the fast path below is straight-line code)
    const uint32_t stg_s = smem_u32(stg);
    uint32_t w_own[CPR];     // this lane's own row, 16-byte chunk j                     (write after the math)
    uint32_t r_mov[CPR];     // the (row, chunk) this lane moves in the coalesced store   (read)
    uint32_t w_res[4];       // residual staging: row (lane / 4) + 8 i, chunk lane % 4   (write, bf16 rows of 64 B)
    uint32_t r_res[4];       // residual staging: own row, chunk j                       (read)
#pragma unroll
    for (int j = 0; j < CPR; ++j) {
        w_own[j] = stg_s + phys(static_cast<int>(lane), j, CPR);
        r_mov[j] = stg_s + phys(sr + RPI * j, sj, CPR);
    }
#pragma unroll
    for (int j = 0; j < 4; ++j) {
        w_res[j] = stg_s + phys(static_cast<int>(lane >> 2) + 8 * j, static_cast<int>(lane & 3), 4);
        r_res[j] = stg_s + phys(static_cast<int>(lane), j, 4);
    }
    const bool alpha_one = p.alpha == 1.0f;
    uint32_t it = 0;
    for (int32_t tile = blockIdx.x; tile < p.num_tiles; tile += gridDim.x, ++it) {
        TileVars tv;
        decompose_tile(p, tile, tv);
        const bool rowsum_tile = (p.flags & EPI_ROWSUM_A) && tv.t[0] == 0 && tv.t[2] == 0 && tv.t[3] == 0 && (p.mh == 2 || grp == 0);
        if (p.mh == 2 && grp == 1) {
#pragma unroll
            for (int i = 0; i < 6; ++i) tv.t[i] += (i == p.pair_var) ? 1 : 0;
        }
        const int32_t gw = tv.t[1] * p.bw + rw, gh = tv.t[2] * p.bh + rh, gn = tv.t[3] * p.bn + rn;
        const bool row_ok = (rn < p.bn) && gw < p.W && gh < p.H && gn < p.N;
        int64_t off = gw * p.ldw + gh * p.ldh + gn * p.ldn;
#pragma unroll
        for (int i = 0; i < 6; ++i) off += tv.t[i] * p.otc[i];
        // rows this lane moves in the coalesced phases (constant over the tile's chunks)
        int64_t off_s[CPR];
        uint32_t ok_s = 0;
#pragma unroll
        for (int i = 0; i < CPR; ++i) {
            const int r = sr + RPI * i;
            off_s[i] = __shfl_sync(0xffffffffu, off, r);
            ok_s |= (__shfl_sync(0xffffffffu, row_ok ? 1u : 0u, r) & 1u) << i;
        }
        const int32_t col0 = tv.t[0] * p.block_n;
        // every row of this warp inside the output: the common case runs without per-row predicates
        const bool all_rows = vec && __all_sync(0xffffffffu, row_ok);
        // GroupNorm statistics: frame of this lane's segment (taken from the segment's first row, which is in bounds
        // whenever any row of the segment is)
        int64_t st_off = 0;
        uint32_t okmask = 0;
        if (has_stats) {
            const int32_t fr = (gw * p.st_cw + gh * p.st_ch + gn * p.st_cn) / p.st_div;
            st_off = int64_t(__shfl_sync(0xffffffffu, fr, lane & ~uint32_t(p.st_seg - 1))) * p.st_ld;
            okmask = __ballot_sync(0xffffffffu, row_ok);
        }
        // ---- software pipeline of the epilogue's global loads.  Per-column additive terms (bias, and the time-embedding row
        // when every row of this warp takes the same one) are fetched two chunks ahead, the residual tile one chunk ahead;
        // the first of them are issued BEFORE the wait for the accumulator, so their latency hides behind the main loop.
        // (every lane takes part in the shuffle: it must not sit behind a short-circuit that depends on row_ok)
        // The row is taken from the warp's first IN-BOUNDS lane: an out-of-bounds lane (frames beyond N in the last tile) would
        // name a row past the end of the time-embedding matrix.  A warp without any valid row folds nothing.
        const int32_t rb_row = has_rb ? gn / p.rb_div : 0;
        const uint32_t ok_lanes = __ballot_sync(0xffffffffu, row_ok);
        const int32_t rb_first = __shfl_sync(0xffffffffu, rb_row, ok_lanes ? __ffs(ok_lanes) - 1 : 0);
        const bool rb_same = !row_ok || rb_row == rb_first;
        const bool rb_uniform = has_rb && vec && ok_lanes != 0u && __all_sync(0xffffffffu, rb_same);
        const float* rb_base = has_rb ? p.rowbias + static_cast<int64_t>(rb_first) * p.rb_ld : nullptr;
        auto load_colterm = [&](int ch) {
            const int32_t c = col0 + ch * 32 + static_cast<int32_t>(lane);
            float b = 0.f;
            if (static_cast<int32_t>(lane) < min(32, min(p.block_n - ch * 32, p.ncols - (col0 + ch * 32)))) {
                if (has_bias) b = __ldg(p.bias + c);
                if (rb_uniform) b += __ldg(rb_base + c);
            }
            return b;
        };
        const bool res_pref = all_rows && has_res;
        const __nv_bfloat16* res_lane = static_cast<const __nv_bfloat16*>(p.residual) + (lane & 3) * 8;
        int64_t off_q[4];   // residual rows this lane fetches: (lane / 4) + 8 i
#pragma unroll
        for (int i = 0; i < 4; ++i) off_q[i] = __shfl_sync(0xffffffffu, off, (lane >> 2) + 8 * i);
        auto load_res = [&](int ch, uint4 (&q)[4]) {
            const int32_t c = col0 + ch * 32;
            if (min(p.block_n - ch * 32, p.ncols - c) >= 32) {
#pragma unroll
                for (int i = 0; i < 4; ++i) q[i] = __ldg(reinterpret_cast<const uint4*>(res_lane + off_q[i] + c));
            }
        };
        float bq0 = 0.f, bq1 = 0.f;
        uint4 rq[4] = {};
        if (c_begin < c_end) bq0 = load_colterm(c_begin);
        if (c_begin + 1 < c_end) bq1 = load_colterm(c_begin + 1);
        if (res_pref && c_begin < c_end) load_res(c_begin, rq);
        const uint32_t as = p.nacc == 2 ? (it & 1u) : 0u;
        const uint32_t aphase = p.nacc == 2 ? ((it >> 1) & 1u) : (it & 1u);
        mbar_wait(&acc_full[as], aphase);
        tc_fence_after();
        const uint32_t taddr = tmem_base + ((quad * 32u) << 16) + as * p.acc_stage_cols + (p.mh == 2 ? grp * p.acc_half_cols : 0u);
        uint32_t acc[32];
        if (c_begin < c_end) tmem_ld32(taddr + c_begin * 32, acc);  // software pipeline: chunk ch+1 is read from TMEM while
        for (int ch = c_begin; ch < c_end; ++ch) {                  // chunk ch goes through staging and out to global memory
            const int32_t col = col0 + ch * 32;
            const int32_t cvalid = min(32, min(p.block_n - ch * 32, p.ncols - col));  // valid columns of this chunk
            __syncwarp();
            const float bcur = bq0;          // this chunk's per-column term; keep the pipeline two chunks deep
            bq0 = bq1;
            if (ch + 2 < c_end) bq1 = load_colterm(ch + 2);
            const bool colterm = has_bias || rb_uniform;
            if (all_rows && cvalid == 32) {
                // ---- fast path: a full 32 x 32 chunk, every row in bounds.  The body is instantiated per combination of
                // (column term, residual, statistics, alpha, per-row time embedding) so that each instantiation is
                // straight-line code without flag tests: ncu counted 334 warp-instructions per chunk in the flag-driven
                // version, a third of them predicate bookkeeping (profiles/r2_ncu_full_gemm_ff_proj_before.txt).
                auto fast = [&](auto kCol, auto kRes, auto kStats, auto kAlpha, auto kRowRb) {
                    // each flag is a ConstFlag (folded at compile time) or a RuntimeFlag (the catch-all instantiation)
                    const bool COL = kCol, RES = kRes, STATS = kStats, ALPHA = kAlpha, ROWRB = kRowRb;
                    if (COL) sbias[lane] = bcur;
                    if (RES) {   // residual (prefetched one chunk ago): registers -> staging, then fetch the next chunk's
#pragma unroll
                        for (int i = 0; i < 4; ++i) sts128(w_res[i], rq[i]);
                        if (ch + 1 < c_end) load_res(ch + 1, rq);
                    }
                    tmem_ld_wait();
                    float* v = reinterpret_cast<float*>(acc);   // in place: the accumulator registers are not reloaded until the
                    if (ALPHA) {                      // values have been packed into the staging tile
#pragma unroll
                        for (int j = 0; j < 32; ++j) v[j] *= p.alpha;
                    }
                    if (COL || RES) __syncwarp();
                    if (COL) {
#pragma unroll
                        for (int j = 0; j < 8; ++j) {
                            const float4 b = *reinterpret_cast<const float4*>(sbias + 4 * j);  // shared-memory broadcast
                            v[4 * j] += b.x; v[4 * j + 1] += b.y; v[4 * j + 2] += b.z; v[4 * j + 3] += b.w;
                        }
                    }
                    if (ROWRB) {
                        const float4* b4 = reinterpret_cast<const float4*>(p.rowbias + (gn / p.rb_div) * p.rb_ld + col);
#pragma unroll
                        for (int j = 0; j < 8; ++j) {
                            const float4 b = __ldg(b4 + j);
                            v[4 * j] += b.x; v[4 * j + 1] += b.y; v[4 * j + 2] += b.z; v[4 * j + 3] += b.w;
                        }
                    }
                    if (RES) {
#pragma unroll
                        for (int j = 0; j < 4; ++j) {
                            const uint4 q = lds128(r_res[j]);
                            v[8 * j + 0] += bf16_lo(q.x); v[8 * j + 1] += bf16_hi(q.x);
                            v[8 * j + 2] += bf16_lo(q.y); v[8 * j + 3] += bf16_hi(q.y);
                            v[8 * j + 4] += bf16_lo(q.z); v[8 * j + 5] += bf16_hi(q.z);
                            v[8 * j + 6] += bf16_lo(q.w); v[8 * j + 7] += bf16_hi(q.w);
                        }
                        __syncwarp();
                    }
                    if constexpr (ESZ == 2) {
#pragma unroll
                        for (int j = 0; j < 4; ++j) {
                            uint4 q;
                            q.x = pack_bf16(v[8 * j + 0], v[8 * j + 1]);
                            q.y = pack_bf16(v[8 * j + 2], v[8 * j + 3]);
                            q.z = pack_bf16(v[8 * j + 4], v[8 * j + 5]);
                            q.w = pack_bf16(v[8 * j + 6], v[8 * j + 7]);
                            sts128(w_own[j], q);
                        }
                    } else {
#pragma unroll
                        for (int j = 0; j < CPR; ++j) sts128(w_own[j], make_uint4(acc[4 * j], acc[4 * j + 1], acc[4 * j + 2], acc[4 * j + 3]));
                    }
                    if (ch + 1 < c_end) tmem_ld32(taddr + (ch + 1) * 32, acc);   // next chunk: in flight during the store phase
                    __syncwarp();
                    if (ESZ == 2 && STATS) {   // GroupNorm statistics of the staged bf16 tile (see the general path below)
                        const uint32_t half = lane >> 4, cw = lane & 15u;
                        float s0 = 0.f, q0 = 0.f, s1 = 0.f, q1 = 0.f;
#pragma unroll
                        for (int i = 0; i < 16; ++i) {
                            const int r = half ? 16 + ((i + 1) & 15) : i;
                            const uint32_t w = lds32(stg_s + phys(r, int(cw >> 2), 4) + (cw & 3u) * 4u);
                            const float lo = bf16_lo(w), hi = bf16_hi(w);
                            s0 += lo; q0 += lo * lo;
                            s1 += hi; q1 += hi * hi;
                        }
                        if (p.st_seg == 32) {
                            s0 += __shfl_xor_sync(0xffffffffu, s0, 16); q0 += __shfl_xor_sync(0xffffffffu, q0, 16);
                            s1 += __shfl_xor_sync(0xffffffffu, s1, 16); q1 += __shfl_xor_sync(0xffffffffu, q1, 16);
                        }
                        if (p.st_seg == 16 || half == 0) red_add_f32x4(p.stats + (st_off + col + 2 * cw) * 2, s0, q0, s1, q1);
                    }
                    // staging -> global: 16 bytes per lane, whole row segments
                    char* ob = static_cast<char*>(p.out) + (static_cast<int64_t>(col) + sj * EPC) * ESZ;
#pragma unroll
                    for (int i = 0; i < CPR; ++i) {
                        const uint4 q = lds128(r_mov[i]);
                        char* o = ob + off_s[i] * ESZ;
                        if (ESZ == 2 || p.out_mode == OUT_F32) {
                            *reinterpret_cast<uint4*>(o) = q;
                        } else {
                            red_add_f32x4(reinterpret_cast<float*>(o), __uint_as_float(q.x), __uint_as_float(q.y), __uint_as_float(q.z),
                                          __uint_as_float(q.w));
                        }
                    }
                };
                using T = ConstFlag<true>;
                using F = ConstFlag<false>;
                const bool row_rb = has_rb && !rb_uniform;
                const int key = (colterm ? 1 : 0) | (has_res ? 2 : 0) | ((ESZ == 2 && has_stats) ? 4 : 0) | (alpha_one ? 0 : 8) | (row_rb ? 16 : 0);
                switch (key) {
                    case 0: fast(F{}, F{}, F{}, F{}, F{}); break;     // plain (dgrad, attention products)
                    case 1: fast(T{}, F{}, F{}, F{}, F{}); break;     // bias (+ folded time embedding)
                    case 2: fast(F{}, T{}, F{}, F{}, F{}); break;     // residual
                    case 3: fast(T{}, T{}, F{}, F{}, F{}); break;     // bias + residual
                    case 4: fast(F{}, F{}, T{}, F{}, F{}); break;     // ... the same with GroupNorm statistics
                    case 5: fast(T{}, F{}, T{}, F{}, F{}); break;
                    case 6: fast(F{}, T{}, T{}, F{}, F{}); break;
                    case 7: fast(T{}, T{}, T{}, F{}, F{}); break;
                    default: {                                         // alpha != 1 / per-row time embedding: rare, runtime flags
                        using R = RuntimeFlag;
                        fast(R{colterm}, R{has_res}, R{ESZ == 2 && has_stats}, R{!alpha_one}, R{row_rb});
                        break;
                    }
                }
                continue;
            }
            if (res_pref && ch + 1 < c_end) load_res(ch + 1, rq);   // keep the residual pipeline primed for a following fast chunk
            const float bval = bcur;   // bias (+ the warp-uniform time-embedding row) of this chunk, fetched two chunks ago
            if (vec && has_res && cvalid > 0) {         // residual: coalesced global -> staging (bf16, 4 chunks per row)
#pragma unroll
                for (int i = 0; i < 4; ++i) {
                    const int r = (lane >> 2) + 8 * i, j = lane & 3;
                    const int64_t off_r = __shfl_sync(0xffffffffu, off, r);
                    const bool ok_r = __shfl_sync(0xffffffffu, row_ok ? 1 : 0, r);
                    uint4 q = make_uint4(0, 0, 0, 0);
                    if (ok_r && j * 8 + 8 <= cvalid)
                        q = __ldg(reinterpret_cast<const uint4*>(static_cast<const __nv_bfloat16*>(p.residual) + off_r + col + j * 8));
                    *reinterpret_cast<uint4*>(stg + phys(r, j, 4)) = q;
                }
            }
            sbias[lane] = bval;
            tmem_ld_wait();
            float v[32];
#pragma unroll
            for (int j = 0; j < 32; ++j) v[j] = __uint_as_float(acc[j]) * p.alpha;
            if (ch + 1 < c_end) tmem_ld32(taddr + (ch + 1) * 32, acc);
            __syncwarp();
            if (cvalid <= 0) continue;
            if (vec) {
                if (colterm) {
#pragma unroll
                    for (int j = 0; j < 8; ++j) {
                        const float4 b = *reinterpret_cast<const float4*>(sbias + 4 * j);  // shared-memory broadcast
                        v[4 * j] += b.x; v[4 * j + 1] += b.y; v[4 * j + 2] += b.z; v[4 * j + 3] += b.w;
                    }
                }
                if (has_rb && !rb_uniform && row_ok) {
                    const float4* b4 = reinterpret_cast<const float4*>(p.rowbias + (gn / p.rb_div) * p.rb_ld + col);
#pragma unroll
                    for (int j = 0; j < 8; ++j) {
                        if (j * 4 + 4 <= cvalid) {
                            const float4 b = __ldg(b4 + j);
                            v[4 * j] += b.x; v[4 * j + 1] += b.y; v[4 * j + 2] += b.z; v[4 * j + 3] += b.w;
                        }
                    }
                }
                if (has_res) {
#pragma unroll
                    for (int j = 0; j < 4; ++j) {
                        const uint4 q = *reinterpret_cast<const uint4*>(stg + phys(lane, j, 4));
                        v[8 * j + 0] += bf16_lo(q.x); v[8 * j + 1] += bf16_hi(q.x);
                        v[8 * j + 2] += bf16_lo(q.y); v[8 * j + 3] += bf16_hi(q.y);
                        v[8 * j + 4] += bf16_lo(q.z); v[8 * j + 5] += bf16_hi(q.z);
                        v[8 * j + 6] += bf16_lo(q.w); v[8 * j + 7] += bf16_hi(q.w);
                    }
                    __syncwarp();
                }
                // own row -> staging
                if (ESZ == 2) {
#pragma unroll
                    for (int j = 0; j < 4; ++j) {
                        uint4 q;
                        q.x = pack_bf16(v[8 * j + 0], v[8 * j + 1]);
                        q.y = pack_bf16(v[8 * j + 2], v[8 * j + 3]);
                        q.z = pack_bf16(v[8 * j + 4], v[8 * j + 5]);
                        q.w = pack_bf16(v[8 * j + 6], v[8 * j + 7]);
                        *reinterpret_cast<uint4*>(stg + phys(lane, j, 4)) = q;
                    }
                } else {
#pragma unroll
                    for (int j = 0; j < 8; ++j)
                        *reinterpret_cast<float4*>(stg + phys(lane, j, 8)) = make_float4(v[4 * j], v[4 * j + 1], v[4 * j + 2], v[4 * j + 3]);
                }
                __syncwarp();
                if (ESZ == 2 && has_stats) {
                    // GroupNorm statistics of the tile just staged (the bf16 values the consumer will read): lanes 0-15 walk
                    // rows 0-15, lanes 16-31 rows 16-31 (staggered by one row so the two halves hit different banks), each lane
                    // owns the column pair (2c, 2c+1), c = lane % 16.  16-row segments: every half is one frame; 32-row
                    // segments: the halves are combined with one shuffle.
                    const uint32_t half = lane >> 4, cw = lane & 15u;
                    float s0 = 0.f, q0 = 0.f, s1 = 0.f, q1 = 0.f;
#pragma unroll
                    for (int i = 0; i < 16; ++i) {
                        const int r = half ? 16 + ((i + 1) & 15) : i;
                        const uint32_t w = *reinterpret_cast<const uint32_t*>(stg + phys(r, int(cw >> 2), 4) + (cw & 3u) * 4u);
                        if ((okmask >> r) & 1u) {
                            const float lo = bf16_lo(w), hi = bf16_hi(w);
                            s0 += lo; q0 += lo * lo;
                            s1 += hi; q1 += hi * hi;
                        }
                    }
                    if (p.st_seg == 32) {
                        s0 += __shfl_xor_sync(0xffffffffu, s0, 16); q0 += __shfl_xor_sync(0xffffffffu, q0, 16);
                        s1 += __shfl_xor_sync(0xffffffffu, s1, 16); q1 += __shfl_xor_sync(0xffffffffu, q1, 16);
                    }
                    // a segment without a valid row has no frame (its offset is meaningless): it contributes nothing
                    const bool any_row = p.st_seg == 16 ? ((okmask >> (half * 16u)) & 0xFFFFu) != 0u : okmask != 0u;
                    if (any_row && (p.st_seg == 16 || half == 0) && int(2 * cw) < cvalid)
                        red_add_f32x4(p.stats + (st_off + col + 2 * cw) * 2, s0, q0, s1, q1);
                }
                // staging -> global: 16 bytes per lane, whole row segments
                const int nval = cvalid - sj * EPC;  // valid elements in this lane's chunk
#pragma unroll
                for (int i = 0; i < CPR; ++i) {
                    if (!((ok_s >> i) & 1u) || nval <= 0) continue;
                    const int r = sr + RPI * i;
                    const uint4 q = *reinterpret_cast<const uint4*>(stg + phys(r, sj, CPR));
                    const int64_t o = off_s[i] + col + sj * EPC;
                    if (nval >= EPC) {
                        if (ESZ == 2) {
                            *reinterpret_cast<uint4*>(static_cast<__nv_bfloat16*>(p.out) + o) = q;
                        } else if (p.out_mode == OUT_F32) {
                            *reinterpret_cast<uint4*>(static_cast<float*>(p.out) + o) = q;
                        } else {
                            red_add_f32x4(static_cast<float*>(p.out) + o, __uint_as_float(q.x), __uint_as_float(q.y), __uint_as_float(q.z),
                                          __uint_as_float(q.w));
                        }
                    } else {  // ragged last chunk: element-wise
                        // (static indexing only: a runtime-indexed register array would be demoted to local memory)
                        const uint32_t w4[4] = {q.x, q.y, q.z, q.w};
#pragma unroll
                        for (int e = 0; e < EPC; ++e) {
                            if (e >= nval) break;
                            if (ESZ == 2) {
                                const uint32_t h = (w4[e >> 1] >> ((e & 1) * 16)) & 0xFFFFu;
                                reinterpret_cast<uint16_t*>(p.out)[o + e] = static_cast<uint16_t>(h);
                            } else if (p.out_mode == OUT_F32) {
                                static_cast<float*>(p.out)[o + e] = __uint_as_float(w4[e & 3]);
                            } else {
                                atomicAdd(static_cast<float*>(p.out) + o + e, __uint_as_float(w4[e & 3]));
                            }
                        }
                    }
                }
            } else if (row_ok) {
                // unaligned buffers: per-thread scalar path
#pragma unroll
                for (int j = 0; j < 32; ++j) {
                    if (j >= cvalid) break;
                    float x = v[j];
                    if (has_bias) x += p.bias[col + j];
                    if (has_rb) x += p.rowbias[(gn / p.rb_div) * p.rb_ld + col + j];
                    if (has_res) x += __bfloat162float(static_cast<const __nv_bfloat16*>(p.residual)[off + col + j]);
                    if (ESZ == 2) static_cast<__nv_bfloat16*>(p.out)[off + col + j] = __float2bfloat16_rn(x);
                    else if (p.out_mode == OUT_F32) static_cast<float*>(p.out)[off + col + j] = x;
                    else atomicAdd(static_cast<float*>(p.out) + off + col + j, x);
                }
            }
        }
        if (rowsum_tile) {
            // row sums of operand A (weight gradient: the bias gradient): 16 identical accumulator columns, column 0 is taken.
            // mh == 1: both warp groups see all 128 rows (they split the columns) - group 0 alone adds them.
            tmem_ld32(taddr + p.rowsum_col, acc);
            tmem_ld_wait();
            if (row_ok) atomicAdd(p.rowsum + gw, __uint_as_float(acc[0]));
        }
        tc_fence_before();
        mbar_arrive(&acc_empty[as]);
    }
}

template <bool A_MN, bool B_MN>
__global__ void __launch_bounds__(kNumThreads, 1) gemm_tc_kernel(const __grid_constant__ GemmParams p) {
    extern __shared__ uint8_t smem_raw[];
    // 1024-byte alignment is required by the 128-byte swizzle atoms.
    uint8_t* smem = smem_raw + ((1024u - (smem_u32(smem_raw) & 1023u)) & 1023u);
    const int S = p.num_stages;
    const bool rowsum_a = p.flags & EPI_ROWSUM_A;
    const uint8_t* ones_tile = smem;              // EPI_ROWSUM_A: 16 rows x 128 bytes of bf16 1.0 (any layout: all elements equal)
    if (rowsum_a) smem += kOnesTileBytes;
    uint8_t* smem_a = smem;
    uint8_t* smem_b = smem + static_cast<size_t>(S) * p.stage_bytes_a;
    uint64_t* bars = reinterpret_cast<uint64_t*>(smem_b + static_cast<size_t>(S) * p.stage_bytes_b);
    uint64_t* full_bar = bars;                       // [S]   TMA -> MMA
    uint64_t* empty_bar = bars + kMaxStages;         // [S]   MMA -> TMA
    uint64_t* acc_full = bars + 2 * kMaxStages;      // [2]   MMA -> epilogue
    uint64_t* acc_empty = bars + 2 * kMaxStages + 2; // [2]   epilogue -> MMA
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 2 * kMaxStages + 4);

    const uint32_t warp = threadIdx.x >> 5;
    const uint32_t lane = lane_id();

    if (warp == 0 && lane == 0) {
        tma_prefetch_desc(&p.a.map);
        tma_prefetch_desc(&p.b.map);
        for (int s = 0; s < S; ++s) {
            mbar_init(&full_bar[s], 1);
            mbar_init(&empty_bar[s], 1);
        }
        for (int s = 0; s < 2; ++s) {
            mbar_init(&acc_full[s], 1);
            mbar_init(&acc_empty[s], 256);
        }
        fence_mbar_init();
    }
    if (warp == 1) tmem_alloc(tmem_slot, kTmemCols);
    if (rowsum_a && warp == 2) {   // the constant B operand of the row-sum MMAs; made visible to the tensor core (async proxy)
        const uint32_t one2 = 0x3F803F80u;         // two bf16 1.0
        uint4* o = reinterpret_cast<uint4*>(const_cast<uint8_t*>(ones_tile));
        for (uint32_t i = lane; i < kOnesTileBytes / 16; i += 32) o[i] = make_uint4(one2, one2, one2, one2);
        asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
    }
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem_base = *tmem_slot;
    pdl_sync();  // the prologue above (barriers, TMEM, descriptor prefetch) overlaps the tail of the previous kernel

    if (warp == 0) {
        // ------------------------------------------------------------------ TMA producer
        if (elect_one()) {
            uint32_t stage = 0, phase = 0;
            const uint32_t tx_bytes = p.stage_bytes_a + p.stage_bytes_b;  // stage_bytes_a covers both row halves when mh == 2
            int32_t half_delta[5];
#pragma unroll
            for (int d = 0; d < 5; ++d) {
                int32_t v = 0;
#pragma unroll
                for (int i = 0; i < 6; ++i) v = (i == p.pair_var) ? p.a.tcoef[d][i] : v;
                half_delta[d] = v;
            }
            for (int32_t tile = blockIdx.x; tile < p.num_tiles; tile += gridDim.x) {
                TileVars tv;
                decompose_tile(p, tile, tv);
                int32_t kb0, kb1;
                k_range(p, tv, kb0, kb1);
                // k-loop digits advance as a mixed-radix counter; coordinates are updated incrementally (adds only) and
                // recomputed from scratch only when the fastest digit wraps
                int32_t kv[3];
                kv[0] = kb0 % p.kdim[0];
                kv[1] = (kb0 / p.kdim[0]) % p.kdim[1];
                kv[2] = kb0 / (p.kdim[0] * p.kdim[1]);
                int32_t ca[5], cb[5];
                operand_coords(p.a, tv, kv, ca);
                operand_coords(p.b, tv, kv, cb);
                for (int32_t kb = kb0; kb < kb1; ++kb) {
                    mbar_wait(&empty_bar[stage], phase ^ 1);
                    mbar_expect_tx(&full_bar[stage], tx_bytes);
                    issue_boxes(p.a, ca, smem_a + static_cast<size_t>(stage) * p.stage_bytes_a, &full_bar[stage]);
                    if (p.mh == 2) {
                        int32_t c1[5];
#pragma unroll
                        for (int d = 0; d < 5; ++d) c1[d] = ca[d] + half_delta[d];
                        issue_boxes(p.a, c1, smem_a + static_cast<size_t>(stage) * p.stage_bytes_a + kBlockM * 128, &full_bar[stage]);
                    }
                    issue_boxes(p.b, cb, smem_b + static_cast<size_t>(stage) * p.stage_bytes_b, &full_bar[stage]);
                    if (++kv[0] < p.kdim[0]) {
#pragma unroll
                        for (int d = 0; d < 5; ++d) {
                            ca[d] += p.a.kcoef[d][0];
                            cb[d] += p.b.kcoef[d][0];
                        }
                    } else {
                        kv[0] = 0;
                        if (++kv[1] == p.kdim[1]) {
                            kv[1] = 0;
                            ++kv[2];
                        }
                        operand_coords(p.a, tv, kv, ca);
                        operand_coords(p.b, tv, kv, cb);
                    }
                    if (++stage == static_cast<uint32_t>(S)) {
                        stage = 0;
                        phase ^= 1;
                    }
                }
            }
        }
    } else if (warp == 1) {
        // ------------------------------------------------------------------ MMA issuer
        if (elect_one()) {
            const uint32_t idesc = make_idesc_bf16(static_cast<uint32_t>(p.block_n), A_MN, B_MN);
            // EPI_ROWSUM_A: D[128 x 16] += A_tile * ones: the same A descriptor against a K-major tile of ones (never advanced
            // along K - every element is 1.0), accumulated next to the tile's own columns
            const uint32_t idesc_ones = make_idesc_bf16(16u, A_MN, false);
            const uint64_t d_ones = make_sw128_desc(smem_u32(ones_tile), 0u, 1024);
            // descriptor advance per UMMA_K=16 step, in 16-byte units
            constexpr uint32_t a_kstep = A_MN ? (16u * 128u) >> 4 : 32u >> 4;
            constexpr uint32_t b_kstep = B_MN ? (16u * 128u) >> 4 : 32u >> 4;
            constexpr uint32_t a_lbo = A_MN ? 64u * 128u : 0u;
            constexpr uint32_t b_lbo = B_MN ? 64u * 128u : 0u;
            uint32_t stage = 0, phase = 0;
            uint32_t it = 0;
            for (int32_t tile = blockIdx.x; tile < p.num_tiles; tile += gridDim.x, ++it) {
                TileVars tv;
                decompose_tile(p, tile, tv);
                int32_t kb0, kb1;
                k_range(p, tv, kb0, kb1);
                const uint32_t as = p.nacc == 2 ? (it & 1u) : 0u;
                const uint32_t aphase = p.nacc == 2 ? ((it >> 1) & 1u) : (it & 1u);
                mbar_wait(&acc_empty[as], aphase ^ 1);
                tc_fence_after();
                const uint32_t tmem_d = tmem_base + as * p.acc_stage_cols;
                const bool rowsum_tile = rowsum_a && tv.t[0] == 0 && tv.t[2] == 0 && tv.t[3] == 0;
                for (int32_t kb = kb0; kb < kb1; ++kb) {
                    mbar_wait(&full_bar[stage], phase);
                    tc_fence_after();
                    const uint64_t da = make_sw128_desc(smem_u32(smem_a + static_cast<size_t>(stage) * p.stage_bytes_a), a_lbo, 1024);
                    const uint64_t db = make_sw128_desc(smem_u32(smem_b + static_cast<size_t>(stage) * p.stage_bytes_b), b_lbo, 1024);
#pragma unroll
                    for (uint32_t k = 0; k < kBlockK / 16; ++k) {
                        umma_f16(tmem_d, da + k * a_kstep, db + k * b_kstep, idesc, (kb > kb0 || k > 0) ? 1u : 0u);
                        if (p.mh == 2)  // second 128-row half: same B tile, A half 1 (16 KB further), accumulator half 1
                            umma_f16(tmem_d + p.acc_half_cols, da + ((kBlockM * 128) >> 4) + k * a_kstep, db + k * b_kstep, idesc,
                                     (kb > kb0 || k > 0) ? 1u : 0u);
                        if (rowsum_tile) {
                            umma_f16(tmem_d + p.rowsum_col, da + k * a_kstep, d_ones, idesc_ones, (kb > kb0 || k > 0) ? 1u : 0u);
                            if (p.mh == 2)
                                umma_f16(tmem_d + p.acc_half_cols + p.rowsum_col, da + ((kBlockM * 128) >> 4) + k * a_kstep, d_ones, idesc_ones,
                                         (kb > kb0 || k > 0) ? 1u : 0u);
                        }
                    }
                    umma_commit(&empty_bar[stage]);  // frees the smem stage once these MMAs have read it
                    if (++stage == static_cast<uint32_t>(S)) {
                        stage = 0;
                        phase ^= 1;
                    }
                }
                umma_commit(&acc_full[as]);  // accumulator complete -> epilogue
            }
        }
    } else {
        // ------------------------------------------------------------------ epilogue (warps 2..9)
        uint8_t* stg_base = reinterpret_cast<uint8_t*>(tmem_slot + 4);
        if (p.out_mode == OUT_BF16) run_epilogue<2>(p, warp, lane, tmem_base, acc_full, acc_empty, stg_base);
        else run_epilogue<4>(p, warp, lane, tmem_base, acc_full, acc_empty, stg_base);
    }

    tc_fence_before();
    __syncthreads();
    if (warp == 1) {
        __syncwarp();
        tmem_dealloc(tmem_base, kTmemCols);
    }
}

// ---------------------------------------------------------------------------------------------- host side

static int g_sm_count = 0;
int device_sm_count() {
    if (g_sm_count == 0) {
        int dev = 0;
        cudaGetDevice(&dev);
        cudaDeviceGetAttribute(&g_sm_count, cudaDevAttrMultiProcessorCount, dev);
        if (g_sm_count <= 0) g_sm_count = 148;
    }
    return g_sm_count;
}

typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*,
                                  const cuuint64_t*, const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave,
                                  CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

static EncodeTiledFn get_encode_fn() {
    static EncodeTiledFn fn = nullptr;
    static std::once_flag once;
    std::call_once(once, [] {
        void* sym = nullptr;
        cudaDriverEntryPointQueryResult qres;
        if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &sym, cudaEnableDefault, &qres) == cudaSuccess &&
            qres == cudaDriverEntryPointSuccess)
            fn = reinterpret_cast<EncodeTiledFn>(sym);
    });
    return fn;
}

int encode_tmap_bf16(CUtensorMap* map, const void* base, int rank, const uint64_t* dims, const uint64_t* strides,
                     const uint32_t* box, const uint32_t* elem_strides) {
    EncodeTiledFn fn = get_encode_fn();
    if (!fn) return -1;
    // cuTensorMapEncodeTiled is a driver-API call: it needs a context bound to the calling thread.  Autograd worker
    // threads may reach this point before any runtime-API call has bound the primary context, so bind it once.
    static thread_local bool ctx_bound = false;
    if (!ctx_bound) {
        cudaFree(nullptr);
        ctx_bound = true;
    }
    cuuint64_t gdim[5], gstr[4];
    cuuint32_t bdim[5], estr[5];
    for (int i = 0; i < rank; ++i) {
        gdim[i] = dims[i];
        bdim[i] = box[i];
        estr[i] = elem_strides ? elem_strides[i] : 1;
    }
    for (int i = 0; i + 1 < rank; ++i) gstr[i] = strides[i];
    CUresult r = fn(map, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, static_cast<cuuint32_t>(rank), const_cast<void*>(base), gdim,
                    gstr, bdim, estr, CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B,
                    CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    return r == CUDA_SUCCESS ? 0 : -static_cast<int>(r) - 1000;
}

template <bool A_MN, bool B_MN>
static int launch_impl(const GemmParams& p, cudaStream_t stream) {
    static std::once_flag attr_once;   // forward runs on the Python thread, backward on autograd worker threads
    static cudaError_t attr_err = cudaSuccess;
    const size_t smem = static_cast<size_t>(p.num_stages) * (p.stage_bytes_a + p.stage_bytes_b) + 1024 /*align*/ + 256 /*barriers*/ +
                        kEpilogueStagingBytes + ((p.flags & EPI_ROWSUM_A) ? kOnesTileBytes : 0);
    auto kern = gemm_tc_kernel<A_MN, B_MN>;
    std::call_once(attr_once, [&] { attr_err = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, 232448); });
    if (attr_err != cudaSuccess) return static_cast<int>(attr_err);
    int grid = p.num_tiles < device_sm_count() ? p.num_tiles : device_sm_count();
    if (grid < 1) return 0;
    return static_cast<int>(launch_pdl(kern, dim3(grid), dim3(kNumThreads), smem, stream, p));
}

int launch_gemm(const GemmParams& p, bool a_mn, bool b_mn, cudaStream_t stream) {
    if (!a_mn && !b_mn) return launch_impl<false, false>(p, stream);
    if (!a_mn && b_mn) return launch_impl<false, true>(p, stream);
    if (a_mn && b_mn) return launch_impl<true, true>(p, stream);
    return launch_impl<true, false>(p, stream);
}

}  // namespace t2v
