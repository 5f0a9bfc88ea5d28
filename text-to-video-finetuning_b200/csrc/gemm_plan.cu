// Host-side planners: turn convolution / batched-GEMM problems into coefficient tables for the affine TMA GEMM
// (gemm_tc.cuh) and export them through the C ABI declared in include/t2v_b200.h.
#include "common.h"
#include "gemm_tc.cuh"

#include "ptx.cuh"

#include <algorithm>
#include <cstdlib>
#include <cstring>
#include <cuda_bf16.h>

using namespace t2v;

namespace t2v {
int launch_channel_stats(const void* x, float* stats, int S, int64_t P, int C, int64_t ld, cudaStream_t st);  // norms.cu
}

namespace {

struct Box3 {
    int w, h, n;
};

// Factor `prod` (a power of two) into (w,h,n) box extents over a (W,H,N) pixel space, maximising useful coverage.
Box3 choose_pixel_box(int prod, int W, int H, int N) {
    Box3 best{prod, 1, 1};
    double best_eff = -1.0;
    for (int bw = prod; bw >= 1; bw >>= 1) {
        for (int bh = prod / bw; bh >= 1; bh >>= 1) {
            const int bn = prod / (bw * bh);
            if (bw > 256 || bh > 256 || bn > 256) continue;
            const double cover = double((W + bw - 1) / bw) * bw * double((H + bh - 1) / bh) * bh * double((N + bn - 1) / bn) * bn;
            const double eff = double(W) * H * N / cover;
            if (eff > best_eff + 1e-9) {
                best_eff = eff;
                best = Box3{bw, bh, bn};
            }
        }
    }
    return best;
}

// UMMA N for a problem with `row_tiles` row tiles and `ncols` columns: fewest waves, then least padding.
int env_int(const char* name) {
    const char* v = std::getenv(name);
    return v ? std::atoi(v) : 0;
}

struct Tiling {
    int bn;        // UMMA N
    int mh;        // 128-row halves per tile (1 or 2)
    int pair_var;  // tile variable (1..3) the halves pair up, -1 if mh == 1
};

// Picks (block_n, row halves) with a small roofline model of the kernel.  Measured on B200 (profiles/): the main loop is
// bound by the chip-wide L2->SM operand bandwidth (~5200 B/clk over all active CTAs) before it is bound by tcgen05
// (2*bn clk per 64-deep k-block and 128-row half), so wide tiles and 256-row tiles (B shared by two accumulators) win
// whenever enough tiles remain to fill the 148 SMs.
//   row_dims[0..2] = tile counts of tile variables 1..3 (unpaired), extra = product of the remaining tile variables.
//   splits_out != nullptr: the reduction may also be split over `s` CTAs per tile (fp32 red.add into scratch + a finishing
//   pass, see launch_split) - chosen jointly, otherwise few-tile problems would be pushed to tiny, operand-hungry tiles.
Tiling choose_tiling(const int row_dims[3], int ncols, bool mn_major_b, int kblocks, int64_t extra, bool allow_pair,
                     int* splits_out = nullptr) {
    const int sms = device_sm_count();
    const int forced_bn = env_int("T2V_FORCE_BN"), forced_mh = env_int("T2V_FORCE_MH");
    const int forced_s = splits_out ? env_int("T2V_FORCE_FWD_SPLITS") : 0;   // sweep hook (tools/plan_sweep.py)
    Tiling best{16, 1, -1};
    double best_cost = 1e30;
    for (int mh = 1; mh <= 2; ++mh) {
        if (forced_mh && mh != forced_mh) continue;
        int pv = -1;
        int64_t m_tiles = int64_t(row_dims[0]) * row_dims[1] * row_dims[2];
        if (mh == 2) {
            if (!allow_pair) continue;
            double best_waste = 1e30;
            for (int v = 2; v >= 0; --v) {
                if (row_dims[v] < 2) continue;
                const double waste = double((row_dims[v] + 1) / 2 * 2) / row_dims[v];
                if (waste < best_waste - 1e-9) {
                    best_waste = waste;
                    pv = v + 1;
                }
            }
            if (pv < 0) continue;
            m_tiles = m_tiles / row_dims[pv - 1] * ((row_dims[pv - 1] + 1) / 2);
        }
        for (int bn = 256; bn >= 16; bn -= 16) {
            if (forced_bn && bn != forced_bn) continue;
            if (bn > 16 && bn - 16 >= ncols) continue;  // strictly more padding than needed
            const int b_bytes = mn_major_b ? ((bn + 63) / 64) * 8192 : bn * 128;
            const int stage_bytes = mh * kBlockM * 128 + b_bytes;
            const int budget = 232448 - 1024 - 256 - kEpilogueStagingBytes;
            if (budget / stage_bytes < 3) continue;
            const int64_t base_tiles = m_tiles * ((ncols + bn - 1) / bn) * extra;
            int max_s = (splits_out && base_tiles * 2 <= sms && kblocks >= 8) ? std::min(kblocks / 4, 64) : 1;
            if (forced_s) max_s = std::min(forced_s, kblocks);
            for (int s = 1; s <= max_s; ++s) {
                const int kper = (kblocks + s - 1) / s;
                if (forced_s ? s != max_s : (kblocks + kper - 1) / kper != s) continue;  // same schedule as a smaller s
                const int64_t tiles = base_tiles * s;
                const int64_t waves = (tiles + sms - 1) / sms;
                const double active = double(std::min<int64_t>(tiles, sms));
                // three ceilings per k-block: tcgen05 issue rate, chip-wide L2->SM bandwidth shared by the active CTAs,
                // and the per-SM shared-memory fill rate (~48 B/clk measured)
                const double t_kb = std::max({double(mh) * 2.0 * bn, stage_bytes * active / 5200.0, stage_bytes / 48.0, 260.0});
                const bool overlap = mh == 1 || bn <= 128;  // two TMEM accumulator stages available
                const double t_epi = (bn / 32.0 + 1.0) * 450.0 * (mh == 2 ? 1.0 : 0.5) * (overlap ? 0.35 : 1.0);
                double cost = double(waves) * (kper * t_kb + t_epi + 1800.0);
                // split: red.add traffic (~2500 B/clk chip-wide) + the memset and finishing launches
                if (s > 1) cost += double(tiles) * mh * kBlockM * bn * 4.0 / 2500.0 + 10000.0;
                if (cost < best_cost - 1e-9) {
                    best_cost = cost;
                    best = Tiling{bn, mh, pv};
                    if (splits_out) *splits_out = s;
                }
            }
        }
    }
    return best;
}

void apply_tiling(GemmParams& p, const Tiling& t) {
    p.block_n = t.bn;
    p.mh = t.mh;
    p.pair_var = t.pair_var;
    if (t.mh == 2) {
        p.tdim[t.pair_var] = (p.tdim[t.pair_var] + 1) / 2;
        p.acc_half_cols = t.bn <= 128 ? 128 : 256;
        p.acc_stage_cols = 2 * p.acc_half_cols;
    } else {
        p.acc_half_cols = 0;
        p.acc_stage_cols = 256;
    }
    p.nacc = 512 / p.acc_stage_cols;
}

// Split-K factor for accumulate-mode problems: minimise  waves(base_tiles * s) * (k-blocks per split * T_kblock + T_epilogue).
int choose_splits(int64_t base_tiles, int kb_total) {
    const int sms = device_sm_count();
    if (const int forced = env_int("T2V_FORCE_SPLITS")) return std::max(1, std::min(forced, kb_total));
    int best = 1;
    double best_cost = 1e30;
    for (int s = 1; s <= kb_total && s <= 128; ++s) {
        const int64_t waves = (base_tiles * s + sms - 1) / sms;
        const int kper = (kb_total + s - 1) / s;
        const double cost = double(waves) * (kper * 320.0 + 2500.0);
        if (cost < best_cost - 1e-9) {
            best_cost = cost;
            best = s;
        }
    }
    return best;
}

void finish_common(GemmParams& p, bool b_mn) {
    p.stage_bytes_a = p.mh * kBlockM * 128;
    p.stage_bytes_b = b_mn ? ((p.block_n + 63) / 64) * 8192 : p.block_n * 128;
    const int budget = 232448 - 1024 - 256 - kEpilogueStagingBytes;
    p.num_stages = std::min<int>(kMaxStages, budget / (p.stage_bytes_a + p.stage_bytes_b));
    if (const int forced = env_int("T2V_FORCE_STAGES")) p.num_stages = std::min(p.num_stages, forced);
    int64_t tiles = 1;
    for (int i = 0; i < 6; ++i) tiles *= p.tdim[i];
    p.num_tiles = static_cast<int32_t>(tiles);
    for (int i = 0; i < 6; ++i) {  // q = umulhi(n, mul) >> shr for 0 <= n < 2^31 (CUTLASS FastDivmod construction)
        const uint32_t d = static_cast<uint32_t>(p.tdim[i]);
        if (d <= 1) {
            p.tdiv_mul[i] = 0;
            p.tdiv_shr[i] = 0;
            continue;
        }
        uint32_t lg = 0;
        while ((1ull << lg) < d) ++lg;
        const uint32_t pw = 31 + lg;
        p.tdiv_mul[i] = static_cast<uint32_t>(((1ull << pw) + d - 1) / d);
        p.tdiv_shr[i] = pw - 32;
    }
    p.kb_total = p.kdim[0] * p.kdim[1] * p.kdim[2];
}

bool aligned16(const void* p) { return (reinterpret_cast<uintptr_t>(p) & 15u) == 0; }

void fill_epilogue(GemmParams& p, const T2VEpilogue* e, void* out, int out_mode_default) {
    p.out = out;
    p.alpha = e ? e->alpha : 1.0f;
    p.bias = e ? e->bias : nullptr;
    p.rowbias = e ? e->rowbias : nullptr;
    p.residual = e ? e->residual : nullptr;
    p.out_mode = e ? (e->out_fp32 ? OUT_F32 : OUT_BF16) : out_mode_default;
    p.rb_div = (e && e->rowbias_div > 0) ? e->rowbias_div : 1;
    p.flags = 0;
    if (p.bias) p.flags |= EPI_BIAS;
    if (p.rowbias) p.flags |= EPI_ROWBIAS;
    if (p.residual) p.flags |= EPI_RESIDUAL;
}

void set_vec_flag(GemmParams& p) {
    bool ok = aligned16(p.out) && (!p.residual || aligned16(p.residual)) && (!p.bias || aligned16(p.bias)) &&
              (!p.rowbias || aligned16(p.rowbias));
    const int64_t strides[] = {p.ldw, p.ldh, p.ldn, p.otc[0], p.otc[1], p.otc[2], p.otc[3], p.otc[4], p.otc[5], p.rb_ld};
    for (int64_t s : strides) ok = ok && (s % 8 == 0);
    if (ok) p.flags |= EPI_VEC;
}

int check_channels(int c, const char* what) {
    if (c <= 0 || c % 8 != 0) return fail(-2, "%s=%d must be a positive multiple of 8 (16-byte TMA rows)", what, c);
    return 0;
}


// ---- split-K for forward / data-gradient problems with few output tiles (the 4x4 / 8x8 / 16x16 levels of the UNet:
// 16-40 tiles of up to 180 k-blocks each would leave most of the 148 SMs idle).  The k-range is split over tile variable
// 4, partial tiles are reduced with red.global.add.f32 into an fp32 scratch image of the output, and one elementwise
// kernel applies alpha / bias / row bias / residual and writes the bf16 (or fp32) result.
void apply_fwd_splits(GemmParams& p, int splits, int kb_total) {
    p.kb_per_split = (kb_total + splits - 1) / splits;
    p.tdim[4] = (kb_total + p.kb_per_split - 1) / p.kb_per_split;
    p.ksplit_var = 4;
}

// GroupNorm statistics in the epilogue (T2VEpilogue.stats): express "frame of an output row" in the tile's row coordinates
// and check that aligned runs of 32 (or 16) accumulator rows never straddle two frames.  Returns false when this tiling
// cannot do it (the caller then runs the standalone statistics pass over the finished output).
bool plan_epilogue_stats(GemmParams& p, const T2VEpilogue* e, int Wo, int Ho, int N) {
    if (!e || !e->stats || e->stats_rows <= 0) return false;
    const int64_t pf = e->stats_rows;     // output rows (flattened [N][Ho][Wo]) per statistics sample
    const int64_t frame = int64_t(Wo) * Ho;
    int cw = 0, ch = 0, cn = 0, div = 1;
    // does every aligned run of `seg` accumulator rows (tile row order: w fastest, then h, then n) share a sample?
    auto pick_seg = [&](auto&& ok) { return ok(32) ? 32 : (ok(16) ? 16 : 0); };
    int seg = 0;
    if (pf % frame == 0) {                // a sample is k whole images n (k = 1: per frame; k = F: per clip)
        const int64_t k = pf / frame, fp = int64_t(p.bw) * p.bh;
        cn = 1; div = int(k);
        seg = pick_seg([&](int sg) {
            if (fp % sg == 0) return true;                       // the run stays inside one image
            if (sg % fp) return false;
            const int64_t m = sg / fp;                           // the run covers m consecutive images of the tile
            return k % m == 0 && p.bn % m == 0;
        });
    } else if (pf % Wo == 0 && Ho % (pf / Wo) == 0) {   // a sample is k h-lines (temporal layout [B][F][HW]: h = frame index)
        const int64_t k = pf / Wo;
        ch = 1; cn = Ho; div = int(k);
        seg = pick_seg([&](int sg) {
            if (p.bw % sg == 0) return true;                     // the run stays inside one h-line
            if (sg % p.bw) return false;
            const int64_t m = sg / p.bw;                         // the run covers m consecutive h-lines of the tile
            return k % m == 0 && p.bh % m == 0;
        });
    } else if (Ho == 1 && N == 1 && Wo % pf == 0) {   // token matrix: sample = row / pf
        cw = 1; div = int(pf);
        seg = (p.bh == 1 && p.bn == 1) ? pick_seg([&](int sg) { return pf % sg == 0; }) : 0;
    }
    if (!seg) return false;
    p.stats = e->stats;
    p.st_ld = e->stats_ld;
    p.st_cw = cw; p.st_ch = ch; p.st_cn = cn; p.st_div = div; p.st_seg = seg;
    p.flags |= EPI_STATS;
    return true;
}

// Finishing pass of a split-K problem that also produces the GroupNorm statistics of its output: a block owns RPB
// consecutive rows of ONE statistics sample and 256 column quads; a thread loads its quad of all RPB rows up front (independent
// loads: one memory round trip), finishes them, and adds the column sums with two red.add.v4.
template <int RPB>
__global__ void __launch_bounds__(256) splitk_finish_stats_kernel(const float* __restrict__ acc, const float* __restrict__ bias,
                                                                  const float* __restrict__ rowbias, const __nv_bfloat16* __restrict__ residual,
                                                                  void* __restrict__ out, int64_t rows, int C, int64_t rows_per_sample, int rb_div,
                                                                  int64_t rb_ld, float alpha, int out_fp32, float* __restrict__ stats,
                                                                  int64_t st_ld, int st_rows) {
    pdl_sync();
    const int64_t r0 = int64_t(blockIdx.x) * RPB;
    const int c = (blockIdx.y * blockDim.x + threadIdx.x) * 4;
    if (c >= C) return;
    float4 v[RPB];
    uint2 w[RPB];
#pragma unroll
    for (int i = 0; i < RPB; ++i) {
        const int64_t r = r0 + i;
        if (r < rows) {
            v[i] = __ldcg(reinterpret_cast<const float4*>(acc + r * C + c));
            if (residual) w[i] = __ldg(reinterpret_cast<const uint2*>(residual + r * C + c));
        }
    }
    float4 b4 = make_float4(0.f, 0.f, 0.f, 0.f);
    if (bias) b4 = __ldg(reinterpret_cast<const float4*>(bias + c));
    float s[4] = {0.f, 0.f, 0.f, 0.f}, q[4] = {0.f, 0.f, 0.f, 0.f};
#pragma unroll
    for (int i = 0; i < RPB; ++i) {
        const int64_t r = r0 + i;
        if (r >= rows) break;
        float4 y = v[i];
        y.x = y.x * alpha + b4.x; y.y = y.y * alpha + b4.y; y.z = y.z * alpha + b4.z; y.w = y.w * alpha + b4.w;
        if (rowbias) {
            const float4 b = __ldg(reinterpret_cast<const float4*>(rowbias + (r / rows_per_sample / rb_div) * rb_ld + c));
            y.x += b.x; y.y += b.y; y.z += b.z; y.w += b.w;
        }
        if (residual) {
            y.x += bf16_lo(w[i].x); y.y += bf16_hi(w[i].x); y.z += bf16_lo(w[i].y); y.w += bf16_hi(w[i].y);
        }
        if (out_fp32) {
            reinterpret_cast<float4*>(static_cast<float*>(out) + r * C)[c >> 2] = y;
        } else {
            uint2 o;
            o.x = pack_bf16(y.x, y.y);
            o.y = pack_bf16(y.z, y.w);
            reinterpret_cast<uint2*>(static_cast<__nv_bfloat16*>(out) + r * C)[c >> 2] = o;
            // the statistics describe what the consumer reads: the rounded values
            y.x = bf16_lo(o.x); y.y = bf16_hi(o.x); y.z = bf16_lo(o.y); y.w = bf16_hi(o.y);
        }
        s[0] += y.x; s[1] += y.y; s[2] += y.z; s[3] += y.w;
        q[0] += y.x * y.x; q[1] += y.y * y.y; q[2] += y.z * y.z; q[3] += y.w * y.w;
    }
    float* sp = stats + ((r0 / st_rows) * st_ld + c) * 2;
    red_add_f32x4(sp, s[0], q[0], s[1], q[1]);
    red_add_f32x4(sp + 4, s[2], q[2], s[3], q[3]);
}

__global__ void splitk_finish_kernel(const float* __restrict__ acc, const float* __restrict__ bias, const float* __restrict__ rowbias,
                                     const __nv_bfloat16* __restrict__ residual, void* __restrict__ out, int64_t rows, int C,
                                     int64_t rows_per_sample, int rb_div, int64_t rb_ld, float alpha, int out_fp32) {
    pdl_sync();
    const int V = C >> 2;  // 4 columns per thread
    const int64_t total = rows * V;
    for (int64_t i = blockIdx.x * int64_t(blockDim.x) + threadIdx.x; i < total; i += int64_t(gridDim.x) * blockDim.x) {
        const int64_t r = i / V;
        const int c = int(i % V) * 4;
        float4 v = __ldcg(reinterpret_cast<const float4*>(acc + r * C + c));
        v.x *= alpha; v.y *= alpha; v.z *= alpha; v.w *= alpha;
        if (bias) {
            const float4 b = __ldg(reinterpret_cast<const float4*>(bias + c));
            v.x += b.x; v.y += b.y; v.z += b.z; v.w += b.w;
        }
        if (rowbias) {
            const float4 b = __ldg(reinterpret_cast<const float4*>(rowbias + (r / rows_per_sample / rb_div) * rb_ld + c));
            v.x += b.x; v.y += b.y; v.z += b.z; v.w += b.w;
        }
        if (residual) {
            const uint2 q = __ldg(reinterpret_cast<const uint2*>(residual + r * C + c));
            v.x += bf16_lo(q.x); v.y += bf16_hi(q.x); v.z += bf16_lo(q.y); v.w += bf16_hi(q.y);
        }
        if (out_fp32) {
            reinterpret_cast<float4*>(static_cast<float*>(out) + r * C)[c >> 2] = v;
        } else {
            uint2 q;
            q.x = pack_bf16(v.x, v.y);
            q.y = pack_bf16(v.z, v.w);
            reinterpret_cast<uint2*>(static_cast<__nv_bfloat16*>(out) + r * C)[c >> 2] = q;
        }
    }
}

// Runs the planned problem split over k into `ws` and finishes into the real output.  `e` is the caller's epilogue.
int launch_split(GemmParams& p, bool a_mn, bool b_mn, const T2VEpilogue& e, void* out, int64_t rows, int C, int64_t rows_per_sample,
                 cudaStream_t st, const char* what) {
    const size_t bytes = size_t(rows) * C * sizeof(float);
    if (!e.workspace || size_t(e.workspace_bytes) < bytes) return fail(-3, "%s: split-K needs %zu bytes of workspace", what, bytes);
    cudaMemsetAsync(e.workspace, 0, bytes, st);
    p.out = e.workspace;
    p.out_mode = OUT_F32_RED;
    p.alpha = 1.0f;
    p.bias = nullptr; p.rowbias = nullptr; p.residual = nullptr;
    p.flags = 0;
    set_vec_flag(p);
    if (int r = launch_checked(launch_gemm(p, a_mn, b_mn, st), what)) return r;
    if (e.stats && e.stats_rows > 0 && e.stats_ld % 2 == 0 && (reinterpret_cast<uintptr_t>(e.stats) & 15u) == 0) {
        // rows per block: a divisor of the sample's rows, so a block never straddles two samples
        const int rpb = e.stats_rows % 4 == 0 ? 4 : (e.stats_rows % 2 == 0 ? 2 : 1);
        const dim3 grid(unsigned((rows + rpb - 1) / rpb), unsigned((C / 4 + 255) / 256));
        const dim3 block(unsigned(std::min(256, (C / 4 + 31) / 32 * 32)));
        auto go = [&](auto kern) {
            return int(launch_pdl(kern, grid, block, 0, st, static_cast<const float*>(e.workspace), e.bias, e.rowbias,
                                  static_cast<const __nv_bfloat16*>(e.residual), out, rows, C, rows_per_sample,
                                  e.rowbias_div > 0 ? e.rowbias_div : 1, int64_t(C), e.alpha, e.out_fp32, e.stats, e.stats_ld, e.stats_rows));
        };
        const int rc = rpb == 4 ? go(splitk_finish_stats_kernel<4>) : (rpb == 2 ? go(splitk_finish_stats_kernel<2>) : go(splitk_finish_stats_kernel<1>));
        return launch_checked(rc, what);
    }
    const int64_t vec = rows * (C / 4);
    const int grid = int(std::min<int64_t>((vec + 255) / 256, 148 * 8));
    const int rc = int(launch_pdl(splitk_finish_kernel, dim3(grid), dim3(256), 0, st, static_cast<const float*>(e.workspace), e.bias,
                                  e.rowbias, static_cast<const __nv_bfloat16*>(e.residual), out, rows, C, rows_per_sample,
                                  e.rowbias_div > 0 ? e.rowbias_div : 1, int64_t(C), e.alpha, e.out_fp32));
    return launch_checked(rc, what);
}

}  // namespace

extern "C" {

static int conv_fwd_impl(const void* x, const void* w, void* y, int32_t N, int32_t H, int32_t W, int32_t Cin, int32_t Cout,
                         int32_t KH, int32_t KW, int32_t stride, int32_t pad_h0, int32_t pad_h1, int32_t pad_w0, int32_t pad_w1,
                         const T2VEpilogue* epi, void* stream, bool plan_only) {
    if (int r = check_channels(Cin, "Cin")) return r;
    if (Cout <= 0 || N <= 0 || H <= 0 || W <= 0) return fail(-2, "conv_fwd: bad shape");
    if (stride != 1 && stride != 2) return fail(-2, "conv_fwd: stride %d unsupported", stride);
    const int Ho = (H + pad_h0 + pad_h1 - KH) / stride + 1, Wo = (W + pad_w0 + pad_w1 - KW) / stride + 1;
    if (Ho <= 0 || Wo <= 0) return fail(-2, "conv_fwd: empty output");
    GemmParams p;
    std::memset(&p, 0, sizeof(p));
    const Box3 bx = choose_pixel_box(kBlockM, Wo, Ho, N);
    p.tdim[1] = (Wo + bx.w - 1) / bx.w;
    p.tdim[2] = (Ho + bx.h - 1) / bx.h;
    p.tdim[3] = (N + bx.n - 1) / bx.n;
    p.tdim[4] = p.tdim[5] = 1;
    p.kdim[0] = (Cin + kBlockK - 1) / kBlockK;
    p.kdim[1] = KW;
    p.kdim[2] = KH;
    p.ksplit_var = -1;
    int splits = 1;
    {
        const int rd[3] = {p.tdim[1], p.tdim[2], p.tdim[3]};
        const bool may_split = (plan_only || (epi && epi->workspace)) && Cout % 8 == 0 && !env_int("T2V_NO_SPLIT");
        apply_tiling(p, choose_tiling(rd, Cout, false, p.kdim[0] * KW * KH, 1, true, may_split ? &splits : nullptr));
    }
    if (plan_only) return splits;
    p.tdim[0] = (Cout + p.block_n - 1) / p.block_n;
    // A: activations, K-major pixel box with tap shifts
    {
        TmaOperand& a = p.a;
        const uint64_t dims[4] = {uint64_t(Cin), uint64_t(W), uint64_t(H), uint64_t(N)};
        const uint64_t str[3] = {uint64_t(Cin) * 2, uint64_t(W) * Cin * 2, uint64_t(H) * W * Cin * 2};
        const uint32_t box[4] = {64, uint32_t(bx.w * stride), uint32_t(bx.h * stride), uint32_t(bx.n)};
        const uint32_t est[4] = {1, uint32_t(stride), uint32_t(stride), 1};
        if (!plan_only)
            if (int r = encode_tmap_bf16(&a.map, x, 4, dims, str, box, est)) return fail(r, "conv_fwd: A tensor map (%d)", r);
        a.rank = 4; a.nbox = 1; a.box_dim = 0; a.box_step = 0; a.box_bytes = bx.w * bx.h * bx.n * 128;
        a.base[1] = -pad_w0; a.base[2] = -pad_h0;
        a.tcoef[1][1] = bx.w * stride; a.tcoef[2][2] = bx.h * stride; a.tcoef[3][3] = bx.n;
        a.kcoef[0][0] = kBlockK; a.kcoef[1][1] = 1; a.kcoef[2][2] = 1;
    }
    // B: weights [Cout][KH*KW*Cin], K-major rows
    {
        TmaOperand& b = p.b;
        const uint64_t kt = uint64_t(KH) * KW * Cin;
        const uint64_t dims[2] = {kt, uint64_t(Cout)};
        const uint64_t str[1] = {kt * 2};
        const uint32_t box[2] = {64, uint32_t(p.block_n)};
        if (!plan_only)
            if (int r = encode_tmap_bf16(&b.map, w, 2, dims, str, box, nullptr)) return fail(r, "conv_fwd: B tensor map (%d)", r);
        b.rank = 2; b.nbox = 1; b.box_bytes = p.block_n * 128;
        b.tcoef[1][0] = p.block_n;
        b.kcoef[0][0] = kBlockK; b.kcoef[0][1] = Cin; b.kcoef[0][2] = KW * Cin;
    }
    finish_common(p, false);
    p.bw = bx.w; p.bh = bx.h; p.bn = bx.n;
    p.W = Wo; p.H = Ho; p.N = N; p.ncols = Cout;
    p.ldw = Cout; p.ldh = int64_t(Wo) * Cout; p.ldn = int64_t(Ho) * Wo * Cout;
    p.rb_ld = Cout;
    fill_epilogue(p, epi, y, OUT_BF16);
    const int kb_total = p.kdim[0] * p.kdim[1] * p.kdim[2];
    if (splits > 1) {
        apply_fwd_splits(p, splits, kb_total);
        finish_common(p, false);
        return launch_split(p, false, false, *epi, y, int64_t(N) * Ho * Wo, Cout, int64_t(Ho) * Wo, static_cast<cudaStream_t>(stream), "conv_fwd");
    }
    set_vec_flag(p);
    bool stats_fallback = false;
    if (epi && epi->stats) {
        if (!(p.flags & EPI_VEC) || epi->out_fp32 || !plan_epilogue_stats(p, epi, Wo, Ho, N)) stats_fallback = true;
    }
    if (int r = launch_checked(launch_gemm(p, false, false, static_cast<cudaStream_t>(stream)), "conv_fwd")) return r;
    if (stats_fallback) {   // this tiling cannot produce the statistics in the epilogue: one extra read of the (L2-hot) output
        if (epi->out_fp32 || epi->stats_rows <= 0 || (int64_t(N) * Ho * Wo) % epi->stats_rows)
            return fail(-2, "conv_fwd: statistics requested for an unsupported output (fp32 or ragged frames)");
        const int frames = int(int64_t(N) * Ho * Wo / epi->stats_rows);
        return launch_checked(launch_channel_stats(y, epi->stats, frames, epi->stats_rows, Cout, epi->stats_ld, static_cast<cudaStream_t>(stream)),
                              "conv_fwd(stats)");
    }
    return 0;
}

static int conv_dgrad_impl(const void* dy, const void* w, void* dx, int32_t N, int32_t H, int32_t W, int32_t Cin, int32_t Cout,
                           int32_t KH, int32_t KW, int32_t stride, int32_t pad_h0, int32_t pad_h1, int32_t pad_w0,
                           int32_t pad_w1, const T2VEpilogue* epi, void* stream, bool plan_only) {
    if (int r = check_channels(Cin, "Cin")) return r;
    if (int r = check_channels(Cout, "Cout")) return r;
    if (stride != 1 && stride != 2) return fail(-2, "conv_dgrad: stride %d unsupported", stride);
    const int Ho = (H + pad_h0 + pad_h1 - KH) / stride + 1, Wo = (W + pad_w0 + pad_w1 - KW) / stride + 1;
    const int s = stride;
    // One launch per output parity class (a single class when stride == 1).
    for (int ph = 0; ph < s; ++ph) {
        for (int pw = 0; pw < s; ++pw) {
            const int Hc = (H - ph + s - 1) / s, Wc = (W - pw + s - 1) / s;  // outputs in this class
            if (Hc <= 0 || Wc <= 0) continue;
            const int th0 = (ph + pad_h0) % s, tw0 = (pw + pad_w0) % s;      // first contributing tap
            const int nth = th0 < KH ? (KH - th0 + s - 1) / s : 0, ntw = tw0 < KW ? (KW - tw0 + s - 1) / s : 0;
            if (nth == 0 || ntw == 0) return fail(-2, "conv_dgrad: parity class without taps is unsupported");
            GemmParams p;
            std::memset(&p, 0, sizeof(p));
            const Box3 bx = choose_pixel_box(kBlockM, Wc, Hc, N);
            p.tdim[1] = (Wc + bx.w - 1) / bx.w;
            p.tdim[2] = (Hc + bx.h - 1) / bx.h;
            p.tdim[3] = (N + bx.n - 1) / bx.n;
            p.tdim[4] = p.tdim[5] = 1;
            p.kdim[0] = (Cout + kBlockK - 1) / kBlockK;
            p.kdim[1] = ntw;
            p.kdim[2] = nth;
            p.ksplit_var = -1;
            int splits = 1;
            {
                const int rd[3] = {p.tdim[1], p.tdim[2], p.tdim[3]};
                // split-K only for the single-class (stride 1) problem: the scratch image is the whole dx
                const bool may_split = s == 1 && (plan_only || (epi && epi->workspace)) && !env_int("T2V_NO_SPLIT");
                apply_tiling(p, choose_tiling(rd, Cin, true, p.kdim[0] * ntw * nth, 1, true, may_split ? &splits : nullptr));
            }
            if (plan_only) return splits;
            p.tdim[0] = (Cin + p.block_n - 1) / p.block_n;
            {
                TmaOperand& a = p.a;  // dy, K-major (K = Cout), pixel box shifted against the tap
                const uint64_t dims[4] = {uint64_t(Cout), uint64_t(Wo), uint64_t(Ho), uint64_t(N)};
                const uint64_t str[3] = {uint64_t(Cout) * 2, uint64_t(Wo) * Cout * 2, uint64_t(Ho) * Wo * Cout * 2};
                const uint32_t box[4] = {64, uint32_t(bx.w), uint32_t(bx.h), uint32_t(bx.n)};
                if (!plan_only)
                    if (int r = encode_tmap_bf16(&a.map, dy, 4, dims, str, box, nullptr)) return fail(r, "conv_dgrad: A tensor map (%d)", r);
                a.rank = 4; a.nbox = 1; a.box_bytes = bx.w * bx.h * bx.n * 128;
                a.base[1] = (pw + pad_w0 - tw0) / s; a.base[2] = (ph + pad_h0 - th0) / s;
                a.tcoef[1][1] = bx.w; a.tcoef[2][2] = bx.h; a.tcoef[3][3] = bx.n;
                a.kcoef[0][0] = kBlockK; a.kcoef[1][1] = -1; a.kcoef[2][2] = -1;
            }
            {
                TmaOperand& b = p.b;  // w viewed as [Cout (k)][tap][Cin (n, contiguous)] -> MN-major
                const uint64_t dims[3] = {uint64_t(Cin), uint64_t(KH) * KW, uint64_t(Cout)};
                const uint64_t str[2] = {uint64_t(Cin) * 2, uint64_t(KH) * KW * Cin * 2};
                const uint32_t box[3] = {64, 1, 64};
                if (!plan_only)
                    if (int r = encode_tmap_bf16(&b.map, w, 3, dims, str, box, nullptr)) return fail(r, "conv_dgrad: B tensor map (%d)", r);
                b.rank = 3; b.nbox = (p.block_n + 63) / 64; b.box_dim = 0; b.box_step = 64; b.box_bytes = 8192;
                b.base[1] = th0 * KW + tw0;
                b.tcoef[0][0] = p.block_n;
                b.kcoef[1][1] = s; b.kcoef[1][2] = s * KW; b.kcoef[2][0] = kBlockK;
            }
            finish_common(p, true);
            p.bw = bx.w; p.bh = bx.h; p.bn = bx.n;
            p.W = Wc; p.H = Hc; p.N = N; p.ncols = Cin;
            p.ldw = int64_t(s) * Cin; p.ldh = int64_t(s) * W * Cin; p.ldn = int64_t(H) * W * Cin;
            p.rb_ld = Cin;
            const int64_t base_off = (int64_t(ph) * W + pw) * Cin;
            T2VEpilogue e = epi ? *epi : T2VEpilogue{nullptr, nullptr, nullptr, 1.0f, 0, 1, nullptr, 0};
            const size_t esz = e.out_fp32 ? 4 : 2;
            fill_epilogue(p, &e, static_cast<char*>(dx) + base_off * esz, OUT_BF16);
            if (p.residual) p.residual = static_cast<const char*>(p.residual) + base_off * 2;
            if (splits > 1) {
                apply_fwd_splits(p, splits, p.kdim[0] * p.kdim[1] * p.kdim[2]);
                finish_common(p, true);
                return launch_split(p, false, true, e, dx, int64_t(N) * H * W, Cin, int64_t(H) * W, static_cast<cudaStream_t>(stream), "conv_dgrad");
            }
            set_vec_flag(p);
            if (int r = launch_checked(launch_gemm(p, false, true, static_cast<cudaStream_t>(stream)), "conv_dgrad")) return r;
        }
    }
    return 0;
}

int t2v_conv_fwd(const void* x, const void* w, void* y, int32_t N, int32_t H, int32_t W, int32_t Cin, int32_t Cout,
                 int32_t KH, int32_t KW, int32_t stride, int32_t pad_h0, int32_t pad_h1, int32_t pad_w0, int32_t pad_w1,
                 const T2VEpilogue* epi, void* stream) {
    return conv_fwd_impl(x, w, y, N, H, W, Cin, Cout, KH, KW, stride, pad_h0, pad_h1, pad_w0, pad_w1, epi, stream, false);
}

int t2v_conv_dgrad(const void* dy, const void* w, void* dx, int32_t N, int32_t H, int32_t W, int32_t Cin, int32_t Cout,
                   int32_t KH, int32_t KW, int32_t stride, int32_t pad_h0, int32_t pad_h1, int32_t pad_w0,
                   int32_t pad_w1, const T2VEpilogue* epi, void* stream) {
    return conv_dgrad_impl(dy, w, dx, N, H, W, Cin, Cout, KH, KW, stride, pad_h0, pad_h1, pad_w0, pad_w1, epi, stream, false);
}

int64_t t2v_conv_workspace_bytes(int32_t dgrad, int32_t N, int32_t H, int32_t W, int32_t Cin, int32_t Cout, int32_t KH, int32_t KW,
                                 int32_t stride, int32_t pad_h0, int32_t pad_h1, int32_t pad_w0, int32_t pad_w1) {
    if (Cin <= 0 || Cout <= 0 || Cin % 8 || Cout % 8 || (stride != 1 && stride != 2)) return 0;
    if (dgrad) {
        const int splits = conv_dgrad_impl(nullptr, nullptr, nullptr, N, H, W, Cin, Cout, KH, KW, stride, pad_h0, pad_h1, pad_w0, pad_w1,
                                           nullptr, nullptr, true);
        return splits > 1 ? int64_t(N) * H * W * Cin * 4 : 0;
    }
    const int Ho = (H + pad_h0 + pad_h1 - KH) / stride + 1, Wo = (W + pad_w0 + pad_w1 - KW) / stride + 1;
    const int splits = conv_fwd_impl(nullptr, nullptr, nullptr, N, H, W, Cin, Cout, KH, KW, stride, pad_h0, pad_h1, pad_w0, pad_w1,
                                     nullptr, nullptr, true);
    return splits > 1 ? int64_t(N) * Ho * Wo * Cout * 4 : 0;
}

// dbias != NULL: the bias gradient dbias[Cout] += sum over output pixels of dy rides in the same launch whenever the tiling
// leaves 32 spare TMEM columns next to the accumulator (EPI_ROWSUM_A, gemm_tc.cuh); otherwise a column-sum pass follows.
static int conv_wgrad_impl(const void* x, const void* dy, float* dw, float* dbias, int32_t N, int32_t H, int32_t W, int32_t Cin,
                           int32_t Cout, int32_t KH, int32_t KW, int32_t stride, int32_t pad_h0, int32_t pad_h1, int32_t pad_w0,
                           int32_t pad_w1, void* stream) {
    if (int r = check_channels(Cin, "Cin")) return r;
    if (int r = check_channels(Cout, "Cout")) return r;
    if (stride != 1 && stride != 2) return fail(-2, "conv_wgrad: stride %d unsupported", stride);
    const int Ho = (H + pad_h0 + pad_h1 - KH) / stride + 1, Wo = (W + pad_w0 + pad_w1 - KW) / stride + 1;
    GemmParams p;
    std::memset(&p, 0, sizeof(p));
    const Box3 kx = choose_pixel_box(kBlockK, Wo, Ho, N);  // 64 output pixels per k-block
    p.kdim[0] = (Wo + kx.w - 1) / kx.w;
    p.kdim[1] = (Ho + kx.h - 1) / kx.h;
    p.kdim[2] = (N + kx.n - 1) / kx.n;
    const int kb_total = p.kdim[0] * p.kdim[1] * p.kdim[2];
    p.tdim[1] = (Cout + kBlockM - 1) / kBlockM;
    p.tdim[2] = KW;
    p.tdim[3] = KH;
    // Joint choice of (UMMA N, 128/256-row tiles, split-K factor).  Cost model fitted to a sweep over 14 weight-gradient
    // shapes of the ms-1.7b UNet on B200 (tools/gemm_sweep.py, profiles/r1_wgrad_sweep.txt; picks within 1% of the best
    // measured configuration on 13 of them): per k-block the slowest of tcgen05 issue, per-SM operand fill (~40 B/clk with
    // MN-major boxes) and chip-wide L2->SM bandwidth (~6000 B/clk); the split-K epilogue pays for its red.global.add
    // traffic at ~2500 B/clk chip-wide.
    struct Choice {
        Tiling t;
        int splits;
        double cost;
    };
    auto search = [&](int max_bn) {
        const int sms = device_sm_count();
        const int forced_bn = env_int("T2V_FORCE_BN"), forced_mh = env_int("T2V_FORCE_MH"), forced_s = env_int("T2V_FORCE_SPLITS");
        const int64_t taps = int64_t(KH) * KW;
        const int mt1 = p.tdim[1];
        Choice best{Tiling{16, 1, -1}, 1, 1e30};
        for (int mh = 1; mh <= 2; ++mh) {
            if ((forced_mh && mh != forced_mh) || (mh == 2 && mt1 < 2)) continue;
            const int mt = mh == 2 ? (mt1 + 1) / 2 : mt1;
            for (int bn = max_bn; bn >= 16; bn -= 16) {
                if (forced_bn && bn != forced_bn) continue;
                if (bn > 16 && bn - 16 >= Cin) continue;
                const int stage_bytes = mh * kBlockM * 128 + ((bn + 63) / 64) * 8192;
                // (the 2 KB tile of ones of the fused bias gradient is budgeted for every candidate, so that the choice with and
                // without dbias differs only by the UMMA N limit)
                const int budget = 232448 - 1024 - 256 - kEpilogueStagingBytes - kOnesTileBytes;
                if (budget / stage_bytes < 3) continue;
                const int64_t base = int64_t(mt) * ((Cin + bn - 1) / bn) * taps;
                for (int s = 1; s <= kb_total && s <= 128; ++s) {
                    if (forced_s && s != std::min(forced_s, kb_total)) continue;
                    const int kper = (kb_total + s - 1) / s;
                    if (!forced_s && (kb_total + kper - 1) / kper != s) continue;  // same schedule as a smaller s
                    const int64_t tiles = base * s;
                    const int64_t waves = (tiles + sms - 1) / sms;
                    const double active = double(std::min<int64_t>(tiles, sms));
                    const double t_kb = std::max({double(mh) * 2.0 * bn, stage_bytes / 40.0, stage_bytes * active / 6000.0, 260.0});
                    const double t_epi = 500.0 + 4.0 * mh * bn;
                    const double red = double(tiles) * mh * kBlockM * bn * 4.0 / 2500.0;
                    const double cost = double(waves) * (kper * t_kb + t_epi) + red;
                    if (cost < best.cost - 1e-9) best = Choice{Tiling{bn, mh, mh == 2 ? 1 : -1}, s, cost};
                }
            }
        }
        return best;
    };
    Choice pick = search(256);
    bool fuse_rowsum = false;
    if (dbias && !env_int("T2V_NO_ROWSUM_FUSE")) {
        // the row-sum accumulator needs the 32 TMEM columns after the tile's own: UMMA N <= 224.  A narrower tile is accepted
        // when it costs less than the separate column-sum launch it saves (launch gap + one pass over dy, ~3000 B/clk).
        const Choice narrow = pick.t.bn <= 224 ? pick : search(224);
        const double colsum_cost = 6000.0 + double(N) * Ho * Wo * Cout * 2.0 / 3000.0;
        if (narrow.cost < 1e29 && narrow.cost <= pick.cost + colsum_cost) {
            pick = narrow;
            fuse_rowsum = true;
        }
    }
    int splits = pick.splits;
    apply_tiling(p, pick.t);
    if (fuse_rowsum && p.mh == 2 && p.block_n > 96 && p.acc_half_cols == 128) {   // make room: one accumulator stage of 2 x 256 columns
        p.acc_half_cols = 256;
        p.acc_stage_cols = 512;
        p.nacc = 1;
    }
    p.tdim[0] = (Cin + p.block_n - 1) / p.block_n;
    p.kb_per_split = (kb_total + splits - 1) / splits;
    splits = (kb_total + p.kb_per_split - 1) / p.kb_per_split;
    p.tdim[4] = splits;
    p.tdim[5] = 1;
    p.ksplit_var = 4;
    {
        TmaOperand& a = p.a;  // dy^T: M = Cout (contiguous), K = pixels -> MN-major, two 64-wide boxes
        const uint64_t dims[4] = {uint64_t(Cout), uint64_t(Wo), uint64_t(Ho), uint64_t(N)};
        const uint64_t str[3] = {uint64_t(Cout) * 2, uint64_t(Wo) * Cout * 2, uint64_t(Ho) * Wo * Cout * 2};
        const uint32_t box[4] = {64, uint32_t(kx.w), uint32_t(kx.h), uint32_t(kx.n)};
        if (int r = encode_tmap_bf16(&a.map, dy, 4, dims, str, box, nullptr)) return fail(r, "conv_wgrad: A tensor map (%d)", r);
        a.rank = 4; a.nbox = 2; a.box_dim = 0; a.box_step = 64; a.box_bytes = 8192;
        a.tcoef[0][1] = kBlockM;
        a.kcoef[1][0] = kx.w; a.kcoef[2][1] = kx.h; a.kcoef[3][2] = kx.n;
    }
    {
        TmaOperand& b = p.b;  // x shifted by the tap: N = Cin (contiguous), K = pixels -> MN-major
        const uint64_t dims[4] = {uint64_t(Cin), uint64_t(W), uint64_t(H), uint64_t(N)};
        const uint64_t str[3] = {uint64_t(Cin) * 2, uint64_t(W) * Cin * 2, uint64_t(H) * W * Cin * 2};
        const uint32_t box[4] = {64, uint32_t(kx.w * stride), uint32_t(kx.h * stride), uint32_t(kx.n)};
        const uint32_t est[4] = {1, uint32_t(stride), uint32_t(stride), 1};
        if (int r = encode_tmap_bf16(&b.map, x, 4, dims, str, box, est)) return fail(r, "conv_wgrad: B tensor map (%d)", r);
        b.rank = 4; b.nbox = (p.block_n + 63) / 64; b.box_dim = 0; b.box_step = 64; b.box_bytes = 8192;
        b.base[1] = -pad_w0; b.base[2] = -pad_h0;
        b.tcoef[0][0] = p.block_n; b.tcoef[1][2] = 1; b.tcoef[2][3] = 1;
        b.kcoef[1][0] = kx.w * stride; b.kcoef[2][1] = kx.h * stride; b.kcoef[3][2] = kx.n;
    }
    finish_common(p, true);
    p.bw = kBlockM; p.bh = 1; p.bn = 1;
    p.W = Cout; p.H = KW; p.N = KH; p.ncols = Cin;
    p.ldw = int64_t(KH) * KW * Cin; p.ldh = Cin; p.ldn = int64_t(KW) * Cin;
    fill_epilogue(p, nullptr, dw, OUT_F32_RED);
    p.alpha = 1.0f;
    set_vec_flag(p);
    if (fuse_rowsum) {
        p.flags |= EPI_ROWSUM_A;
        p.rowsum = dbias;
        p.rowsum_col = static_cast<uint32_t>(p.block_n);
        const int budget = 232448 - 1024 - 256 - kEpilogueStagingBytes - kOnesTileBytes;
        p.num_stages = std::min<int>(p.num_stages, budget / (p.stage_bytes_a + p.stage_bytes_b));
    }
    if (int r = launch_checked(launch_gemm(p, true, true, static_cast<cudaStream_t>(stream)), "conv_wgrad")) return r;
    if (dbias && !fuse_rowsum) return t2v_colsum(dy, dbias, 1, int64_t(N) * Ho * Wo, Cout, stream);
    return 0;
}

int t2v_conv_wgrad(const void* x, const void* dy, float* dw, int32_t N, int32_t H, int32_t W, int32_t Cin, int32_t Cout,
                   int32_t KH, int32_t KW, int32_t stride, int32_t pad_h0, int32_t pad_h1, int32_t pad_w0,
                   int32_t pad_w1, void* stream) {
    return conv_wgrad_impl(x, dy, dw, nullptr, N, H, W, Cin, Cout, KH, KW, stride, pad_h0, pad_h1, pad_w0, pad_w1, stream);
}

int t2v_conv_wgrad_bias(const void* x, const void* dy, float* dw, float* dbias, int32_t N, int32_t H, int32_t W, int32_t Cin,
                        int32_t Cout, int32_t KH, int32_t KW, int32_t stride, int32_t pad_h0, int32_t pad_h1, int32_t pad_w0,
                        int32_t pad_w1, void* stream) {
    return conv_wgrad_impl(x, dy, dw, dbias, N, H, W, Cin, Cout, KH, KW, stride, pad_h0, pad_h1, pad_w0, pad_w1, stream);
}

int t2v_bgemm(const T2VMat* A, const T2VMat* B, void* C, int64_t ldc, int64_t c_stride_z1, int64_t c_stride_z2,
              int32_t M, int32_t N, int32_t K, int32_t Z1, int32_t Z2, float alpha, int32_t out_mode, void* stream) {
    if (M <= 0 || N <= 0 || K <= 0 || Z1 <= 0 || Z2 <= 0) return fail(-2, "bgemm: bad shape");
    if (A->ld % 8 || B->ld % 8) return fail(-2, "bgemm: leading dimensions must be multiples of 8 elements");
    if ((Z1 > 1 && (A->stride_z1 % 8 || B->stride_z1 % 8)) || (Z2 > 1 && (A->stride_z2 % 8 || B->stride_z2 % 8)))
        return fail(-2, "bgemm: batch strides must be multiples of 8 elements");
    const bool a_mn = !A->kmajor, b_mn = !B->kmajor;
    GemmParams p;
    std::memset(&p, 0, sizeof(p));
    p.kdim[0] = (K + kBlockK - 1) / kBlockK;
    p.kdim[1] = p.kdim[2] = 1;
    p.tdim[1] = (M + kBlockM - 1) / kBlockM;
    p.tdim[2] = 1;
    p.tdim[3] = 1;
    p.tdim[4] = Z2;
    p.tdim[5] = Z1;
    {
        const int rd[3] = {p.tdim[1], 1, 1};
        apply_tiling(p, choose_tiling(rd, N, b_mn, out_mode == OUT_F32_RED ? std::max(2, p.kdim[0] / 4) : p.kdim[0], int64_t(Z1) * Z2, true));
    }
    p.tdim[0] = (N + p.block_n - 1) / p.block_n;
    p.ksplit_var = -1;
    int splits = 1;
    if (out_mode == OUT_F32_RED) {
        const int64_t base_tiles = int64_t(p.tdim[0]) * p.tdim[1] * Z1 * Z2;
        splits = choose_splits(base_tiles, p.kdim[0]);
        p.kb_per_split = (p.kdim[0] + splits - 1) / splits;
        splits = (p.kdim[0] + p.kb_per_split - 1) / p.kb_per_split;
        p.tdim[2] = splits;
        p.ksplit_var = 2;
    }
    auto stride_or = [](int64_t s, int64_t fallback) { return uint64_t((s > 0 ? s : fallback) * 2); };
    auto plan_operand = [&](TmaOperand& op, const T2VMat* m, bool mn, int rows_mn, int block_mn, int tile_var) -> int {
        // stored matrix: K-major [rows_mn][K]; MN-major [K][rows_mn]
        const uint64_t inner = mn ? uint64_t(rows_mn) : uint64_t(K), outer = mn ? uint64_t(K) : uint64_t(rows_mn);
        const uint64_t dims[4] = {inner, outer, uint64_t(Z2), uint64_t(Z1)};
        const uint64_t str[3] = {uint64_t(m->ld) * 2, stride_or(m->stride_z2, m->ld * int64_t(outer)),
                                 stride_or(m->stride_z1, m->ld * int64_t(outer) * Z2)};
        const uint32_t box[4] = {64, uint32_t(mn ? 64 : block_mn), 1, 1};
        if (int r = encode_tmap_bf16(&op.map, m->ptr, 4, dims, str, box, nullptr)) return r;
        op.rank = 4;
        op.tcoef[2][4] = 1;
        op.tcoef[3][5] = 1;
        if (mn) {
            op.nbox = (block_mn + 63) / 64; op.box_dim = 0; op.box_step = 64; op.box_bytes = 8192;
            op.tcoef[0][tile_var] = block_mn;
            op.kcoef[1][0] = kBlockK;
        } else {
            op.nbox = 1; op.box_bytes = block_mn * 128;
            op.tcoef[1][tile_var] = block_mn;
            op.kcoef[0][0] = kBlockK;
        }
        return 0;
    };
    if (int r = plan_operand(p.a, A, a_mn, M, kBlockM, 1)) return fail(r, "bgemm: A tensor map (%d)", r);
    if (int r = plan_operand(p.b, B, b_mn, N, p.block_n, 0)) return fail(r, "bgemm: B tensor map (%d)", r);
    finish_common(p, b_mn);
    p.bw = kBlockM; p.bh = 1; p.bn = 1;
    p.W = M; p.H = splits; p.N = 1; p.ncols = N;
    p.ldw = ldc; p.ldh = 0; p.ldn = 0;
    p.otc[4] = c_stride_z2; p.otc[5] = c_stride_z1;
    T2VEpilogue e{nullptr, nullptr, nullptr, alpha, out_mode != OUT_BF16, 1};
    fill_epilogue(p, &e, C, OUT_BF16);
    p.out_mode = out_mode;
    set_vec_flag(p);
    return launch_checked(launch_gemm(p, a_mn, b_mn, static_cast<cudaStream_t>(stream)), "bgemm");
}

}  // extern "C"
