// spmm.cuh -- the gather-accumulate core shared by SpMM, sampled SpMM and the fused
// GAT layer.  Written for sm_100a; see DESIGN.md "Kernels" for the byte model.
//
// Work decomposition
//   * one warp per sparse row (8 rows per 256-thread CTA); rows whose total degree
//     exceeds plan->hub_threshold are skipped by their warp and executed by a whole
//     CTA instead (hub CTAs are the first blockIdx.x values so that the longest tasks
//     start first).  Hub partials are combined through shared memory in fixed warp
//     order: no atomics, bit-reproducible run to run.
//   * a warp walks its edge list 32 edges at a time: one coalesced streaming load of
//     32 column indices (+ weights), software-pipelined one chunk ahead, then
//     the feature rows are gathered with 128-bit loads.  LPR lanes cover one feature
//     row (K = 32 fp32 -> 8 lanes x float4 = one 128-byte line), so 32/LPR edges are
//     in flight per load instruction and LPR independent loads per lane per chunk.
//   * feature tiles wider than one warp pass are mapped to blockIdx.y.
#pragma once
#include "common.cuh"

namespace gala {

// MODE_GAT_DOT: like MODE_GAT, but aR[col] = dot(X[col,:], wR) + bR is recomputed from the
// gathered feature row itself, which removes the second random gather per edge.
// MODE_GAT_COL: the caller stores the features in a reflected basis whose LAST column is (a multiple of) the
// right-hand attention term, aR[j] = sR * X[j, K-1] + bR (see gala_gat_forward_col_f32): the scalar arrives with
// the 128-byte row the edge gathers anyway, so an edge costs 4 sectors instead of 4 + 1, and no dot product.
enum { MODE_PLAIN = 0, MODE_GAT = 1, MODE_GAT_DOT = 2, MODE_GAT_COL = 3 };
#define GALA_IS_GAT(M) ((M) == MODE_GAT || (M) == MODE_GAT_DOT || (M) == MODE_GAT_COL)
#define GALA_GAT_FROM_ROW(M) ((M) == MODE_GAT_DOT || (M) == MODE_GAT_COL)   // edge weight known only after the gather

struct SpmmParams {
    GraphDev g;
    const float* __restrict__ vals;       // per-edge weights or nullptr
    const float* __restrict__ X;          // [ncols, K]
    float* __restrict__ Y;                // [nrows, K]
    int K;
    int64_t ldx, ldy;                     // row pitch of X / Y in elements (>= K; K when the rows are packed)
    const float* __restrict__ row_scale;  // nullable
    const float* __restrict__ col_scale;  // nullable
    int accumulate;
    int relu;
    int scale_after;                      // 1: Y = act(row_scale * (sum + Y_old)) -- last pass of a segment-major run
    TaskParams t;
    // MODE_GAT
    const float* __restrict__ aL;
    const float* __restrict__ aR;
    float slope;
    float* __restrict__ alpha_out;        // nullable
    float seed_total;                     // S * 1e-12f
    const float* __restrict__ wR;         // MODE_GAT_DOT: aR[j] = dot(X[j,:], wR) + bR
    float bR;
    float sR;                             // MODE_GAT_COL: aR[j] = sR * X[j, K-1] + bR
    const float* __restrict__ refl_in;    // MODE_GAT_COL, nullable [K] unit vector v: y <- y - 2 v (v.y) on the normalised
                                          // sum (back from the reflected basis), before the ReLU
    const float* __restrict__ refl_out;   // nullable [K]: the same reflection with this vector after the ReLU (into the
                                          // basis the NEXT layer gathers in)
    // fused dense epilogue on the finished output row y (K <= 128, one feature tile):
    const float* __restrict__ att_w;      // [2, K]: att_out[row] = y.att_w[0] + att_b0, att_out[nrows+row] = y.att_w[1] + att_b1
    float att_b0, att_b1;
    float* __restrict__ att_out;
    const float* __restrict__ cls_wT;     // [K, cls_n] (transposed Linear weight): cls_out[row,:] = y @ cls_wT + cls_b
    const float* __restrict__ cls_b;      // [cls_n] or nullptr
    float* __restrict__ cls_out;          // [nrows, cls_n]
    int cls_n;
    MultiOut mo;                          // count > 0: Y rows go to every GPU (Y itself unused)
    MultiOut att_mo;                      // count > 0: the second projection of the epilogue (next layer's attenR) is also
                                          // stored at element `row` of every GPU's gathered vector
};


// Gather the feature rows of the (up to) 32 edges a warp holds one-per-lane and
// accumulate them.  All loads of a batch are issued before the first FMA that
// consumes them (UNR x ACC independent 16-byte loads in flight per lane).
//   FULL : every lane of the chunk holds a valid edge (no per-edge checks)
//   EXACT: K is a multiple of the tile width (no per-lane feature predicates)
// xlane = X + tile_base + sub*VEC (the lane's first feature); the row address is one
// IMAD.WIDE.U32: xlane + col * row_bytes.
template <int VEC, int LPR, int ACC, bool FULL, bool EXACT>
__device__ __forceinline__ void gather_chunk(const char* __restrict__ xlane, uint32_t row_bytes, int grp,
                                             int c, float w, bool weighted, const bool (&fvalid)[ACC],
                                             float (&acc)[ACC][VEC]) {
    constexpr int EPI = 32 / LPR;
    constexpr int UNR_ = 32 / (ACC * VEC) < 1 ? 1 : 32 / (ACC * VEC);
    constexpr int UNR = UNR_ > LPR ? LPR : (UNR_ > 16 ? 16 : UNR_);
#pragma unroll
    for (int j0 = 0; j0 < LPR; j0 += UNR) {
        int cj[UNR];
        float wj[UNR];
#pragma unroll
        for (int u = 0; u < UNR; ++u) {
            cj[u] = __shfl_sync(kFull, c, (j0 + u) * EPI + grp);
            wj[u] = weighted ? __shfl_sync(kFull, w, (j0 + u) * EPI + grp) : 1.0f;
        }
        Vec<VEC> x[UNR][ACC];
#pragma unroll
        for (int u = 0; u < UNR; ++u) {
            const uint32_t cc = (uint32_t)(FULL ? cj[u] : max(cj[u], 0));
            const char* xr = xlane + (uint64_t)cc * row_bytes;
#pragma unroll
            for (int a = 0; a < ACC; ++a) {
                if (EXACT) {
                    x[u][a].load(reinterpret_cast<const float*>(xr) + a * LPR * VEC);
                } else {
#pragma unroll
                    for (int v = 0; v < VEC; ++v) x[u][a].v[v] = 0.0f;
                    if (fvalid[a]) x[u][a].load(reinterpret_cast<const float*>(xr) + a * LPR * VEC);
                }
            }
        }
#pragma unroll
        for (int u = 0; u < UNR; ++u) {
            const bool ok = FULL || cj[u] >= 0;
            const float ww = ok ? wj[u] : 0.0f;
#pragma unroll
            for (int a = 0; a < ACC; ++a)
#pragma unroll
                for (int v = 0; v < VEC; ++v)
                    acc[a][v] = fmaf(ww, ok ? x[u][a].v[v] : 0.0f, acc[a][v]);
        }
    }
}


// MODE_GAT_DOT chunk: gather the 32 rows (LPR loads per lane, all in flight), form each
// edge's dot product with wR on the LPR lanes that hold the row, reduce the LPR partials of
// the LPR edges of a group with a halving exchange (LPR-1 shuffles; lane `sub` ends up with
// edge `sub` of its group), evaluate the softmax numerator ONCE per edge on that lane, then
// broadcast it back as the weight of the row that is already sitting in registers.
// Requires ACC == 1 and a batch that covers all LPR sub-iterations (LPR * VEC <= 32).
template <int VEC, int LPR, bool FULL>
__device__ __forceinline__ void gather_chunk_dot(const char* __restrict__ xlane, uint32_t row_bytes, int sub, int grp,
                                                 int c, int base, int e1, const float (&wreg)[VEC], float aL_row,
                                                 float bR, float slope, float* __restrict__ alpha_out, float& rs,
                                                 float (&acc)[1][VEC]) {
    constexpr int EPI = 32 / LPR;
    int cj[LPR];
    Vec<VEC> x[LPR];
#pragma unroll
    for (int u = 0; u < LPR; ++u) cj[u] = __shfl_sync(kFull, c, u * EPI + grp);
#pragma unroll
    for (int u = 0; u < LPR; ++u) {
        const uint32_t cc = (uint32_t)(FULL ? cj[u] : max(cj[u], 0));
        x[u].load(reinterpret_cast<const float*>(xlane + (uint64_t)cc * row_bytes));
    }
    float d[LPR];
#pragma unroll
    for (int u = 0; u < LPR; ++u) {
        float t = 0.0f;
#pragma unroll
        for (int v = 0; v < VEC; ++v) t = fmaf(x[u].v[v], wreg[v], t);
        d[u] = t;
    }
#pragma unroll
    for (int o = LPR >> 1, n = LPR; o > 0; o >>= 1, n >>= 1) {
        const bool upper = (sub & o) != 0;
#pragma unroll
        for (int i = 0; i < (n >> 1); ++i) {
            const float send = upper ? d[i] : d[i + (n >> 1)];
            const float keep = upper ? d[i + (n >> 1)] : d[i];
            d[i] = keep + __shfl_xor_sync(kFull, send, o);
        }
    }
    // this lane now owns chunk position sub*EPI + grp
    const int pos = base + sub * EPI + grp;
    float e = 0.0f;
    if (FULL || pos < e1) {
        e = softmax_num(leaky(aL_row + (d[0] + bR), slope));
        rs += e;
        if (alpha_out) alpha_out[pos] = e;
    }
#pragma unroll
    for (int u = 0; u < LPR; ++u) {
        const float wj = __shfl_sync(kFull, e, grp * LPR + u);   // 0 for invalid edges
        const bool ok = FULL || cj[u] >= 0;
#pragma unroll
        for (int v = 0; v < VEC; ++v) acc[0][v] = fmaf(wj, ok ? x[u].v[v] : 0.0f, acc[0][v]);
    }
}


// MODE_GAT_COL chunk: gather the 32 rows (LPR loads per lane, all in flight).  The lane that holds the last 16-byte
// piece of a row (sub == LPR-1) has that edge's attention scalar in its last element: it drops the LPR scalars of its
// group into the warp's 32-float staging line, every lane picks up the scalar of ITS OWN edge (slot = lane), evaluates
// the softmax numerator once, writes it back to the same slot, and the LPR weights of the group come back with
// vector loads; the column ids reach the groups the same way (one store, LPR/4 vector loads).  Three warp-local
// syncs and a handful of shared-memory instructions per chunk replace the second global gather of MODE_GAT with its
// 2 x LPR shuffles, and the LPR-1 shuffles + LPR*VEC FMAs of MODE_GAT_DOT.  TW == K (one exact tile).
#ifndef GALA_GATCOL_SMEM_COLS
#define GALA_GATCOL_SMEM_COLS 1   // 1: the column ids also travel through the staging line (no shuffles at all)
#endif
template <int VEC, int LPR, bool FULL>
__device__ __forceinline__ void gather_chunk_col(const char* __restrict__ xlane, uint32_t row_bytes, int sub, int grp,
                                                 int lane, int c, int base, int e1, float aL_row, float sR, float bR,
                                                 float slope, float* __restrict__ alpha_out, float& rs,
                                                 float (&acc)[1][VEC], float* __restrict__ line,
                                                 int* __restrict__ cline) {
    // Edge at chunk position q = grp*LPR + u is gathered by group grp in sub-iteration u; lane q evaluates its weight.
    int cj[LPR];
    Vec<VEC> x[LPR];
#if GALA_GATCOL_SMEM_COLS
    cline[lane] = c;
    __syncwarp();       // also orders the previous chunk's reads of `line` before this chunk's writes
    if constexpr (LPR % 4 == 0) {
#pragma unroll
        for (int u = 0; u < LPR; u += 4) {
            const int4 t = *reinterpret_cast<const int4*>(cline + grp * LPR + u);
            cj[u] = t.x; cj[u + 1] = t.y; cj[u + 2] = t.z; cj[u + 3] = t.w;
        }
    } else {
#pragma unroll
        for (int u = 0; u < LPR; ++u) cj[u] = cline[grp * LPR + u];
    }
#else
#pragma unroll
    for (int u = 0; u < LPR; ++u) cj[u] = __shfl_sync(kFull, c, grp * LPR + u);   // (also a warp-wide sync point)
#endif
#pragma unroll
    for (int u = 0; u < LPR; ++u) {
        const uint32_t cc = (uint32_t)(FULL ? cj[u] : max(cj[u], 0));
        x[u].load(reinterpret_cast<const float*>(xlane + (uint64_t)cc * row_bytes));
    }
    if (sub == LPR - 1) {
        if constexpr (LPR % 4 == 0) {
#pragma unroll
            for (int u = 0; u < LPR; u += 4)
                *reinterpret_cast<float4*>(line + grp * LPR + u) =
                    make_float4(x[u].v[VEC - 1], x[u + 1].v[VEC - 1], x[u + 2].v[VEC - 1], x[u + 3].v[VEC - 1]);
        } else {
#pragma unroll
            for (int u = 0; u < LPR; ++u) line[grp * LPR + u] = x[u].v[VEC - 1];
        }
    }
    __syncwarp();
    // slot `lane` is this lane's own edge (column c, chunk position lane); it is read and rewritten by this lane only
    float e = 0.0f;
    if (FULL || c >= 0) {
        e = softmax_num(leaky(aL_row + fmaf(sR, line[lane], bR), slope));
        rs += e;
        if (alpha_out) alpha_out[base + lane] = e;
    }
    line[lane] = e;     // 0 for the positions past the row's end
    __syncwarp();
    float wj[LPR];
    if constexpr (LPR % 4 == 0) {
#pragma unroll
        for (int u = 0; u < LPR; u += 4) {
            const float4 t = *reinterpret_cast<const float4*>(line + grp * LPR + u);
            wj[u] = t.x; wj[u + 1] = t.y; wj[u + 2] = t.z; wj[u + 3] = t.w;
        }
    } else {
#pragma unroll
        for (int u = 0; u < LPR; ++u) wj[u] = line[grp * LPR + u];
    }
#pragma unroll
    for (int u = 0; u < LPR; ++u) {
        const bool ok = FULL || cj[u] >= 0;
#pragma unroll
        for (int v = 0; v < VEC; ++v) acc[0][v] = fmaf(wj[u], ok ? x[u].v[v] : 0.0f, acc[0][v]);
    }
#if !GALA_GATCOL_SMEM_COLS
    __syncwarp();       // the line is rewritten by the next chunk
#endif
}

// y <- y - 2 v (v . y) for the K = VEC*LPR features a group of LPR lanes holds (VEC each); every lane of the group
// returns with its reflected elements.  v is a unit vector (a Householder reflection: orthogonal and its own inverse).
template <int VEC, int LPR>
__device__ __forceinline__ void reflect_row(float (&y)[VEC], const float* __restrict__ v, int sub) {
    float vv[VEC];
    float d = 0.0f;
#pragma unroll
    for (int i = 0; i < VEC; ++i) {
        vv[i] = __ldg(v + sub * VEC + i);
        d = fmaf(vv[i], y[i], d);
    }
#pragma unroll
    for (int o = 1; o < LPR; o <<= 1) d += __shfl_xor_sync(kFull, d, o);
    d *= 2.0f;
#pragma unroll
    for (int i = 0; i < VEC; ++i) y[i] = fmaf(-d, vv[i], y[i]);
}


constexpr int kRowBufMax = 128;   // widest output row the fused dense epilogue handles

// Dense epilogue of one finished output row held in shared memory (K floats): the next layer's two
// attention projections and / or the classifier transform, computed by the warp that owns the row.
__device__ __forceinline__ void row_dense_epilogue(const SpmmParams& p, const float* __restrict__ rowbuf, int row,
                                                   int lane) {
    if (p.att_w) {
        float a0 = 0.0f, a1 = 0.0f;
        for (int f = lane; f < p.K; f += 32) {
            const float y = rowbuf[f];
            a0 = fmaf(y, __ldg(p.att_w + f), a0);
            a1 = fmaf(y, __ldg(p.att_w + p.K + f), a1);
        }
        a0 = warp_sum(a0);
        a1 = warp_sum(a1);
        if (lane == 0) {
            p.att_out[row] = a0 + p.att_b0;
            p.att_out[p.g.nrows + row] = a1 + p.att_b1;
            if (p.att_mo.count > 0) {
                Vec<1> o1;
                o1.v[0] = a1 + p.att_b1;
                multi_store<1>(p.att_mo, row, o1, row);
            }
        }
    }
    if (p.cls_wT) {
        for (int n = lane; n < p.cls_n; n += 32) {
            float o = p.cls_b ? __ldg(p.cls_b + n) : 0.0f;
#pragma unroll 8
            for (int f = 0; f < p.K; ++f) o = fmaf(rowbuf[f], __ldg(p.cls_wT + f * p.cls_n + n), o);
            p.cls_out[(int64_t)row * p.cls_n + n] = o;
        }
    }
}

// Minimum resident CTAs per SM the register allocator must allow.  Decides how many of
// the UNR gathers ptxas keeps in flight (it serialises them when squeezed below ~64
// registers); tuned on B200, see profiles/r01_variants.txt.
#ifndef GALA_SPMM_MINB
#define GALA_SPMM_MINB 5
#endif
#ifndef GALA_GATDOT_MINB
#define GALA_GATDOT_MINB 3   // the dot mode keeps 8 rows + 8 partial dots live
// (a software-pipelined form -- two chunks in registers, the gathers of chunk i+1 in flight while chunk i is reduced --
//  was measured in round 2: 1.53 ms at 128 registers / 2 CTAs per SM against 1.16 ms for this one and 1.07 ms for the
//  two-gather MODE_GAT; profiles/r02_variants_dot_pipeline_softmax_cache.txt.  Occupancy, not the dependent chain, is
//  what the gather lives on.)
#endif
#ifndef GALA_GATCOL_MINB
#define GALA_GATCOL_MINB 4   // the column mode keeps the 8 gathered rows of a chunk live until all their weights are
// known: 64 registers (4 CTAs per SM) hold that without spills -- 0.92 ms on the Reddit shape; squeezed to 48 registers
// (5 CTAs) the accumulators spill inside the chunk loop, 1.40 ms; 3 CTAs 1.04 ms (profiles/r02_variants_col.txt)
#endif
#if GALA_SPMM_MINB > 0
#define GALA_SPMM_BOUNDS                                                                                   \
    __launch_bounds__(kCtaThreads, (MODE == MODE_GAT_DOT ? GALA_GATDOT_MINB                                \
                                                         : (MODE == MODE_GAT_COL ? GALA_GATCOL_MINB : GALA_SPMM_MINB)))
#else
#define GALA_SPMM_BOUNDS __launch_bounds__(kCtaThreads)
#endif
template <int VEC, int LPR, int ACC, int MODE, bool EXACT>
__global__ void GALA_SPMM_BOUNDS
spmm_kernel(const __grid_constant__ SpmmParams p) {
    constexpr int TW = VEC * LPR * ACC;  // features covered by one warp pass (32 / LPR edges per load instruction)
    const GraphDev& g = p.g;
    const int lane = threadIdx.x & 31;
    const int warp = threadIdx.x >> 5;
    const int sub = lane % LPR;
    const int grp = lane / LPR;
    const int tile_base = blockIdx.y * TW;
    const RowTask task = row_task(g, p.t);
    if (!task.valid) return;
    const bool hub_cta = task.hub;
    const int row = task.row, lo = task.lo, hi = task.hi;
    // VEC == 8 is the bf16-feature flavour: X holds bf16 rows (2 bytes per feature), see Vec<8>
    constexpr uint32_t XB = VEC == 8 ? 2u : 4u;
    static_assert(VEC != 8 || (ACC == 1 && EXACT && !GALA_GAT_FROM_ROW(MODE)), "bf16 rows: one exact tile, no dot mode");
    static_assert(MODE != MODE_GAT_COL || (ACC == 1 && EXACT), "column mode: the row is one exact tile");
    __shared__ __align__(16) float colline[MODE == MODE_GAT_COL ? kWarpsPerCta : 1][32];
    __shared__ __align__(16) int colcols[MODE == MODE_GAT_COL ? kWarpsPerCta : 1][32];
    const uint32_t row_bytes = (uint32_t)p.ldx * XB;
    // Y rows go out as VEC-wide stores when their pitch and base keep that alignment; otherwise (packed rows of
    // odd width) element by element -- N*K*4 bytes once per launch, nothing next to the gather
    const bool yvec = VEC == 1 || VEC == 8 || (p.ldy % VEC == 0 && (reinterpret_cast<uintptr_t>(p.Y) % (VEC * 4)) == 0);
    // the lane's first feature; lanes past K (non-EXACT shapes) point at the tile start
    const char* xlane = reinterpret_cast<const char*>(p.X) +
                        (size_t)(tile_base + ((EXACT || tile_base + sub * VEC < p.K) ? sub * VEC : 0)) * XB;

    bool fvalid[ACC];
#pragma unroll
    for (int a = 0; a < ACC; ++a) fvalid[a] = tile_base + (a * LPR + sub) * VEC < p.K;

    float acc[ACC][VEC];
#pragma unroll
    for (int a = 0; a < ACC; ++a)
#pragma unroll
        for (int v = 0; v < VEC; ++v) acc[a][v] = 0.0f;

    const bool weighted = GALA_IS_GAT(MODE) || p.vals != nullptr || p.col_scale != nullptr;
    float aL_row = 0.0f, rs = 0.0f;
    if (GALA_IS_GAT(MODE)) aL_row = __ldg(p.aL + row);
    const bool write_alpha = GALA_IS_GAT(MODE) && p.alpha_out != nullptr && blockIdx.y == 0;
    float wreg[VEC];
#pragma unroll
    for (int v = 0; v < VEC; ++v) wreg[v] = 0.0f;
    if (MODE == MODE_GAT_DOT) {
#pragma unroll
        for (int v = 0; v < VEC; ++v)
            if (sub * VEC + v < p.K) wreg[v] = __ldg(p.wR + sub * VEC + v);
    }

    // one edge per lane: column index and the weight the edge contributes with
    auto fetch = [&](int idx, int e1, int& c, float& w) {
        c = -1;
        w = 0.0f;
        if (idx < e1) {
            c = ld_stream(g.cols + idx);
            if (GALA_GAT_FROM_ROW(MODE)) {
                // weight comes from the gathered row (gather_chunk_dot / gather_chunk_col)
            } else if (MODE == MODE_GAT) {
                float e = softmax_num(leaky(aL_row + ld_keep(p.aR + c), p.slope));
                rs += e;
                if (write_alpha) p.alpha_out[idx] = e;
                w = e;
            } else {
                w = p.vals ? ld_stream(p.vals + idx) : 1.0f;
                if (p.col_scale) w *= ld_keep(p.col_scale + c);
            }
        }
    };

    for_each_chunk(g, row, lo, hi, [&](int e0, int e1) {
        int c_nxt;
        float w_nxt;
        fetch(e0 + lane, e1, c_nxt, w_nxt);
        for (int base = e0; base < e1; base += 32) {
            const int c = c_nxt;
            const float w = w_nxt;
            if (base + 32 < e1) fetch(base + 32 + lane, e1, c_nxt, w_nxt);
            if constexpr (MODE == MODE_GAT_DOT) {
                float* ao = write_alpha ? p.alpha_out : nullptr;
                if (base + 32 <= e1)
                    gather_chunk_dot<VEC, LPR, true>(xlane, row_bytes, sub, grp, c, base, e1, wreg, aL_row, p.bR,
                                                     p.slope, ao, rs, acc);
                else
                    gather_chunk_dot<VEC, LPR, false>(xlane, row_bytes, sub, grp, c, base, e1, wreg, aL_row, p.bR,
                                                      p.slope, ao, rs, acc);
            } else if constexpr (MODE == MODE_GAT_COL) {
                float* ao = write_alpha ? p.alpha_out : nullptr;
                if (base + 32 <= e1)
                    gather_chunk_col<VEC, LPR, true>(xlane, row_bytes, sub, grp, lane, c, base, e1, aL_row, p.sR, p.bR,
                                                     p.slope, ao, rs, acc, colline[warp], colcols[warp]);
                else
                    gather_chunk_col<VEC, LPR, false>(xlane, row_bytes, sub, grp, lane, c, base, e1, aL_row, p.sR, p.bR,
                                                      p.slope, ao, rs, acc, colline[warp], colcols[warp]);
            } else if (base + 32 <= e1)
                gather_chunk<VEC, LPR, ACC, true, EXACT>(xlane, row_bytes, grp, c, w, weighted, fvalid, acc);
            else
                gather_chunk<VEC, LPR, ACC, false, EXACT>(xlane, row_bytes, grp, c, w, weighted, fvalid, acc);
        }
    });

    // combine the EPI edge groups of the warp (fixed tree -> deterministic)
#pragma unroll
    for (int o = LPR; o < 32; o <<= 1)
#pragma unroll
        for (int a = 0; a < ACC; ++a)
#pragma unroll
            for (int v = 0; v < VEC; ++v) acc[a][v] += __shfl_xor_sync(kFull, acc[a][v], o);
    if (GALA_IS_GAT(MODE)) rs = warp_sum(rs);

    float scale = p.row_scale ? __ldg(p.row_scale + row) : 1.0f;

    __shared__ float rowbuf[kWarpsPerCta][kRowBufMax];
    const bool dense_ep = p.att_w != nullptr || p.cls_wT != nullptr;   // host guarantees K <= kRowBufMax, one tile

    if constexpr (MODE == MODE_GAT_COL) {
        // One exact tile, K = VEC*LPR <= 32: normalise, reflect back from the gathered basis, ReLU, reflect into the
        // next layer's basis.  Hub rows first combine their 8 warp partials in warp order (as below); every warp of
        // the hub CTA then holds the same full sums and warp 0 finishes the row.
        if (hub_cta) {
            __shared__ float cpart[kWarpsPerCta * 32];
            __shared__ float cpart_rs[kWarpsPerCta];
            if (grp == 0) {
#pragma unroll
                for (int v = 0; v < VEC; ++v) cpart[warp * 32 + sub * VEC + v] = acc[0][v];
            }
            if (lane == 0) cpart_rs[warp] = rs;
            __syncthreads();
            rs = 0.0f;
#pragma unroll
            for (int v = 0; v < VEC; ++v) acc[0][v] = 0.0f;
#pragma unroll
            for (int w = 0; w < kWarpsPerCta; ++w) {
                rs += cpart_rs[w];
#pragma unroll
                for (int v = 0; v < VEC; ++v) acc[0][v] += cpart[w * 32 + sub * VEC + v];
            }
        }
        scale = 1.0f / (rs + p.seed_total);
        if (!hub_cta || warp == 0) {
            float y[VEC];
#pragma unroll
            for (int v = 0; v < VEC; ++v) y[v] = acc[0][v] * scale;
            if (p.refl_in) reflect_row<VEC, LPR>(y, p.refl_in, sub);
            if (p.relu) {
#pragma unroll
                for (int v = 0; v < VEC; ++v) y[v] = fmaxf(y[v], 0.0f);
            }
            if (p.refl_out) reflect_row<VEC, LPR>(y, p.refl_out, sub);
            // the next layer's LEFT-hand projection of the final row (its right-hand one is that layer's last
            // column): from the registers that hold the row, no staging -- att_out[0:nrows] only
            const bool att_only = p.att_w != nullptr && p.cls_wT == nullptr;
            if (att_only) {
                float d = 0.0f;
#pragma unroll
                for (int v = 0; v < VEC; ++v) d = fmaf(y[v], __ldg(p.att_w + sub * VEC + v), d);
#pragma unroll
                for (int o = 1; o < LPR; o <<= 1) d += __shfl_xor_sync(kFull, d, o);
                if (lane == 0) p.att_out[row] = d + p.att_b0;
            }
            if (grp == 0) {
                Vec<VEC> o;
#pragma unroll
                for (int v = 0; v < VEC; ++v) {
                    o.v[v] = y[v];
                    if (dense_ep && !att_only) rowbuf[warp][sub * VEC + v] = y[v];
                }
                if (p.mo.count > 0) multi_store<VEC>(p.mo, (int64_t)row * p.K + sub * VEC, o, row);
                else if (p.Y) o.store(p.Y + (int64_t)row * p.ldy + sub * VEC);
            }
            if (dense_ep && !att_only) {
                __syncwarp();
                row_dense_epilogue(p, rowbuf[warp], row, lane);
            }
        }
        if (write_alpha) {
            __syncwarp();   // the numerators were stored by other lanes of this warp
            for_each_chunk(g, row, lo, hi, [&](int e0, int e1) {
                for (int e = e0 + lane; e < e1; e += 32) p.alpha_out[e] *= scale;
            });
        }
        return;
    }

    if (!hub_cta) {
        if (GALA_IS_GAT(MODE)) scale = 1.0f / (rs + p.seed_total);
        if (grp == 0) {
#pragma unroll
            for (int a = 0; a < ACC; ++a) {
                if (!fvalid[a]) continue;
                const int f0 = tile_base + (a * LPR + sub) * VEC;
                float* y = p.Y + (int64_t)row * p.ldy + f0;
                // whole vector inside the row and aligned: one VEC-wide access; else element by element
                const bool whole = EXACT || (yvec && f0 + VEC <= p.K);
                Vec<VEC> o;
                if (p.accumulate) {
                    if (whole) o.load_rw(y);
                    else {
#pragma unroll
                        for (int v = 0; v < VEC; ++v) o.v[v] = (f0 + v < p.K) ? y[v] : 0.0f;
                    }
                }
#pragma unroll
                for (int v = 0; v < VEC; ++v) {
                    float t = p.scale_after ? acc[a][v] : acc[a][v] * scale;
                    if (p.accumulate) t += o.v[v];
                    if (p.scale_after) t *= scale;
                    if (p.relu) t = fmaxf(t, 0.0f);
                    o.v[v] = t;
                    if (dense_ep && (EXACT || f0 + v < p.K)) rowbuf[warp][f0 + v] = t;
                }
                if constexpr (VEC <= 4) {
                    if (p.mo.count > 0) multi_store<VEC>(p.mo, (int64_t)row * p.K + f0, o, row);
                    else if (p.Y) {
                        if (whole) o.store(y);
                        else {
#pragma unroll
                            for (int v = 0; v < VEC; ++v)
                                if (f0 + v < p.K) y[v] = o.v[v];
                        }
                    }
                } else {
                    if (p.Y) o.store(y);
                }
            }
        }
        if (dense_ep) {
            __syncwarp();
            row_dense_epilogue(p, rowbuf[warp], row, lane);
        }
        if (write_alpha) {
            __syncwarp();   // MODE_GAT_DOT: the numerators were stored by other lanes of this warp
            for_each_chunk(g, row, lo, hi, [&](int e0, int e1) {
                for (int e = e0 + lane; e < e1; e += 32) p.alpha_out[e] *= scale;
            });
        }
        return;
    }

    // ---- hub row: combine the 8 warp partials in warp order ----
    __shared__ float part[kWarpsPerCta * TW];
    __shared__ float part_rs[kWarpsPerCta];
    if (grp == 0) {
#pragma unroll
        for (int a = 0; a < ACC; ++a)
#pragma unroll
            for (int v = 0; v < VEC; ++v) part[warp * TW + (a * LPR + sub) * VEC + v] = acc[a][v];
    }
    if (GALA_IS_GAT(MODE) && lane == 0) part_rs[warp] = rs;
    __syncthreads();
    if (GALA_IS_GAT(MODE)) {
        float t = 0.0f;
#pragma unroll
        for (int w = 0; w < kWarpsPerCta; ++w) t += part_rs[w];
        scale = 1.0f / (t + p.seed_total);
    }
    for (int f = threadIdx.x; f < TW; f += kCtaThreads) {
        if (tile_base + f >= p.K) break;
        float t = 0.0f;
#pragma unroll
        for (int w = 0; w < kWarpsPerCta; ++w) t += part[w * TW + f];
        float* y = p.Y + (int64_t)row * p.ldy + tile_base + f;
        if (!p.scale_after) t *= scale;
        if (p.accumulate) t += *y;
        if (p.scale_after) t *= scale;
        if (p.relu) t = fmaxf(t, 0.0f);
        if (p.mo.count > 0) {
            Vec<1> o1;
            o1.v[0] = t;
            multi_store<1>(p.mo, (int64_t)row * p.K + tile_base + f, o1, row);
        } else if (p.Y) {
            *y = t;
        }
        if (dense_ep) rowbuf[0][f] = t;
    }
    if (dense_ep) {
        __syncthreads();
        if (warp == 0) row_dense_epilogue(p, rowbuf[0], row, lane);
    }
    if (write_alpha) {
        for_each_chunk(g, row, lo, hi, [&](int e0, int e1) {
            for (int e = e0 + lane; e < e1; e += 32) p.alpha_out[e] *= scale;
        });
    }
}

// ---- sampled aggregation (K1s) ---------------------------------------------------
struct SampledParams {
    GraphDev g;
    const float* __restrict__ vals;
    const float* __restrict__ X;
    float* __restrict__ Y;
    int K;
    int64_t ldx, ldy;
    int nsamples, ra, rb;
    int accumulate;
};

template <int VEC, int LPR, int ACC>
__global__ void __launch_bounds__(kCtaThreads)
spmm_sampled_kernel(const __grid_constant__ SampledParams p) {
    constexpr int TW = VEC * LPR * ACC;
    constexpr int EPI = 32 / LPR;
    const GraphDev& g = p.g;
    const int lane = threadIdx.x & 31;
    const int warp = threadIdx.x >> 5;
    const int sub = lane % LPR;
    const int grp = lane / LPR;
    const int tile_base = blockIdx.y * TW;
    const int row = blockIdx.x * kWarpsPerCta + warp;
    if (row >= g.nrows) return;
    const bool yvec = VEC == 1 || (p.ldy % VEC == 0 && (reinterpret_cast<uintptr_t>(p.Y) % (VEC * 4)) == 0);

    bool fvalid[ACC];
#pragma unroll
    for (int a = 0; a < ACC; ++a) fvalid[a] = tile_base + (a * LPR + sub) * VEC < p.K;
    float acc[ACC][VEC];
#pragma unroll
    for (int a = 0; a < ACC; ++a)
#pragma unroll
        for (int v = 0; v < VEC; ++v) acc[a][v] = 0.0f;

    for (int s = 0; s < g.S; ++s) {
        const int* off = g.offsets + (int64_t)s * (g.nrows + 1) + row;
        const int b = __ldg(off);
        const int jmax = __ldg(off + 1) - b;
        if (jmax <= 0) continue;
        const int base = seg_start(g, s) + b;
        for (int ji = grp; ji < p.nsamples; ji += EPI) {
            const int j = (p.ra * ji + p.rb) % jmax;  // cuda.h:320, int arithmetic as emitted
            const int c = __ldg(g.cols + base + j);
            const float w = p.vals ? __ldg(p.vals + base + j) : 1.0f;
            const float* xr = p.X + (int64_t)c * p.ldx + tile_base + sub * VEC;
#pragma unroll
            for (int a = 0; a < ACC; ++a) {
                if (fvalid[a]) {
                    Vec<VEC> x;
                    x.load(xr + a * LPR * VEC);
#pragma unroll
                    for (int v = 0; v < VEC; ++v) acc[a][v] = fmaf(w, x.v[v], acc[a][v]);
                }
            }
        }
    }
#pragma unroll
    for (int o = LPR; o < 32; o <<= 1)
#pragma unroll
        for (int a = 0; a < ACC; ++a)
#pragma unroll
            for (int v = 0; v < VEC; ++v) acc[a][v] += __shfl_xor_sync(kFull, acc[a][v], o);
    if (grp == 0) {
#pragma unroll
        for (int a = 0; a < ACC; ++a) {
            if (!fvalid[a]) continue;
            const int f0 = tile_base + (a * LPR + sub) * VEC;
            float* y = p.Y + (int64_t)row * p.ldy + f0;
            if (yvec && f0 + VEC <= p.K) {
                Vec<VEC> o;
                if (p.accumulate) o.load_rw(y);
#pragma unroll
                for (int v = 0; v < VEC; ++v) o.v[v] = p.accumulate ? o.v[v] + acc[a][v] : acc[a][v];
                o.store(y);
            } else {
#pragma unroll
                for (int v = 0; v < VEC; ++v)
                    if (f0 + v < p.K) y[v] = p.accumulate ? y[v] + acc[a][v] : acc[a][v];
            }
        }
    }
}

}  // namespace gala
