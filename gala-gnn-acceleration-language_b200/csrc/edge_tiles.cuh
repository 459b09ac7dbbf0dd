// edge_tiles.cuh -- edge-parallel ("nnz-split", the 1-D merge-path decomposition) forms of the streaming edge
// kernels: row sums (K3), per-row scaling (K4), edge-softmax forward / backward.
//
// The row-structured kernels of edge_ops.cuh give every row to one warp, so the bytes a warp keeps in flight follow the
// row's length and each row pays two dependent global round trips (sum, then scale).  Here the EDGE array is cut into
// tiles of kTileEdges consecutive edges; tile t owns the rows that START inside [t*kTileEdges, (t+1)*kTileEdges)
// (plan->tile_rows, a binary search per tile on the row pointers at plan time -- cf. the nnz split of
// nnz_ord_row_tile_info, reference src/ops/tiling.h:1656-1708), so a tile's edges are one contiguous, disjoint span
// of the edge arrays whatever the degree distribution.  Persistent, warp-specialised CTAs walk the tiles round-robin
// through a three-stage shared-memory ring:
//   producer warp   moves each tile's span AND its row pointers global -> shared memory with 1-D bulk copies
//                   (cp.async.bulk, the TMA engine: no registers, completion on the stage's "full" mbarrier), two tiles
//                   ahead of the consumers, and writes every finished span back with one bulk store (16-byte aligned
//                   interior) plus <= 6 scalar stores once the consumers have arrived on the stage's "done" mbarrier;
//   16 consumer warps  a flat, row-agnostic pass for the per-edge arithmetic (exp / products), then the rows of the
//                   tile reduced and rescaled IN shared memory by groups of G lanes (G follows the mean degree).
// Every edge array is read once and written once from / to HBM with full-line accesses, independent of row length,
// and no global-memory latency sits between the phases of a tile.
// Rows that do not fit the staged window (longer than the window, or behind such a row in the same tile) are
// processed by the consumer warps together, straight from global memory, two passes, like a hub row of the
// row-structured kernels.
// Single-segment graphs only (the GAT schedules: col_tile >= ncols); column-tiled graphs keep the row-structured form.
#pragma once
#include "edge_ops.cuh"

namespace gala {

constexpr int kTileEdges = 4096;   // edges per tile == gala_plan_t.tile_edges
constexpr int kTileCap = 8192;     // staged window, floats per edge array: a row of <= 4096 edges always fits
constexpr int kTileRowCap = 1024;  // row pointers staged per tile; rows beyond read theirs from global memory
#ifndef GALA_TILE_STAGES
#define GALA_TILE_STAGES 3
#endif
constexpr int kTileStages = GALA_TILE_STAGES;
constexpr int kTileConsumers = 512;                  // 16 consumer warps
constexpr int kTileThreads = kTileConsumers + 32;    // + the producer warp
constexpr int kBatch = 8;          // shared-memory loads in flight per thread in the tile loops

enum : int { TILE_ROWSUM = 0, TILE_SCALE = 1, TILE_SOFTMAX_FWD = 2, TILE_SOFTMAX_BWD = 3 };

__host__ __device__ constexpr int tile_narr(int op) { return op == TILE_SOFTMAX_BWD ? 2 : 1; }   // staged edge arrays
__host__ __device__ constexpr size_t tile_stage_bytes(int op);
__host__ __device__ constexpr int tile_ctas_per_sm(int op);
__host__ __device__ constexpr size_t tile_stage_bytes(int op) {
    return (size_t)tile_narr(op) * kTileCap * 4 + (size_t)(kTileRowCap + 8) * 4;
}

// persistent CTAs per SM: what the shared-memory ring leaves room for (227 KB per SM, 1 KB reserved per CTA)
__host__ __device__ constexpr int tile_ctas_per_sm(int op) {
    return 2 * (kTileStages * tile_stage_bytes(op) + 2048) <= 227 * 1024 ? 2 : 1;
}

struct TileParams {
    const int* __restrict__ offsets;    // [nrows + 1]
    const int2* __restrict__ tiles;     // [n_tiles + 1]: (first row, first edge) of tile t; last entry (nrows, E)
    int nrows, n_tiles, nvals;
    const float* a;        // first edge array   (vals / x / alpha); may alias out
    const float* b;        // second edge array  (dalpha)
    float* out;            // edge output (nullptr for the row reductions)
    const float* row_in;   // TILE_SCALE: per-row factor
    float* row_out;        // per-row output (row sum / reciprocal), nullable where optional
    float seed;
};

__device__ __forceinline__ uint32_t tile_smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void tile_bulk_load(void* smem_dst, const void* gsrc, uint32_t bytes, uint64_t* bar) {
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(
                     tile_smem_u32(smem_dst)),
                 "l"(gsrc), "r"(bytes), "r"(tile_smem_u32(bar))
                 : "memory");
}
__device__ __forceinline__ void tile_bulk_store(void* gdst, const void* smem_src, uint32_t bytes) {
    asm volatile("cp.async.bulk.global.shared::cta.bulk_group [%0], [%1], %2;" ::"l"(gdst), "r"(tile_smem_u32(smem_src)),
                 "r"(bytes)
                 : "memory");
}
__device__ __forceinline__ void tile_wait(uint64_t* bar, uint32_t parity) {
    asm volatile(
        "{\n"
        ".reg .pred p;\n"
        "TILE_WAIT:\n"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n"
        "@p bra TILE_DONE;\n"
        "bra TILE_WAIT;\n"
        "TILE_DONE:\n"
        "}\n" ::"r"(tile_smem_u32(bar)),
        "r"(parity)
        : "memory");
}

template <int G>
__device__ __forceinline__ float group_sum(float x, unsigned mask) {
#pragma unroll
    for (int o = G >> 1; o > 0; o >>= 1) x += __shfl_xor_sync(mask, x, o);
    return x;
}

// Sum over the consumer warps in warp order -- all of their threads get the total.
__device__ __forceinline__ float tile_cta_sum(float warp_total) {   // consumer warps only (named barrier 1)
    __shared__ float s_part[kTileConsumers / 32];
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    if (lane == 0) s_part[warp] = warp_total;
    asm volatile("bar.sync 1, %0;" ::"n"(kTileConsumers) : "memory");
    float t = 0.0f;
#pragma unroll
    for (int w = 0; w < kTileConsumers / 32; ++w) t += s_part[w];
    asm volatile("bar.sync 1, %0;" ::"n"(kTileConsumers) : "memory");
    return t;
}

// Visit the edges [e0, e1) with all consumer warps: 128-bit accesses over the 16-byte aligned body, scalar head and tail.
template <class FS, class FV>
__device__ __forceinline__ void cta_edges(int e0, int e1, FS&& scalar, FV&& vec4) {
    const int tid = threadIdx.x;
    const int a0 = min((e0 + 3) & ~3, e1);
    const int a1 = max(a0, e1 & ~3);
    if (e0 + tid < a0) scalar(e0 + tid);
#pragma unroll 2
    for (int e = a0 + tid * 4; e < a1; e += kTileConsumers * 4) vec4(e);
    if (a1 + tid < e1) scalar(a1 + tid);
}

// The per-op arithmetic of the rows streamed from global memory (the shared-memory path inlines the same formulas).
//   pre(a, b)        contribution of one edge to its row's sum; `a` may be replaced by what post() needs
//   post(a, b, tot)  the edge's output given the row total (seed included; its reciprocal for the softmax forward)
template <int OP>
struct TileOp {
    float row_scalar;   // TILE_SCALE: the row's factor
    __device__ __forceinline__ float pre(float& a, float b) const {
        if constexpr (OP == TILE_ROWSUM) return a;
        if constexpr (OP == TILE_SOFTMAX_FWD) { a = softmax_num(a); return a; }
        if constexpr (OP == TILE_SOFTMAX_BWD) return a * b;
        return 0.0f;
    }
    __device__ __forceinline__ float post(float a, float b, float tot) const {
        if constexpr (OP == TILE_SCALE) return a * row_scalar;
        if constexpr (OP == TILE_SOFTMAX_FWD) return a * tot;
        if constexpr (OP == TILE_SOFTMAX_BWD) return a * b - a * tot;
        return a;
    }
};

// What one tile stages: its edge window [a0, a0 + n_win) (a0 16-byte aligned in the arrays) and the row pointers
// [ra0, ra0 + n_off).  The bulk copies move whole 16-byte units; a unit that would reach past the end of an array
// (last tile only) is left to ordinary loads.
struct TileGeom {
    int r_begin, r_end, eb, ee;
    int a0, n_win, n_bulk;        // edge window: staged floats, floats moved by the bulk copy
    int ra0, n_off, n_off_bulk;   // row pointers: staged ints, ints moved by the bulk copy
};
__device__ __forceinline__ TileGeom tile_geom(const TileParams& p, int2 m0, int2 m1) {
    TileGeom g;
    g.r_begin = m0.x; g.eb = m0.y; g.r_end = m1.x; g.ee = m1.y;
    g.a0 = g.eb & ~3;
    g.n_win = min(g.ee, g.a0 + kTileCap) - g.a0;
    const int up = (g.n_win + 3) & ~3;
    g.n_bulk = g.a0 + up <= p.nvals ? up : (g.n_win & ~3);
    g.ra0 = g.r_begin & ~3;
    g.n_off = min(g.r_end + 1, g.ra0 + kTileRowCap) - g.ra0;
    const int upo = (g.n_off + 3) & ~3;
    g.n_off_bulk = g.ra0 + upo <= p.nrows + 1 ? upo : (g.n_off & ~3);
    return g;
}

// First row of the tile whose edges reach past the staged window (offsets[r + 1] > w_end); exists when ee > w_end.
__device__ __forceinline__ int tile_split_row(const int* __restrict__ offsets, const TileGeom& g, int w_end) {
    int lo = g.r_begin, hi = g.r_end - 1;
    while (lo < hi) {
        const int mid = (lo + hi) >> 1;
        if (__ldg(offsets + mid + 1) > w_end) hi = mid;
        else lo = mid + 1;
    }
    return lo;
}

template <int OP, int G>
__global__ void __launch_bounds__(kTileThreads, tile_ctas_per_sm(OP))
edge_tile_kernel(const __grid_constant__ TileParams p) {
    constexpr int NARR = tile_narr(OP);
    constexpr bool kSum = OP != TILE_SCALE;                            // a row reduction precedes the output
    constexpr bool kEdgeOut = OP != TILE_ROWSUM;
    constexpr bool kFlat = OP == TILE_SOFTMAX_FWD || OP == TILE_SOFTMAX_BWD;
    constexpr int SLOTS = kTileConsumers / G;
    constexpr int RB = G >= 32 ? kBatch : 4;                           // loads in flight per lane in the row loops
    extern __shared__ __align__(128) unsigned char s_raw[];            // kTileStages x [NARR windows | row pointers]
    __shared__ __align__(8) uint64_t s_full[kTileStages], s_done[kTileStages];

    const int tid = threadIdx.x, lane = tid & 31;
    auto win = [&](int s) { return reinterpret_cast<float*>(s_raw + (size_t)s * tile_stage_bytes(OP)); };
    auto offs = [&](int s) { return reinterpret_cast<int*>(s_raw + (size_t)s * tile_stage_bytes(OP) + (size_t)NARR * kTileCap * 4); };
    auto meta = [&](int t) { return __ldg(p.tiles + min(t, p.n_tiles)); };
    const int stride = gridDim.x;
    const int first = blockIdx.x;
    if (first >= p.n_tiles) return;
    const int n_my = (p.n_tiles - first + stride - 1) / stride;        // tiles of this CTA: first + k * stride

    if (tid == 0) {
#pragma unroll
        for (int s = 0; s < kTileStages; ++s) {
            asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(tile_smem_u32(&s_full[s])));
            asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(tile_smem_u32(&s_done[s])), "r"(kTileConsumers / 32));
        }
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    __syncthreads();

    if (tid >= kTileConsumers) {
        // ================= producer warp: bulk loads two tiles ahead, bulk stores of the finished tiles =================
        auto issue = [&](int k) {   // lane 0: arm the stage's barrier, start the copies (an empty tile completes the phase)
            const int s = k % kTileStages, t = first + k * stride;
            const TileGeom g = tile_geom(p, meta(t), meta(t + 1));
            const uint32_t bar = tile_smem_u32(&s_full[s]);
            if (g.r_begin >= g.r_end || (g.n_bulk == 0 && g.n_off_bulk == 0)) {
                asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(bar) : "memory");
                return;
            }
            asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar),
                         "r"((uint32_t)((NARR * g.n_bulk + g.n_off_bulk) * 4))
                         : "memory");
            if (g.n_bulk > 0) {
                tile_bulk_load(win(s), p.a + g.a0, g.n_bulk * 4, &s_full[s]);
                if (NARR > 1) tile_bulk_load(win(s) + kTileCap, p.b + g.a0, g.n_bulk * 4, &s_full[s]);
            }
            if (g.n_off_bulk > 0) tile_bulk_load(offs(s), p.offsets + g.ra0, g.n_off_bulk * 4, &s_full[s]);
        };
        if (lane == 0)
            for (int k = 0; k < min(kTileStages, n_my); ++k) issue(k);
        for (int k = 0; k < n_my; ++k) {
            const int s = k % kTileStages, t = first + k * stride;
            const TileGeom g = tile_geom(p, meta(t), meta(t + 1));     // (issued before the wait: off its critical path)
            tile_wait(&s_done[s], (k / kTileStages) & 1);
            if (kEdgeOut && g.r_begin < g.r_end) {
                const int w_end = g.a0 + g.n_win;
                int e_split = g.ee;                                    // edges [eb, e_split) were finished in the window
                if (g.ee > w_end) e_split = __ldg(p.offsets + tile_split_row(p.offsets, g, w_end));
                const float* s_a = win(s);
                const int b0 = min((g.eb + 3) & ~3, e_split), b1 = max(b0, e_split & ~3);
                if (lane == 0 && b1 > b0) {
                    tile_bulk_store(p.out + b0, s_a + (b0 - g.a0), (uint32_t)(b1 - b0) * 4u);
                    asm volatile("cp.async.bulk.commit_group;" ::: "memory");
                }
                if (lane < b0 - g.eb) p.out[g.eb + lane] = s_a[g.eb + lane - g.a0];
                if (lane >= 8 && lane - 8 < e_split - b1) p.out[b1 + lane - 8] = s_a[b1 + lane - 8 - g.a0];
                __syncwarp();
                // the store reads the window asynchronously: it must have done so before the stage is refilled / released
                if (lane == 0) asm volatile("cp.async.bulk.wait_group.read 0;" ::: "memory");
            }
            if (lane == 0 && k + kTileStages < n_my) issue(k + kTileStages);
            __syncwarp();
        }
        return;
    }

    // ================= consumer warps =================
    const int sub = lane % G;
    const unsigned gmask = G == 32 ? kFull : (((1u << G) - 1u) << (lane - sub));
    auto consumer_sync = [&]() { asm volatile("bar.sync 1, %0;" ::"n"(kTileConsumers) : "memory"); };
    TileOp<OP> op{0.0f};
    int2 m0 = meta(first), m1 = meta(first + 1);
    for (int k = 0; k < n_my; ++k) {
        const int s = k % kTileStages, t = first + k * stride;
        const TileGeom g = tile_geom(p, m0, m1);
        m0 = meta(t + stride); m1 = meta(t + stride + 1);              // next tile's descriptor: consumed one iteration on
        tile_wait(&s_full[s], (k / kTileStages) & 1);
        if (g.r_begin >= g.r_end) {                                    // tile inside one long row: no row starts here
            __syncwarp();
            if (lane == 0) asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(tile_smem_u32(&s_done[s])) : "memory");
            continue;
        }
        float* s_a = win(s);
        float* s_b = s_a + (NARR > 1 ? kTileCap : 0);
        int* s_off = offs(s);
        if (g.n_bulk < g.n_win || g.n_off_bulk < g.n_off) {            // last tile: the units the bulk copies left out
            if (tid < g.n_win - g.n_bulk) {
                s_a[g.n_bulk + tid] = p.a[g.a0 + g.n_bulk + tid];
                if (NARR > 1) s_b[g.n_bulk + tid] = p.b[g.a0 + g.n_bulk + tid];
            }
            if (tid >= 32 && tid - 32 < g.n_off - g.n_off_bulk)
                s_off[g.n_off_bulk + tid - 32] = p.offsets[g.ra0 + g.n_off_bulk + tid - 32];
            consumer_sync();                                           // (CTA-uniform branch)
        }
        // ---- flat pass, balanced whatever the rows look like: the per-edge arithmetic that needs no row total ----
        // (loads batched ahead of the arithmetic and the stores: eight shared-memory reads in flight per thread)
        if (kFlat) {
            for (int i = tid; i < g.n_win; i += kBatch * kTileConsumers) {
                float va[kBatch], vb[kBatch];
#pragma unroll
                for (int u = 0; u < kBatch; ++u) {
                    const int j = min(i + u * kTileConsumers, kTileCap - 1);
                    va[u] = s_a[j];
                    if (OP == TILE_SOFTMAX_BWD) vb[u] = s_b[j];
                }
#pragma unroll
                for (int u = 0; u < kBatch; ++u)
                    if (i + u * kTileConsumers < g.n_win) {
                        if (OP == TILE_SOFTMAX_FWD) s_a[i + u * kTileConsumers] = softmax_num(va[u]);
                        else s_b[i + u * kTileConsumers] = va[u] * vb[u];
                    }
            }
            consumer_sync();
        }
        const int a0 = g.a0, w_end = g.a0 + g.n_win;
        auto row_ptr = [&](int r) { return r - g.ra0 < g.n_off ? s_off[r - g.ra0] : __ldg(p.offsets + r); };
        // rows [split, r_end) reach past the window (CTA-uniform; only tiles holding a row longer than the window)
        const int split = g.ee > w_end ? tile_split_row(p.offsets, g, w_end) : g.r_end;

        // ---- rows whose edges lie inside the window: G lanes per row, sums and rescaling in shared memory ----
        for (int r = g.r_begin + tid / G; r < split; r += SLOTS) {
            const int lo = row_ptr(r), hi = row_ptr(r + 1);
            if (OP == TILE_SCALE) op.row_scalar = __ldg(p.row_in + r);
            float tot = 0.0f;
            const int end = hi - a0;
            if (kSum) {
                const float* src = OP == TILE_SOFTMAX_BWD ? s_b : s_a;
                float acc[4] = {0.0f, 0.0f, 0.0f, 0.0f};
                for (int i = lo - a0 + sub; i < end; i += RB * G) {
                    float v[RB];
#pragma unroll
                    for (int u = 0; u < RB; ++u) v[u] = i + u * G < end ? src[i + u * G] : 0.0f;
#pragma unroll
                    for (int u = 0; u < RB; ++u) acc[u & 3] += v[u];
                }
                tot = group_sum<G>((acc[0] + acc[1]) + (acc[2] + acc[3]), gmask) + p.seed;
                if (OP == TILE_SOFTMAX_FWD) tot = 1.0f / tot;
                if ((OP == TILE_ROWSUM || OP == TILE_SOFTMAX_FWD) && sub == 0 && p.row_out) p.row_out[r] = tot;
            }
            if (kEdgeOut) {
                for (int i = lo - a0 + sub; i < end; i += RB * G) {
                    float va[RB], vb[RB];
#pragma unroll
                    for (int u = 0; u < RB; ++u) {
                        const int j = min(i + u * G, kTileCap - 1);
                        va[u] = s_a[j];
                        if (OP == TILE_SOFTMAX_BWD) vb[u] = s_b[j];
                    }
#pragma unroll
                    for (int u = 0; u < RB; ++u)
                        if (i + u * G < end) {
                            if (OP == TILE_SOFTMAX_BWD) s_a[i + u * G] = vb[u] - va[u] * tot;   // alpha*dalpha - alpha*tot
                            else s_a[i + u * G] = op.post(va[u], 0.0f, tot);
                        }
                }
            }
        }
        // hand the stage to the producer warp (bulk store, then the next bulk loads)
        if (kEdgeOut) asm volatile("fence.proxy.async.shared::cta;" ::: "memory");   // generic writes -> bulk store
        __syncwarp();
        if (lane == 0) asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(tile_smem_u32(&s_done[s])) : "memory");

        // ---- rows past the window: all consumer warps per row, two passes over global memory ----
        for (int r = split; r < g.r_end; ++r) {
            const int lo = __ldg(p.offsets + r), hi = __ldg(p.offsets + r + 1);
            if (OP == TILE_SCALE) op.row_scalar = __ldg(p.row_in + r);
            float tot = 0.0f;
            if (kSum) {
                float acc = 0.0f;
                cta_edges(lo, hi,
                          [&](int e) {
                              float a = p.a[e];
                              acc += op.pre(a, NARR > 1 ? p.b[e] : 0.0f);
                          },
                          [&](int e) {
                              float4 a = *reinterpret_cast<const float4*>(p.a + e);
                              float4 b = NARR > 1 ? *reinterpret_cast<const float4*>(p.b + e) : make_float4(0, 0, 0, 0);
                              acc += (op.pre(a.x, b.x) + op.pre(a.y, b.y)) + (op.pre(a.z, b.z) + op.pre(a.w, b.w));
                          });
                tot = tile_cta_sum(warp_sum(acc)) + p.seed;
                if (OP == TILE_SOFTMAX_FWD) tot = 1.0f / tot;
                if ((OP == TILE_ROWSUM || OP == TILE_SOFTMAX_FWD) && tid == 0 && p.row_out) p.row_out[r] = tot;
            }
            auto fin = [&](float a, float b) {
                if (OP == TILE_SOFTMAX_FWD) a = softmax_num(a);
                return op.post(a, b, tot);
            };
            if (kEdgeOut) {
                cta_edges(lo, hi, [&](int e) { p.out[e] = fin(p.a[e], NARR > 1 ? p.b[e] : 0.0f); },
                          [&](int e) {
                              float4 a = *reinterpret_cast<const float4*>(p.a + e);
                              float4 b = NARR > 1 ? *reinterpret_cast<const float4*>(p.b + e) : make_float4(0, 0, 0, 0);
                              float4 o = make_float4(fin(a.x, b.x), fin(a.y, b.y), fin(a.z, b.z), fin(a.w, b.w));
                              *reinterpret_cast<float4*>(p.out + e) = o;
                          });
            }
        }
    }
}

// tiles[t] = (first row r with offsets[r] >= t * kTileEdges, offsets[r]) for t < n_tiles;  tiles[n_tiles] = (nrows, E).
__global__ void __launch_bounds__(256) plan_tiles_kernel(const int* __restrict__ offsets, int nrows, int n_tiles,
                                                         int2* __restrict__ tiles) {
    const int t = blockIdx.x * blockDim.x + threadIdx.x;
    if (t > n_tiles) return;
    int lo = nrows;
    if (t < n_tiles) {
        const int target = t * kTileEdges;  // < E <= INT_MAX
        int hi = nrows;                     // answer in [0, nrows]: offsets[nrows] = E >= target
        lo = 0;
        while (lo < hi) {
            const int mid = (lo + hi) >> 1;
            if (__ldg(offsets + mid) >= target) hi = mid;
            else lo = mid + 1;
        }
    }
    tiles[t] = make_int2(lo, __ldg(offsets + lo));
}

}  // namespace gala
