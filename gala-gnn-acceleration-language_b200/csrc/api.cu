// api.cu -- the C-ABI (include/gala_b200.h): argument checks, shape dispatch, launches.
#include <algorithm>
#include <cstdio>
#include <cstdlib>
#include <cstring>

#include "edge_ops.cuh"
#include "edge_tiles.cuh"
#include "plan.cuh"
#include "spmm.cuh"

using namespace gala;

namespace {

inline cudaStream_t S(gala_stream_t s) { return reinterpret_cast<cudaStream_t>(s); }

inline bool aligned(const void* p, size_t a) { return (reinterpret_cast<uintptr_t>(p) % a) == 0; }

int check_graph(const gala_graph_t* g) {
    if (!g) return GALA_ERR_NULL_POINTER;
    if (g->nrows < 0 || g->ncols < 0 || g->nvals < 0 || g->segments < 1) return GALA_ERR_BAD_SHAPE;
    if (g->nvals > 0x7fffffffLL) return GALA_ERR_UNSUPPORTED;  // int32 edge ids (common.h:1682)
    if (g->nrows > 0 && !g->offsets) return GALA_ERR_NULL_POINTER;
    if (g->nvals > 0 && !g->cols) return GALA_ERR_NULL_POINTER;
    if (g->segments > 1 && !g->bounds) return GALA_ERR_NULL_POINTER;
    // more segments than the kernel parameters hold: the kernels read the segment starts from a device copy of bounds
    if (g->segments > kMaxSeg && !g->bounds_dev) return GALA_ERR_NULL_POINTER;
    return GALA_OK;
}

GraphDev make_dev(const gala_graph_t* g) {
    GraphDev d;
    d.offsets = g->offsets;
    d.cols = g->cols;
    d.nrows = g->nrows;
    d.S = g->segments;
    d.bounds_dev = g->bounds_dev;
    for (int s = 0; s < kMaxSeg; ++s) d.seg_base[s] = 0;
    if (g->bounds)
        for (int s = 0; s < std::min<int>(g->segments, kMaxSeg); ++s) d.seg_base[s] = g->bounds[2 * s];
    return d;
}

struct HubView {
    const int* rows;
    int n, thr;
    const int* order;
    int n_ordered;
};
HubView hub_of(const gala_plan_t* plan, const gala_graph_t* g) {
    HubView h = {nullptr, 0, 0x7fffffff, nullptr, g->nrows};
    if (!plan) return h;
    if (plan->n_hub > 0 && plan->hub_rows) {
        h.rows = plan->hub_rows;
        h.n = plan->n_hub;
        h.thr = plan->hub_threshold;
    }
    if (plan->row_order && plan->n_ordered + plan->n_hub == g->nrows) {
        h.order = plan->row_order;
        h.n_ordered = plan->n_ordered;
    }
    return h;
}

TaskParams task_of(const HubView& h) {
    TaskParams t;
    t.hub_rows = h.rows;
    t.row_order = h.order;
    t.n_hub = h.n;
    t.n_ordered = h.n_ordered;
    t.hub_threshold = h.thr;
    return t;
}

inline int last_error() {
    cudaError_t e = cudaGetLastError();
    return e == cudaSuccess ? GALA_OK : (int)e;
}

// Pick the vector width / lanes-per-row / accumulators-per-lane for a feature width.
// The vector width follows the alignment of the GATHERED rows only (base pointer and row pitch `ld`): a row of
// K = 41 floats stored with pitch 44 is gathered with 128-bit loads (the last one reads 3 floats of padding).
struct Shape {
    int vec, lpr, acc;
};
Shape shape_for(int K, int vec) {
    Shape s;
    s.vec = vec;
    int units = (K + s.vec - 1) / s.vec;
    s.lpr = 1;
    while (s.lpr < units && s.lpr < 32) s.lpr <<= 1;
    s.acc = units <= 32 ? 1 : units <= 64 ? 2 : 4;
    return s;
}
Shape pick_shape(int K, const void* base, int64_t ld) {
    return shape_for(K, (ld % 4 == 0 && aligned(base, 16)) ? 4 : (ld % 2 == 0 && aligned(base, 8)) ? 2 : 1);
}

#define GALA_SHAPE_SWITCH(SH, CALL)                                        \
    do {                                                                   \
        const int key_ = (SH).vec * 10000 + (SH).lpr * 100 + (SH).acc;     \
        switch (key_) {                                                    \
            case 10101: CALL(1, 1, 1); break;                              \
            case 10201: CALL(1, 2, 1); break;                              \
            case 10401: CALL(1, 4, 1); break;                              \
            case 10801: CALL(1, 8, 1); break;                              \
            case 11601: CALL(1, 16, 1); break;                             \
            case 13201: CALL(1, 32, 1); break;                             \
            case 13202: CALL(1, 32, 2); break;                             \
            case 13204: CALL(1, 32, 4); break;                             \
            case 20101: CALL(2, 1, 1); break;                              \
            case 20201: CALL(2, 2, 1); break;                              \
            case 20401: CALL(2, 4, 1); break;                              \
            case 20801: CALL(2, 8, 1); break;                              \
            case 21601: CALL(2, 16, 1); break;                             \
            case 23201: CALL(2, 32, 1); break;                             \
            case 23202: CALL(2, 32, 2); break;                             \
            case 23204: CALL(2, 32, 4); break;                             \
            case 40101: CALL(4, 1, 1); break;                              \
            case 40201: CALL(4, 2, 1); break;                              \
            case 40401: CALL(4, 4, 1); break;                              \
            case 40801: CALL(4, 8, 1); break;                              \
            case 41601: CALL(4, 16, 1); break;                             \
            case 43201: CALL(4, 32, 1); break;                             \
            case 43202: CALL(4, 32, 2); break;                             \
            case 43204: CALL(4, 32, 4); break;                             \
            default: return GALA_ERR_UNSUPPORTED;                          \
        }                                                                  \
    } while (0)

template <int MODE>
int launch_spmm(const SpmmParams& p, cudaStream_t st) {
    if (p.g.nrows == 0 || p.K == 0) return GALA_OK;
    if (p.ldx < p.K || p.ldy < p.K) return GALA_ERR_BAD_SHAPE;
    Shape sh = pick_shape(p.K, p.X, p.ldx);
    const int tw = sh.vec * sh.lpr * sh.acc;
    dim3 grid(p.t.n_hub + (p.t.n_ordered + kWarpsPerCta - 1) / kWarpsPerCta, (p.K + tw - 1) / tw);
    const bool y16 = !p.Y || (aligned(p.Y, 16) && p.ldy % 4 == 0);
    if (sh.vec == 4 && p.K % tw == 0 && y16) {
        // K is a whole number of tiles: no per-lane feature predicates in the gather loop
#define CALL(V, L, A) spmm_kernel<4, L, A, MODE, true><<<grid, kCtaThreads, 0, st>>>(p)
        switch (sh.lpr * 100 + sh.acc) {
            case 101: CALL(4, 1, 1); break;
            case 201: CALL(4, 2, 1); break;
            case 401: CALL(4, 4, 1); break;
            case 801: CALL(4, 8, 1); break;
            case 1601: CALL(4, 16, 1); break;
            case 3201: CALL(4, 32, 1); break;
            case 3202: CALL(4, 32, 2); break;
            case 3204: CALL(4, 32, 4); break;
            default: return GALA_ERR_UNSUPPORTED;
        }
#undef CALL
        return last_error();
    }
#define CALL(V, L, A) spmm_kernel<V, L, A, MODE, false><<<grid, kCtaThreads, 0, st>>>(p)
    GALA_SHAPE_SWITCH(sh, CALL);
#undef CALL
    return last_error();
}


// bf16 feature rows (Vec<8>): K = 8 * LPR with LPR a power of two <= 32, 16-byte aligned rows
template <int MODE>
int launch_spmm_bf16(const SpmmParams& p, cudaStream_t st) {
    if (p.g.nrows == 0) return GALA_OK;
    if (p.K < 8 || p.K > 256 || (p.K & (p.K - 1)) != 0) return GALA_ERR_UNSUPPORTED;
    if (!aligned(p.X, 16) || !aligned(p.Y, 16)) return GALA_ERR_MISALIGNED;
    dim3 grid(p.t.n_hub + (p.t.n_ordered + kWarpsPerCta - 1) / kWarpsPerCta, 1);
    switch (p.K / 8) {
        case 1: spmm_kernel<8, 1, 1, MODE, true><<<grid, kCtaThreads, 0, st>>>(p); break;
        case 2: spmm_kernel<8, 2, 1, MODE, true><<<grid, kCtaThreads, 0, st>>>(p); break;
        case 4: spmm_kernel<8, 4, 1, MODE, true><<<grid, kCtaThreads, 0, st>>>(p); break;
        case 8: spmm_kernel<8, 8, 1, MODE, true><<<grid, kCtaThreads, 0, st>>>(p); break;
        case 16: spmm_kernel<8, 16, 1, MODE, true><<<grid, kCtaThreads, 0, st>>>(p); break;
        case 32: spmm_kernel<8, 32, 1, MODE, true><<<grid, kCtaThreads, 0, st>>>(p); break;
        default: return GALA_ERR_UNSUPPORTED;
    }
    return last_error();
}

}  // namespace

namespace gala {
// bandwidth probe: grid-stride 128-bit reads that bypass L1 (ld.global.cg), 4 independent loads per thread in flight
__global__ void __launch_bounds__(256) probe_read_kernel(const uint4* __restrict__ buf, int64_t n16, int repeats,
                                                         uint4* __restrict__ sink) {
    const int64_t stride = (int64_t)gridDim.x * blockDim.x;
    uint4 acc = make_uint4(0, 0, 0, 0);
    for (int r = 0; r < repeats; ++r) {
        int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
        for (; i + 3 * stride < n16; i += 4 * stride) {
            uint4 v[4];
#pragma unroll
            for (int u = 0; u < 4; ++u)
                asm volatile("ld.global.cg.v4.u32 {%0,%1,%2,%3}, [%4];"
                             : "=r"(v[u].x), "=r"(v[u].y), "=r"(v[u].z), "=r"(v[u].w)
                             : "l"(buf + i + u * stride));
#pragma unroll
            for (int u = 0; u < 4; ++u) {
                acc.x ^= v[u].x; acc.y ^= v[u].y; acc.z ^= v[u].z; acc.w ^= v[u].w;
            }
        }
        for (; i < n16; i += stride) {
            uint4 v;
            asm volatile("ld.global.cg.v4.u32 {%0,%1,%2,%3}, [%4];" : "=r"(v.x), "=r"(v.y), "=r"(v.z), "=r"(v.w) : "l"(buf + i));
            acc.x ^= v.x; acc.y ^= v.y; acc.z ^= v.z; acc.w ^= v.w;
        }
    }
    if ((acc.x ^ acc.y ^ acc.z ^ acc.w) == 0x9e3779b9u) *sink = acc;   // data-dependent, practically never taken
}
}  // namespace gala

namespace gala {
// Xp[r, 0:K] = X[r, 0:K], Xp[r, K:ld_out] = 0: re-pitch packed rows so that every row starts 16-byte aligned
// (coalesced: consecutive threads write consecutive output elements)
__global__ void __launch_bounds__(256) pad_rows_kernel(const float* __restrict__ X, int64_t nrows, int K, int64_t ld_in,
                                                       float* __restrict__ Xp, int64_t ld_out) {
    const int64_t total = nrows * ld_out;
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (int64_t)gridDim.x * blockDim.x) {
        const int64_t r = i / ld_out;
        const int c = (int)(i - r * ld_out);
        Xp[i] = c < K ? __ldg(X + r * ld_in + c) : 0.0f;
    }
}
}  // namespace gala

// Rows (and one optional scalar per row) copied from a local matrix into every GPU that gathers them: the exchange
// step of a partitioned layer as its own light kernel, for producers whose epilogue cannot write full lines (the
// tcgen05 transform holds one output row per thread: 16-byte pieces of 32 different rows per store instruction) or
// whose rows should travel while the NEXT row block is still being computed.  Every store instruction of a warp
// writes whole 128-byte lines of the destination (K % 4 == 0: K/4 lanes per row, float4 each).
struct PushParams {
    const float* __restrict__ X;
    const float* __restrict__ scalars;
    int64_t M, ldx;
    int K;
    MultiOut mo, smo;
};
__global__ void __launch_bounds__(256) push_rows_kernel(const __grid_constant__ PushParams p) {
    const int vpr = p.K >> 2;                                         // float4 per row
    const int64_t total = p.M * vpr;
    const int64_t stride = (int64_t)gridDim.x * blockDim.x;
    constexpr int U = 4;                                              // independent loads in flight per thread
    for (int64_t i0 = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i0 < total; i0 += stride * U) {
        Vec<4> v[U];
        int64_t row[U];
        int c[U];
#pragma unroll
        for (int u = 0; u < U; ++u) {
            const int64_t i = i0 + u * stride;
            row[u] = i / vpr;
            c[u] = (int)(i - row[u] * vpr);
            if (i < total) v[u].load_rw(p.X + row[u] * p.ldx + c[u] * 4);
        }
#pragma unroll
        for (int u = 0; u < U; ++u)
            if (i0 + u * stride < total) {
                multi_store<4>(p.mo, row[u] * p.K + c[u] * 4, v[u], row[u]);
                if (p.scalars && c[u] == 0 && p.smo.count > 0) {
                    Vec<1> s;
                    s.v[0] = __ldg(p.scalars + row[u]);
                    multi_store<1>(p.smo, row[u], s, row[u]);
                }
            }
    }
}

extern "C" {

int gala_b200_abi_version(void) { return GALA_B200_ABI_VERSION; }

int gala_b200_probe_read(const void* buf, size_t bytes, int32_t repeats, void* sink, gala_stream_t stream) {
    if (!buf || !sink) return GALA_ERR_NULL_POINTER;
    if (!aligned(buf, 16) || repeats < 0) return GALA_ERR_MISALIGNED;
    if (bytes < 16 || repeats == 0) return GALA_OK;
    probe_read_kernel<<<148 * 8, 256, 0, S(stream)>>>(static_cast<const uint4*>(buf), (int64_t)(bytes / 16), repeats,
                                                       static_cast<uint4*>(sink));
    return last_error();
}

const char* gala_b200_error_string(int code) {
    switch (code) {
        case GALA_OK: return "success";
        case GALA_ERR_NULL_POINTER: return "gala_b200: required pointer is NULL";
        case GALA_ERR_BAD_SHAPE: return "gala_b200: negative or inconsistent size";
        case GALA_ERR_UNSUPPORTED: return "gala_b200: unsupported configuration (nvals >= 2^31, width outside the kernel's range, ...)";
        case GALA_ERR_WORKSPACE: return "gala_b200: workspace too small";
        case GALA_ERR_MISALIGNED: return "gala_b200: pointer is not 4-byte aligned";
        default: return code > 0 ? cudaGetErrorString((cudaError_t)code) : "gala_b200: unknown error";
    }
}

static size_t plan_tiles(const gala_graph_t* g) {   // edge tiles of a single-segment graph (0: no table)
    return (g->segments == 1 && g->nvals > 0 && g->nvals <= 0x7fffffffLL) ? (size_t)((g->nvals + kTileEdges - 1) / kTileEdges) : 0;
}

size_t gala_plan_workspace_bytes(const gala_graph_t* g) {
    if (!g || g->nrows < 0) return 0;
    return (2 * (size_t)g->nrows + 4 + 2 * kPlanBins + 2 * (plan_tiles(g) + 1) + 4) * sizeof(int32_t);
}

int gala_plan_build(const gala_graph_t* g, int32_t hub_threshold, void* workspace, size_t workspace_bytes,
                    gala_plan_t* plan, gala_stream_t stream) {
    if (int rc = check_graph(g)) return rc;
    if (!plan || !workspace) return GALA_ERR_NULL_POINTER;
    if (workspace_bytes < gala_plan_workspace_bytes(g)) return GALA_ERR_WORKSPACE;
    if (hub_threshold < 1) return GALA_ERR_BAD_SHAPE;
    int* ws = static_cast<int*>(workspace);
    int* hub_rows = ws + 4;
    int* row_order = hub_rows + g->nrows;
    int* hist = row_order + g->nrows;
    int* cursor = hist + kPlanBins;
    plan->hub_rows = hub_rows;
    plan->row_order = row_order;
    plan->n_hub = 0;
    plan->n_ordered = g->nrows;
    plan->hub_threshold = hub_threshold;
    plan->tile_rows = nullptr;
    plan->n_tiles = 0;
    plan->tile_edges = kTileEdges;
    plan->tile_policy = GALA_TILES_AUTO;
    if (g->nrows == 0) return GALA_OK;
    cudaStream_t st = S(stream);
    if (const size_t nt = plan_tiles(g)) {   // (first row, first edge) of every edge tile: one binary search per tile
        int* tile_rows = cursor + kPlanBins;
        tile_rows += (4 - ((reinterpret_cast<uintptr_t>(tile_rows) / 4) & 3)) & 3;   // 16-byte aligned pairs
        plan_tiles_kernel<<<(unsigned)((nt + 1 + 255) / 256), 256, 0, st>>>(g->offsets, g->nrows, (int)nt,
                                                                            reinterpret_cast<int2*>(tile_rows));
        plan->tile_rows = tile_rows;
        plan->n_tiles = (int)nt;
    }
    cudaError_t e = cudaMemsetAsync(ws, 0, 4 * sizeof(int), st);
    if (e != cudaSuccess) return (int)e;
    e = cudaMemsetAsync(hist, 0, 2 * kPlanBins * sizeof(int), st);
    if (e != cudaSuccess) return (int)e;
    GraphDev d = make_dev(g);
    const int blocks = std::min((g->nrows + 255) / 256, 148 * 8);
    plan_hist_kernel<<<blocks, 256, 0, st>>>(d, hub_threshold, ws, hub_rows, hist);
    plan_scan_kernel<<<1, 1024, 0, st>>>(hist, cursor);
    plan_scatter_kernel<<<blocks, 256, 0, st>>>(d, hub_threshold, cursor, row_order);
    if (int rc = last_error()) return rc;
    int n = 0;
    e = cudaMemcpyAsync(&n, ws, sizeof(int), cudaMemcpyDeviceToHost, st);
    if (e != cudaSuccess) return (int)e;
    e = cudaStreamSynchronize(st);
    if (e != cudaSuccess) return (int)e;
    plan->n_hub = n;
    plan->n_ordered = g->nrows - n;
    return GALA_OK;
}

// Feature matrices larger than this do not stay resident in the 126 MB L2 while the whole
// column range is gathered from; column-tiled graphs are then executed segment by segment.
constexpr int64_t kL2ResidentBytes = 96ll << 20;

int gala_pad_rows_f32(const float* X, int64_t nrows, int32_t K, int64_t ld_in, float* Xp, int64_t ld_out,
                      gala_stream_t stream) {
    if (nrows < 0 || K < 0 || ld_in < K || ld_out < K) return GALA_ERR_BAD_SHAPE;
    if (nrows == 0 || ld_out == 0) return GALA_OK;
    if (!X || !Xp) return GALA_ERR_NULL_POINTER;
    const int64_t total = nrows * ld_out;
    const int blocks = (int)std::min<int64_t>((total + 1023) / 1024, 148 * 16);
    pad_rows_kernel<<<blocks, 256, 0, S(stream)>>>(X, nrows, K, ld_in, Xp, ld_out);
    return last_error();
}

int gala_spmm_f32(const gala_graph_t* g, const float* vals, const float* X, int32_t K, float* Y,
                  const gala_epilogue_t* ep, const gala_plan_t* plan, gala_stream_t stream) {
    if (int rc = check_graph(g)) return rc;
    if (K < 0) return GALA_ERR_BAD_SHAPE;
    const bool pushes = ep && ep->multi_out && ep->multi_out->count > 0;
    if ((g->nrows > 0 && K > 0) && (!X || (!Y && !pushes))) return GALA_ERR_NULL_POINTER;
    if (!aligned(X, 4) || !aligned(Y, 4)) return GALA_ERR_MISALIGNED;
    SpmmParams p;
    std::memset(&p, 0, sizeof(p));
    p.g = make_dev(g);
    p.vals = vals;
    p.X = X;
    p.Y = Y;
    p.K = K;
    p.ldx = p.ldy = K;
    int schedule = 0;
    if (ep) {
        p.row_scale = ep->row_scale;
        p.col_scale = ep->col_scale;
        p.accumulate = ep->accumulate;
        p.relu = ep->relu;
        schedule = ep->schedule;
        if (ep->ldx > 0) p.ldx = ep->ldx;
        if (ep->ldy > 0) p.ldy = ep->ldy;
        if (ep->multi_out && ep->multi_out->count > 0) {
            if (ep->multi_out->count > kMaxPeers) return GALA_ERR_UNSUPPORTED;
            p.mo.count = ep->multi_out->count;
            p.mo.mc_base = ep->multi_out->multicast_base;
            p.mo.need = ep->multi_out->need_mask;
            for (int q = 0; q < p.mo.count; ++q) p.mo.base[q] = ep->multi_out->base[q];
        }
    }
    if (p.ldx < K || p.ldy < K) return GALA_ERR_BAD_SHAPE;
    HubView h = hub_of(plan, g);
    p.t = task_of(h);
    // Segment-major schedule: one launch per column segment, accumulating into Y, so that the slice of
    // X a segment gathers from stays in L2 (GALA's col_tile idea sized for a 126 MB L2).  The reference
    // also launches per segment, but concurrently on separate streams (cuda.h:470-476).
    bool seg_major = g->segments > 1 &&
                     (schedule == GALA_SCHEDULE_SEGMENT_MAJOR ||
                      (schedule == GALA_SCHEDULE_AUTO && (int64_t)g->ncols * K * 4 > kL2ResidentBytes &&
                       g->nvals / std::max<int64_t>(1, (int64_t)g->nrows * g->segments) >= 128));
    // (measured, profiles/r01_tiling_*.txt: segment-major only pays when every row still has >= ~100 edges per
    //  segment -- Reddit shape at K = 602: 21.2 -> 19.1 ms; on the Products shape, mean degree 50, it loses)
    if (p.accumulate && p.row_scale) seg_major = false;   // Y_old must not be scaled: keep the single launch
    if (p.mo.count > 0) seg_major = false;                // rows are pushed once, when they are complete
    // hub rows without a row order: the natural-order fallback decides "hub or not" from the degree it sees, which
    // is per segment in a segment-major launch -- a row could then be run by its hub CTA AND its warp
    if (h.n > 0 && !h.order) seg_major = false;
    if (!seg_major) return launch_spmm<MODE_PLAIN>(p, S(stream));
    const float* row_scale = p.row_scale;
    const int relu = p.relu;
    for (int s = 0; s < g->segments; ++s) {
        const bool last = s == g->segments - 1;
        p.g.offsets = g->offsets + (int64_t)s * (g->nrows + 1);
        p.g.S = 1;
        p.g.seg_base[0] = g->bounds[2 * s];
        p.accumulate = (s > 0) || (ep && ep->accumulate);
        p.row_scale = last ? row_scale : nullptr;
        p.relu = last ? relu : 0;
        p.scale_after = last ? 1 : 0;
        if (int rc = launch_spmm<MODE_PLAIN>(p, S(stream))) return rc;
    }
    return GALA_OK;
}

int gala_gat_forward_f32(const gala_graph_t* g, const float* aL, const float* aR, const float* X, int32_t K,
                         float slope, float* Y, float* alpha_out, int32_t relu, const gala_plan_t* plan,
                         gala_stream_t stream) {
    if (int rc = check_graph(g)) return rc;
    if (K < 0) return GALA_ERR_BAD_SHAPE;
    if (g->nrows > 0 && (!aL || !aR || (K > 0 && (!X || !Y)))) return GALA_ERR_NULL_POINTER;
    if (!aligned(X, 4) || !aligned(Y, 4)) return GALA_ERR_MISALIGNED;
    if (K == 0) return GALA_ERR_UNSUPPORTED;
    SpmmParams p;
    std::memset(&p, 0, sizeof(p));
    p.g = make_dev(g);
    p.X = X;
    p.Y = Y;
    p.K = K;
    p.ldx = p.ldy = K;
    p.relu = relu;
    p.aL = aL;
    p.aR = aR;
    p.slope = slope;
    p.alpha_out = alpha_out;
    p.seed_total = (float)g->segments * 1e-12f;
    HubView h = hub_of(plan, g);
    p.t = task_of(h);
    return launch_spmm<MODE_GAT>(p, S(stream));
}

int gala_spmm_bf16(const gala_graph_t* g, const float* vals, const uint16_t* X, int32_t K, float* Y,
                   const gala_epilogue_t* ep, const gala_plan_t* plan, gala_stream_t stream) {
    if (int rc = check_graph(g)) return rc;
    if (K < 0) return GALA_ERR_BAD_SHAPE;
    if ((g->nrows > 0 && K > 0) && (!X || !Y)) return GALA_ERR_NULL_POINTER;
    SpmmParams p;
    std::memset(&p, 0, sizeof(p));
    p.g = make_dev(g);
    p.vals = vals;
    p.X = reinterpret_cast<const float*>(X);
    p.Y = Y;
    p.K = K;
    p.ldx = p.ldy = K;
    if (ep) {
        p.row_scale = ep->row_scale;
        p.col_scale = ep->col_scale;
        p.accumulate = ep->accumulate;
        p.relu = ep->relu;
    }
    HubView h = hub_of(plan, g);
    p.t = task_of(h);
    return launch_spmm_bf16<MODE_PLAIN>(p, S(stream));
}

int gala_gat_forward_bf16(const gala_graph_t* g, const float* aL, const float* aR, const uint16_t* X, int32_t K,
                          float slope, float* Y, float* alpha_out, int32_t relu, const gala_plan_t* plan,
                          gala_stream_t stream) {
    if (int rc = check_graph(g)) return rc;
    if (K <= 0) return K < 0 ? GALA_ERR_BAD_SHAPE : GALA_ERR_UNSUPPORTED;
    if (g->nrows > 0 && (!aL || !aR || !X || !Y)) return GALA_ERR_NULL_POINTER;
    SpmmParams p;
    std::memset(&p, 0, sizeof(p));
    p.g = make_dev(g);
    p.X = reinterpret_cast<const float*>(X);
    p.Y = Y;
    p.K = K;
    p.ldx = p.ldy = K;
    p.relu = relu;
    p.aL = aL;
    p.aR = aR;
    p.slope = slope;
    p.alpha_out = alpha_out;
    p.seed_total = (float)g->segments * 1e-12f;
    HubView h = hub_of(plan, g);
    p.t = task_of(h);
    return launch_spmm_bf16<MODE_GAT>(p, S(stream));
}

int gala_gat_forward_ex_f32(const gala_graph_t* g, const float* aL, const float* aR, const float* X, int32_t K,
                            float slope, float* Y, float* alpha_out, int32_t relu, const gala_dense_epilogue_t* ep,
                            const gala_plan_t* plan, gala_stream_t stream) {
    if (int rc = check_graph(g)) return rc;
    if (K <= 0) return GALA_ERR_BAD_SHAPE;
    const bool has_ep = ep && (ep->att_w || ep->cls_wT);
    const bool multi = ep && ep->multi_out && ep->multi_out->count > 0;
    if (g->nrows > 0 && (!aL || !aR || !X || (!Y && !(has_ep && ep->cls_wT) && !multi))) return GALA_ERR_NULL_POINTER;
    if (multi && ep->multi_out->count > kMaxPeers) return GALA_ERR_UNSUPPORTED;
    if (!aligned(X, 4) || !aligned(Y, 4)) return GALA_ERR_MISALIGNED;
    SpmmParams p;
    std::memset(&p, 0, sizeof(p));
    p.g = make_dev(g);
    p.X = X;
    p.Y = Y;
    p.K = K;
    p.relu = relu;
    p.aL = aL;
    p.aR = aR;
    p.slope = slope;
    p.alpha_out = alpha_out;
    p.seed_total = (float)g->segments * 1e-12f;
    p.ldx = (ep && ep->ldx > 0) ? ep->ldx : K;
    p.ldy = (ep && ep->ldy > 0) ? ep->ldy : K;
    if (p.ldx < K || p.ldy < K) return GALA_ERR_BAD_SHAPE;
    if (has_ep) {
        if (K > kRowBufMax) return GALA_ERR_UNSUPPORTED;
        if (ep->att_w && !ep->att_out) return GALA_ERR_NULL_POINTER;
        if (ep->cls_wT && (!ep->cls_out || ep->cls_n <= 0)) return GALA_ERR_NULL_POINTER;
        p.att_w = ep->att_w;
        p.att_b0 = ep->att_b[0];
        p.att_b1 = ep->att_b[1];
        p.att_out = ep->att_out;
        p.cls_wT = ep->cls_wT;
        p.cls_b = ep->cls_b;
        p.cls_out = ep->cls_out;
        p.cls_n = ep->cls_n;
    }
    if (multi) {
        p.mo.count = ep->multi_out->count;
        p.mo.mc_base = ep->multi_out->multicast_base;
        p.mo.need = ep->multi_out->need_mask;
        for (int q = 0; q < p.mo.count; ++q) p.mo.base[q] = ep->multi_out->base[q];
    }
    if (ep && ep->att_multi_out && ep->att_multi_out->count > 0) {
        if (!ep->att_w || ep->att_multi_out->count > kMaxPeers) return GALA_ERR_UNSUPPORTED;
        p.att_mo.count = ep->att_multi_out->count;
        p.att_mo.mc_base = ep->att_multi_out->multicast_base;
        p.att_mo.need = ep->att_multi_out->need_mask;
        for (int q = 0; q < p.att_mo.count; ++q) p.att_mo.base[q] = ep->att_multi_out->base[q];
    }
    HubView h = hub_of(plan, g);
    p.t = task_of(h);
    if (has_ep) {   // the dense epilogue needs the whole row in one warp pass: check the tiling the dispatcher picks
        Shape sh = pick_shape(K, X, p.ldx);
        if (sh.vec * sh.lpr * sh.acc < K) return GALA_ERR_UNSUPPORTED;
    }
    return launch_spmm<MODE_GAT>(p, S(stream));
}

int gala_gat_forward_dot_f32(const gala_graph_t* g, const float* aL, const float* wR, float bR, const float* X,
                             int32_t K, float slope, float* Y, float* alpha_out, int32_t relu,
                             const gala_plan_t* plan, gala_stream_t stream) {
    if (int rc = check_graph(g)) return rc;
    if (K <= 0 || K > 32 || K % 4 != 0) return GALA_ERR_UNSUPPORTED;   // one 8-lane pass must hold the row
    if (g->nrows > 0 && (!aL || !wR || !X || !Y)) return GALA_ERR_NULL_POINTER;
    if (!aligned(X, 16) || !aligned(Y, 16)) return GALA_ERR_UNSUPPORTED;
    if (g->nrows == 0) return GALA_OK;
    SpmmParams p;
    std::memset(&p, 0, sizeof(p));
    p.g = make_dev(g);
    p.X = X;
    p.Y = Y;
    p.K = K;
    p.ldx = p.ldy = K;
    p.relu = relu;
    p.aL = aL;
    p.wR = wR;
    p.bR = bR;
    p.slope = slope;
    p.alpha_out = alpha_out;
    p.seed_total = (float)g->segments * 1e-12f;
    HubView h = hub_of(plan, g);
    p.t = task_of(h);
    dim3 grid(p.t.n_hub + (p.t.n_ordered + kWarpsPerCta - 1) / kWarpsPerCta, 1);
    cudaStream_t st = S(stream);
    const int units = K / 4;
    if (units <= 1) spmm_kernel<4, 1, 1, MODE_GAT_DOT, false><<<grid, kCtaThreads, 0, st>>>(p);
    else if (units <= 2) spmm_kernel<4, 2, 1, MODE_GAT_DOT, false><<<grid, kCtaThreads, 0, st>>>(p);
    else if (units <= 4) spmm_kernel<4, 4, 1, MODE_GAT_DOT, false><<<grid, kCtaThreads, 0, st>>>(p);
    else if (K == 32) spmm_kernel<4, 8, 1, MODE_GAT_DOT, true><<<grid, kCtaThreads, 0, st>>>(p);
    else spmm_kernel<4, 8, 1, MODE_GAT_DOT, false><<<grid, kCtaThreads, 0, st>>>(p);
    return last_error();
}

int gala_reflection_f32(const float* w, int32_t K, float* v, float* sR) {
    if (!w || !v || !sR) return GALA_ERR_NULL_POINTER;
    if (K <= 0) return GALA_ERR_BAD_SHAPE;
    double n2 = 0.0;
    for (int i = 0; i < K; ++i) n2 += (double)w[i] * (double)w[i];
    if (!(n2 > 0.0)) return GALA_ERR_BAD_SHAPE;
    const double n = std::sqrt(n2);
    const double sgn = w[K - 1] < 0.0f ? -1.0 : 1.0;
    // v ~ e_{K-1} + sgn * w/|w|: the last component 1 + |w[K-1]|/|w| >= 1, no cancellation
    double vn2 = 0.0;
    std::vector<double> t(K);
    for (int i = 0; i < K; ++i) {
        t[i] = sgn * (double)w[i] / n + (i == K - 1 ? 1.0 : 0.0);
        vn2 += t[i] * t[i];
    }
    const double vn = std::sqrt(vn2);
    for (int i = 0; i < K; ++i) v[i] = (float)(t[i] / vn);
    *sR = (float)(-sgn * n);
    return GALA_OK;
}

int gala_gat_forward_col_f32(const gala_graph_t* g, const float* aL, float sR, float bR, const float* X, int32_t K,
                             float slope, float* Y, float* alpha_out, int32_t relu, const float* reflect_in,
                             const float* reflect_out, const gala_dense_epilogue_t* ep, const gala_plan_t* plan,
                             gala_stream_t stream) {
    if (int rc = check_graph(g)) return rc;
    if (K != 4 && K != 8 && K != 16 && K != 32) return GALA_ERR_UNSUPPORTED;   // K/4 lanes hold the row, last lane the scalar
    const bool has_ep = ep && (ep->att_w || ep->cls_wT);
    const bool multi = ep && ep->multi_out && ep->multi_out->count > 0;
    if (g->nrows > 0 && (!aL || !X || (!Y && !(has_ep && ep->cls_wT) && !multi))) return GALA_ERR_NULL_POINTER;
    if (multi && ep->multi_out->count > kMaxPeers) return GALA_ERR_UNSUPPORTED;
    if (ep && ep->att_multi_out && ep->att_multi_out->count > 0) return GALA_ERR_UNSUPPORTED;   // no scalar to exchange here
    const int64_t ldx = (ep && ep->ldx > 0) ? ep->ldx : K, ldy = (ep && ep->ldy > 0) ? ep->ldy : K;
    if (ldx < K || ldy < K) return GALA_ERR_BAD_SHAPE;
    if (!aligned(X, 16) || !aligned(Y, 16) || ldx % 4 != 0 || ldy % 4 != 0) return GALA_ERR_UNSUPPORTED;
    if (g->nrows == 0) return GALA_OK;
    SpmmParams p;
    std::memset(&p, 0, sizeof(p));
    p.g = make_dev(g);
    p.X = X;
    p.Y = Y;
    p.K = K;
    p.ldx = ldx;
    p.ldy = ldy;
    p.relu = relu;
    p.aL = aL;
    p.sR = sR;
    p.bR = bR;
    p.refl_in = reflect_in;
    p.refl_out = reflect_out;
    p.slope = slope;
    p.alpha_out = alpha_out;
    p.seed_total = (float)g->segments * 1e-12f;
    if (has_ep) {
        if (ep->att_w && !ep->att_out) return GALA_ERR_NULL_POINTER;
        if (ep->cls_wT && (!ep->cls_out || ep->cls_n <= 0)) return GALA_ERR_NULL_POINTER;
        p.att_w = ep->att_w;
        p.att_b0 = ep->att_b[0];
        p.att_b1 = ep->att_b[1];
        p.att_out = ep->att_out;
        p.cls_wT = ep->cls_wT;
        p.cls_b = ep->cls_b;
        p.cls_out = ep->cls_out;
        p.cls_n = ep->cls_n;
    }
    if (multi) {
        p.mo.count = ep->multi_out->count;
        p.mo.mc_base = ep->multi_out->multicast_base;
        p.mo.need = ep->multi_out->need_mask;
        for (int q = 0; q < p.mo.count; ++q) p.mo.base[q] = ep->multi_out->base[q];
    }
    HubView h = hub_of(plan, g);
    p.t = task_of(h);
    dim3 grid(p.t.n_hub + (p.t.n_ordered + kWarpsPerCta - 1) / kWarpsPerCta, 1);
    cudaStream_t st = S(stream);
    switch (K) {
        case 4: spmm_kernel<4, 1, 1, MODE_GAT_COL, true><<<grid, kCtaThreads, 0, st>>>(p); break;
        case 8: spmm_kernel<4, 2, 1, MODE_GAT_COL, true><<<grid, kCtaThreads, 0, st>>>(p); break;
        case 16: spmm_kernel<4, 4, 1, MODE_GAT_COL, true><<<grid, kCtaThreads, 0, st>>>(p); break;
        default: spmm_kernel<4, 8, 1, MODE_GAT_COL, true><<<grid, kCtaThreads, 0, st>>>(p); break;
    }
    return last_error();
}

int gala_spmm_sampled_f32(const gala_graph_t* g, const float* vals, const float* X, int32_t K, float* Y,
                          int32_t nsamples, int32_t ra, int32_t rb, int32_t accumulate, int64_t ldx, int64_t ldy,
                          gala_stream_t stream) {
    if (int rc = check_graph(g)) return rc;
    if (K < 0 || nsamples < 0) return GALA_ERR_BAD_SHAPE;
    if ((g->nrows > 0 && K > 0) && (!X || !Y)) return GALA_ERR_NULL_POINTER;
    if (!aligned(X, 4) || !aligned(Y, 4)) return GALA_ERR_MISALIGNED;
    if (g->nrows == 0 || K == 0) return GALA_OK;
    SampledParams p;
    std::memset(&p, 0, sizeof(p));
    p.g = make_dev(g);
    p.vals = vals;
    p.X = X;
    p.Y = Y;
    p.K = K;
    p.nsamples = nsamples;
    p.ra = ra;
    p.rb = rb;
    p.accumulate = accumulate;
    p.ldx = ldx > 0 ? ldx : K;
    p.ldy = ldy > 0 ? ldy : K;
    if (p.ldx < K || p.ldy < K) return GALA_ERR_BAD_SHAPE;
    Shape sh = pick_shape(K, X, p.ldx);
    const int tw = sh.vec * sh.lpr * sh.acc;
    dim3 grid((g->nrows + kWarpsPerCta - 1) / kWarpsPerCta, (K + tw - 1) / tw);
    cudaStream_t st = S(stream);
#define CALL(V, L, A) spmm_sampled_kernel<V, L, A><<<grid, kCtaThreads, 0, st>>>(p)
    GALA_SHAPE_SWITCH(sh, CALL);
#undef CALL
    return last_error();
}

// Edge-parallel form (edge_tiles.cuh) of a streaming edge kernel: single-segment graph, a plan that carries the tile
// table, 16-byte aligned edge arrays (bulk copies).  Returns false when the row-structured kernel must run instead.
extern "C++" {
template <int OP>
static bool launch_tiles(const gala_graph_t* g, const gala_plan_t* plan, TileParams& tp, cudaStream_t st, int* rc) {
#ifdef GALA_NO_EDGE_TILES
    return false;
#endif
    if (!plan || !plan->tile_rows || plan->tile_edges != kTileEdges || g->segments != 1 || g->nvals <= 0) return false;
    if (plan->n_tiles != (int)((g->nvals + kTileEdges - 1) / kTileEdges)) return false;
    if (!aligned(tp.a, 16) || (tp.b && !aligned(tp.b, 16)) || (tp.out && !aligned(tp.out, 16))) return false;
    if (!aligned(g->offsets, 16) || !aligned(plan->tile_rows, 8)) return false;
    tp.offsets = g->offsets;
    tp.tiles = reinterpret_cast<const int2*>(plan->tile_rows);
    tp.nrows = g->nrows;
    tp.n_tiles = plan->n_tiles;
    tp.nvals = (int)g->nvals;
    const size_t smem = kTileStages * tile_stage_bytes(OP);
    const int per_sm = tile_ctas_per_sm(OP);                 // persistent CTAs: what the shared-memory ring allows
    int dev = 0, sms = 148;
    if (cudaGetDevice(&dev) == cudaSuccess) cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
    const unsigned grid = (unsigned)std::min<int64_t>(plan->n_tiles, (int64_t)sms * per_sm);
    const int64_t deg = g->nvals / std::max(g->nrows, 1);    // lanes per row follow the mean degree
    // Measured on B200 (profiles/r02_edge_tiles.txt): the tile pipeline wins 2-3x on short-row graphs (Products shape,
    // mean degree 50: softmax 1.04 -> 0.44 ms) and for the reduction-free row scaling on every shape (Reddit 0.22 ->
    // 0.17 ms, 81 % of the HBM peak); with rows of hundreds of edges the one-warp-per-row loops inside a tile are the
    // long pole and the row-structured kernels (LPT row order, 64 warps per SM) are as fast or faster.
    if (plan->tile_policy != GALA_TILES_ALWAYS && OP != TILE_SCALE && deg >= 96) return false;
    auto go = [&](auto kern) {
        cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
        if (e == cudaSuccess) {
            kern<<<grid, kTileThreads, smem, st>>>(tp);
            e = cudaGetLastError();
        }
        *rc = e == cudaSuccess ? GALA_OK : (int)e;
    };
    if (deg >= 96) go(edge_tile_kernel<OP, 32>);
    else if (deg >= 32) go(edge_tile_kernel<OP, 16>);
    else if (deg >= 12) go(edge_tile_kernel<OP, 8>);
    else go(edge_tile_kernel<OP, 4>);
    return true;
}
}  // extern "C++"

static int edge_common(const gala_graph_t* g, const gala_plan_t* plan, EdgeParams& p, dim3& grid) {
    if (int rc = check_graph(g)) return rc;
    std::memset(&p, 0, sizeof(p));
    p.g = make_dev(g);
    HubView h = hub_of(plan, g);
    p.t = task_of(h);
    p.slope = 1.0f;
    grid = dim3(h.n + (h.n_ordered + kWarpsPerCta - 1) / kWarpsPerCta);
    return GALA_OK;
}

int gala_edge_rowsum_f32(const gala_graph_t* g, const float* vals, float* out, float seed,
                         const gala_plan_t* plan, gala_stream_t stream) {
    EdgeParams p;
    dim3 grid;
    if (int rc = edge_common(g, plan, p, grid)) return rc;
    if (g->nrows == 0) return GALA_OK;
    if (!out || (g->nvals > 0 && !vals)) return GALA_ERR_NULL_POINTER;
    p.a = vals;
    p.out = out;
    const bool v4 = aligned(vals, 16);
    // reference: `local_C = 1e-12` once per segment, then C += local_C (cuda.h:512-522)
    p.seed = (float)g->segments * seed;
    {
        TileParams tp = {};
        tp.a = vals;
        tp.row_out = out;
        tp.seed = p.seed;
        int rc = GALA_OK;
        if (launch_tiles<TILE_ROWSUM>(g, plan, tp, S(stream), &rc)) return rc;
    }
    if (v4) edge_rowsum_kernel<true><<<grid, kCtaThreads, 0, S(stream)>>>(p);
    else edge_rowsum_kernel<false><<<grid, kCtaThreads, 0, S(stream)>>>(p);
    return last_error();
}

int gala_edge_scale_rows_f32(const gala_graph_t* g, float* vals, const float* rowval, const gala_plan_t* plan,
                             gala_stream_t stream) {
    EdgeParams p;
    dim3 grid;
    if (int rc = edge_common(g, plan, p, grid)) return rc;
    if (g->nrows == 0 || g->nvals == 0) return GALA_OK;
    if (!vals || !rowval) return GALA_ERR_NULL_POINTER;
    p.a = rowval;
    p.out = vals;
    const bool v4 = aligned(vals, 16);
    {
        TileParams tp = {};
        tp.a = vals;
        tp.out = vals;
        tp.row_in = rowval;
        int rc = GALA_OK;
        if (launch_tiles<TILE_SCALE>(g, plan, tp, S(stream), &rc)) return rc;
    }
    if (v4) edge_scale_kernel<true><<<grid, kCtaThreads, 0, S(stream)>>>(p);
    else edge_scale_kernel<false><<<grid, kCtaThreads, 0, S(stream)>>>(p);
    return last_error();
}

int gala_sddvv_f32(const gala_graph_t* g, const float* A, const float* B, float* out, int32_t op,
                   float leaky_slope, const gala_plan_t* plan, gala_stream_t stream) {
    EdgeParams p;
    dim3 grid;
    if (int rc = edge_common(g, plan, p, grid)) return rc;
    if (op != GALA_SDDVV_ADD && op != GALA_SDDVV_MUL) return GALA_ERR_UNSUPPORTED;
    if (g->nrows == 0 || g->nvals == 0) return GALA_OK;
    if (!A || !B || !out) return GALA_ERR_NULL_POINTER;
    p.a = A;
    p.b = B;
    p.out = out;
    p.op = op;
    p.slope = leaky_slope;
    const bool v4 = aligned(g->cols, 16) && aligned(out, 16);
    if (v4) sddvv_kernel<true><<<grid, kCtaThreads, 0, S(stream)>>>(p);
    else sddvv_kernel<false><<<grid, kCtaThreads, 0, S(stream)>>>(p);
    return last_error();
}

int gala_edge_softmax_fwd_f32(const gala_graph_t* g, const float* x, float* alpha, float* recip,
                              const gala_plan_t* plan, gala_stream_t stream) {
    EdgeParams p;
    dim3 grid;
    if (int rc = edge_common(g, plan, p, grid)) return rc;
    if (g->nrows == 0) return GALA_OK;
    if (g->nvals > 0 && (!x || !alpha)) return GALA_ERR_NULL_POINTER;
    p.a = x;
    p.out = alpha;
    p.out2 = recip;
    const bool v4 = aligned(x, 16) && aligned(alpha, 16);
    p.seed = (float)g->segments * 1e-12f;
    {
        TileParams tp = {};
        tp.a = x;
        tp.out = alpha;
        tp.row_out = recip;
        tp.seed = p.seed;
        int rc = GALA_OK;
        if (launch_tiles<TILE_SOFTMAX_FWD>(g, plan, tp, S(stream), &rc)) return rc;
    }
    if (v4) edge_softmax_fwd_kernel<true><<<grid, kCtaThreads, 0, S(stream)>>>(p);
    else edge_softmax_fwd_kernel<false><<<grid, kCtaThreads, 0, S(stream)>>>(p);
    return last_error();
}

int gala_edge_softmax_bwd_f32(const gala_graph_t* g, const float* alpha, const float* dalpha, float* out,
                              const gala_plan_t* plan, gala_stream_t stream) {
    EdgeParams p;
    dim3 grid;
    if (int rc = edge_common(g, plan, p, grid)) return rc;
    if (g->nrows == 0 || g->nvals == 0) return GALA_OK;
    if (!alpha || !dalpha || !out) return GALA_ERR_NULL_POINTER;
    p.a = alpha;
    p.b = dalpha;
    p.out = out;
    p.seed = (float)g->segments * 1e-12f;
    const bool v4 = aligned(alpha, 16) && aligned(dalpha, 16) && aligned(out, 16);
    {
        TileParams tp = {};
        tp.a = alpha;
        tp.b = dalpha;
        tp.out = out;
        tp.seed = p.seed;
        int rc = GALA_OK;
        if (launch_tiles<TILE_SOFTMAX_BWD>(g, plan, tp, S(stream), &rc)) return rc;
    }
    if (v4) edge_softmax_bwd_kernel<true><<<grid, kCtaThreads, 0, S(stream)>>>(p);
    else edge_softmax_bwd_kernel<false><<<grid, kCtaThreads, 0, S(stream)>>>(p);
    return last_error();
}

int gala_gat_backward_att_f32(const gala_graph_t* g, const float* alpha, const float* dalpha, const float* aL,
                              const float* aR, float slope, float* d_att, const gala_plan_t* plan,
                              gala_stream_t stream) {
    GatBwdParams q;
    dim3 grid;
    if (int rc = edge_common(g, plan, q.e, grid)) return rc;
    if (g->nrows == 0) return GALA_OK;
    if (!d_att || !aL || (g->nvals > 0 && (!alpha || !dalpha || !aR))) return GALA_ERR_NULL_POINTER;
    q.e.a = alpha;
    q.e.b = dalpha;
    q.e.out = d_att;
    q.e.seed = (float)g->segments * 1e-12f;
    q.e.slope = slope;
    q.aL = aL;
    q.aR = aR;
    // the 128-bit body also reads the column ids as int4: the segment bases must keep 16-byte phase
    bool v4 = aligned(alpha, 16) && aligned(dalpha, 16) && aligned(g->cols, 16);
    if (v4) gat_bwd_att_kernel<true><<<grid, kCtaThreads, 0, S(stream)>>>(q);
    else gat_bwd_att_kernel<false><<<grid, kCtaThreads, 0, S(stream)>>>(q);
    return last_error();
}

int gala_sddmm_f32(const gala_graph_t* g, const float* A, const float* B, int32_t K, float* out,
                   const gala_plan_t* plan, gala_stream_t stream) {
    if (int rc = check_graph(g)) return rc;
    if (K < 0) return GALA_ERR_BAD_SHAPE;
    if (g->nrows == 0 || g->nvals == 0) return GALA_OK;
    if (!out || (K > 0 && (!A || !B))) return GALA_ERR_NULL_POINTER;
    if (!aligned(A, 4) || !aligned(B, 4)) return GALA_ERR_MISALIGNED;
    SddmmParams p;
    std::memset(&p, 0, sizeof(p));
    p.g = make_dev(g);
    HubView h = hub_of(plan, g);
    p.t = task_of(h);
    p.A = A;
    p.B = B;
    p.out = out;
    p.K = K;
    const bool a16 = aligned(A, 16) && aligned(B, 16) && K % 4 == 0;
    const bool a8 = aligned(A, 8) && aligned(B, 8) && K % 2 == 0;
    Shape sh = shape_for(K > 0 ? K : 1, a16 ? 4 : a8 ? 2 : 1);
    dim3 grid(h.n + (h.n_ordered + kWarpsPerCta - 1) / kWarpsPerCta);
    cudaStream_t st = S(stream);
    if (sh.vec == 4 && K == sh.vec * sh.lpr * sh.acc) {   // one exact tile (hidden widths 4..512): predicate-free loop
#define CALL(V, L, A_) sddmm_kernel<4, L, A_, true><<<grid, kCtaThreads, 0, st>>>(p)
        switch (sh.lpr * 100 + sh.acc) {
            case 101: CALL(4, 1, 1); break;
            case 201: CALL(4, 2, 1); break;
            case 401: CALL(4, 4, 1); break;
            case 801: CALL(4, 8, 1); break;
            case 1601: CALL(4, 16, 1); break;
            case 3201: CALL(4, 32, 1); break;
            case 3202: CALL(4, 32, 2); break;
            case 3204: CALL(4, 32, 4); break;
            default: return GALA_ERR_UNSUPPORTED;
        }
#undef CALL
        return last_error();
    }
#define CALL(V, L, A_) sddmm_kernel<V, L, A_><<<grid, kCtaThreads, 0, st>>>(p)
    GALA_SHAPE_SWITCH(sh, CALL);
#undef CALL
    return last_error();
}

int gala_push_rows_f32(const float* X, int64_t M, int32_t K, int64_t ldx, const float* scalars,
                       const gala_multi_out_t* multi_out, const gala_multi_out_t* scalar_multi_out, int32_t max_ctas,
                       gala_stream_t stream) {
    if (M < 0 || K <= 0) return GALA_ERR_BAD_SHAPE;
    if (M == 0) return GALA_OK;
    if (!X || !multi_out || multi_out->count <= 0) return GALA_ERR_NULL_POINTER;
    if (multi_out->count > kMaxPeers || (K & 3) != 0) return GALA_ERR_UNSUPPORTED;
    PushParams p;
    std::memset(&p, 0, sizeof(p));
    p.X = X;
    p.scalars = scalars;
    p.M = M;
    p.ldx = ldx > 0 ? ldx : K;
    p.K = K;
    if (p.ldx < K || p.ldx % 4 != 0 || !aligned(X, 16)) return GALA_ERR_MISALIGNED;
    p.mo.count = multi_out->count;
    p.mo.mc_base = multi_out->multicast_base;
    p.mo.need = multi_out->need_mask;
    for (int q = 0; q < multi_out->count; ++q) p.mo.base[q] = multi_out->base[q];
    if (scalars && scalar_multi_out && scalar_multi_out->count > 0) {
        if (scalar_multi_out->count > kMaxPeers) return GALA_ERR_UNSUPPORTED;
        p.smo.count = scalar_multi_out->count;
        p.smo.mc_base = scalar_multi_out->multicast_base;
        p.smo.need = scalar_multi_out->need_mask;
        for (int q = 0; q < scalar_multi_out->count; ++q) p.smo.base[q] = scalar_multi_out->base[q];
    }
    const int64_t units = M * (K / 4);
    const int64_t cap = max_ctas > 0 ? max_ctas : 148;
    const unsigned grid = (unsigned)std::max<int64_t>(1, std::min<int64_t>((units + 1023) / 1024, cap));
    push_rows_kernel<<<grid, 256, 0, S(stream)>>>(p);
    return last_error();
}

}  // extern "C"
