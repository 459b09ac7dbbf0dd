// common.cuh -- device-side helpers shared by the sm_100a kernels.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

#include "../../include/gala_b200.h"

namespace gala {

constexpr int kMaxSeg = 64;       // column segments whose start offsets travel in the kernel parameters; graphs with
                                  // more segments read them from gala_graph_t.bounds_dev (see seg_start)
constexpr int kWarpsPerCta = 8;   // 256-thread CTAs everywhere
constexpr int kCtaThreads = kWarpsPerCta * 32;
constexpr unsigned kFull = 0xffffffffu;

// Device view of a column-tiled graph, passed by value in kernel params.
struct GraphDev {
    const int* __restrict__ offsets;  // [S*(N+1)]
    const int* __restrict__ cols;     // [E]
    int nrows;
    int S;
    int seg_base[kMaxSeg];            // bounds[2s] of the segments in this launch (S <= kMaxSeg)
    const int* __restrict__ bounds_dev;   // device copy of bounds[2S], read instead when S > kMaxSeg
};

// Start offset of segment s in cols / vals: from the kernel parameters up to kMaxSeg segments, from the caller's
// device copy of `bounds` beyond (gala_graph_t.bounds_dev; the reference accepts any number of column segments,
// tiling.h:222-283).  No state is carried between segments: the gather kernels sit at their register limit.
__device__ __forceinline__ int seg_start(const GraphDev& g, int s) {
    return g.S <= kMaxSeg ? g.seg_base[s] : __ldg(g.bounds_dev + 2 * s);
}

// ---- cache-hinted loads / stores -------------------------------------------
// Streams that are read exactly once (column indices, edge values) go through
// ld.global.cs / st.global.cs so they are first in line for eviction and leave L1/L2
// to the gathered feature rows, which are re-used.
__device__ __forceinline__ int ld_stream(const int* p) { return __ldcs(p); }
__device__ __forceinline__ float ld_stream(const float* p) { return __ldcs(p); }
__device__ __forceinline__ void st_stream(float* p, float v) { __stcs(p, v); }

// Gathered feature rows.  GALA_ROW_POLICY selects the L1 policy of these loads:
//   0  ld.global.nc (default)          1  ld.global.cg (L2 only)
//   2  ld.global.nc.L1::evict_first     3  ld.global.nc.L1::no_allocate
// (measurements: profiles/r01_variants.txt, profiles/r01_l1_policy_variants.txt)
#ifndef GALA_ROW_POLICY
#define GALA_ROW_POLICY 0
#endif
#if GALA_ROW_POLICY == 2
#define GALA_ROW_QUAL ".L1::evict_first"
#elif GALA_ROW_POLICY == 3
#define GALA_ROW_QUAL ".L1::no_allocate"
#endif
__device__ __forceinline__ float gather_ld(const float* p) {
#if GALA_ROW_POLICY == 0
    return __ldg(p);
#elif GALA_ROW_POLICY == 1
    return __ldcg(p);
#else
    float v;
    asm volatile("ld.global.nc" GALA_ROW_QUAL ".f32 %0, [%1];" : "=f"(v) : "l"(p));
    return v;
#endif
}
__device__ __forceinline__ float2 gather_ld(const float2* p) {
#if GALA_ROW_POLICY == 0
    return __ldg(p);
#elif GALA_ROW_POLICY == 1
    return __ldcg(p);
#else
    float2 v;
    asm volatile("ld.global.nc" GALA_ROW_QUAL ".v2.f32 {%0,%1}, [%2];" : "=f"(v.x), "=f"(v.y) : "l"(p));
    return v;
#endif
}
__device__ __forceinline__ float4 gather_ld(const float4* p) {
#if GALA_ROW_POLICY == 0
    return __ldg(p);
#elif GALA_ROW_POLICY == 1
    return __ldcg(p);
#else
    float4 v;
    asm volatile("ld.global.nc" GALA_ROW_QUAL ".v4.f32 {%0,%1,%2,%3}, [%4];" : "=f"(v.x), "=f"(v.y), "=f"(v.z), "=f"(v.w) : "l"(p));
    return v;
#endif
}
#define GALA_GATHER_LD(p) gather_ld(p)

// Per-column scalars gathered once per edge (aR[col], norm[col]): 4 bytes that cost a whole 32-byte sector
// when they miss.  GALA_SCALAR_POLICY = 1 asks L1 to keep them (ld.global.nc.L1::evict_last).
#ifndef GALA_SCALAR_POLICY
#define GALA_SCALAR_POLICY 0
#endif
__device__ __forceinline__ float ld_keep(const float* p) {
#if GALA_SCALAR_POLICY == 1
    float v;
    asm volatile("ld.global.nc.L1::evict_last.f32 %0, [%1];" : "=f"(v) : "l"(p));
    return v;
#else
    return __ldg(p);
#endif
}

template <int VEC>
struct Vec;
template <>
struct Vec<1> {
    float v[1];
    __device__ __forceinline__ void load(const float* p) { v[0] = GALA_GATHER_LD(p); }
    __device__ __forceinline__ void store(float* p) const { *p = v[0]; }
    __device__ __forceinline__ void load_rw(const float* p) { v[0] = *p; }
};
template <>
struct Vec<2> {
    float v[2];
    __device__ __forceinline__ void load(const float* p) {
        float2 t = GALA_GATHER_LD(reinterpret_cast<const float2*>(p));
        v[0] = t.x; v[1] = t.y;
    }
    __device__ __forceinline__ void store(float* p) const {
        *reinterpret_cast<float2*>(p) = make_float2(v[0], v[1]);
    }
    __device__ __forceinline__ void load_rw(const float* p) {
        float2 t = *reinterpret_cast<const float2*>(p);
        v[0] = t.x; v[1] = t.y;
    }
};
template <>
struct Vec<4> {
    float v[4];
    __device__ __forceinline__ void load(const float* p) {
        float4 t = GALA_GATHER_LD(reinterpret_cast<const float4*>(p));
        v[0] = t.x; v[1] = t.y; v[2] = t.z; v[3] = t.w;
    }
    __device__ __forceinline__ void store(float* p) const {
        *reinterpret_cast<float4*>(p) = make_float4(v[0], v[1], v[2], v[3]);
    }
    __device__ __forceinline__ void load_rw(const float* p) {
        float4 t = *reinterpret_cast<const float4*>(p);
        v[0] = t.x; v[1] = t.y; v[2] = t.z; v[3] = t.w;
    }
};

// Eight features per lane whose SOURCE rows are stored as bf16 (16 bytes per lane, the optional
// reduced-precision feature storage: half the gathered bytes).  Accumulators, outputs and read-modify-write
// stay fp32.  bf16 -> fp32 is exact (a 16-bit shift).
template <>
struct Vec<8> {
    float v[8];
    __device__ __forceinline__ void load(const float* p) {   // p points at 8 bf16 values
        const uint4 t = __ldg(reinterpret_cast<const uint4*>(p));
        v[0] = __uint_as_float(t.x << 16); v[1] = __uint_as_float(t.x & 0xffff0000u);
        v[2] = __uint_as_float(t.y << 16); v[3] = __uint_as_float(t.y & 0xffff0000u);
        v[4] = __uint_as_float(t.z << 16); v[5] = __uint_as_float(t.z & 0xffff0000u);
        v[6] = __uint_as_float(t.w << 16); v[7] = __uint_as_float(t.w & 0xffff0000u);
    }
    __device__ __forceinline__ void store(float* p) const {
        reinterpret_cast<float4*>(p)[0] = make_float4(v[0], v[1], v[2], v[3]);
        reinterpret_cast<float4*>(p)[1] = make_float4(v[4], v[5], v[6], v[7]);
    }
    __device__ __forceinline__ void load_rw(const float* p) {
        const float4 a = reinterpret_cast<const float4*>(p)[0], b = reinterpret_cast<const float4*>(p)[1];
        v[0] = a.x; v[1] = a.y; v[2] = a.z; v[3] = a.w; v[4] = b.x; v[5] = b.y; v[6] = b.z; v[7] = b.w;
    }
};


__device__ __forceinline__ float warp_sum(float x) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) x += __shfl_xor_sync(kFull, x, o);
    return x;
}

// exp -> clamp(0, 1e12) of the reference's edge-softmax (common.h:760-761).
// GALA_FAST_EXP = 0: expf.  1: 2^(x*log2e) on the SFU (ex2.approx, max relative error 2^-22) with the rounding error of
// the product x*log2e carried as a first-order correction, ~3e-7 relative over the whole range that survives the clamp
// (x <= 27.7) -- inside the 1e-5 parity bound, which plain __expf (error growing with |x|) is not guaranteed to be.
#ifndef GALA_FAST_EXP
#define GALA_FAST_EXP 0
#endif
__device__ __forceinline__ float softmax_exp(float x) {
#if GALA_FAST_EXP
    const float t = x * 1.4426950216293335f;                         // hi part of log2(e)
    float r = fmaf(x, 1.4426950216293335f, -t);                      // exact rounding error of the product
    r = fmaf(x, 1.9259629911266175e-8f, r);                          // + x * lo part of log2(e)
    float e;
    asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(e) : "f"(t));
    return fmaf(e, r * 0.6931471805599453f, e);                      // 2^(t+r) = 2^t * (1 + r ln2 + ...)
#else
    return expf(x);
#endif
}
__device__ __forceinline__ float softmax_num(float x) {
    float e = softmax_exp(x);
    return e > 1e12f ? 1e12f : e;
}

__device__ __forceinline__ float leaky(float x, float slope) { return x > 0.0f ? x : x * slope; }

// Visit the pieces of row `row` that fall inside [lo, hi) of the row's edge list
// taken as the concatenation of its per-segment chunks (segment order == column
// order == the order of the untiled CSR row).  f(e0, e1) receives absolute edge
// ranges into cols / vals.
template <class F>
__device__ __forceinline__ void for_each_chunk(const GraphDev& g, int row, int lo, int hi, F&& f) {
    int pos = 0;
#pragma unroll 1
    for (int s = 0; s < g.S; ++s) {
        const int* off = g.offsets + (int64_t)s * (g.nrows + 1) + row;
        int b = __ldg(off), e = __ldg(off + 1);
        int len = e - b;
        int a0 = max(lo - pos, 0), a1 = min(hi - pos, len);
        if (a0 < a1) {
            const int sb = seg_start(g, s);
            f(sb + b + a0, sb + b + a1);
        }
        pos += len;
    }
}

// Visit the edges [e0, e1) of one row with a whole warp: 128-bit accesses over the 16-byte
// aligned body (4 edges per lane per step, 4 steps unrolled = 2 KB in flight per warp),
// scalar accesses on the ragged head and tail.  VEC4 = false keeps everything scalar
// (edge arrays that are not 16-byte aligned).
template <bool VEC4, int UNR = 4, class FS, class FV>
__device__ __forceinline__ void warp_edges(int e0, int e1, int lane, FS&& scalar, FV&& vec4) {
    if (!VEC4) {
#pragma unroll 4
        for (int e = e0 + lane; e < e1; e += 32) scalar(e);
        return;
    }
    const int a0 = min((e0 + 3) & ~3, e1);
    const int a1 = max(a0, e1 & ~3);
    if (e0 + lane < a0) scalar(e0 + lane);
#pragma unroll UNR
    for (int e = a0 + lane * 4; e < a1; e += 128) vec4(e);
    if (a1 + lane < e1) scalar(a1 + lane);
}

__device__ __forceinline__ float4 ld_stream4(const float* p) { return __ldcs(reinterpret_cast<const float4*>(p)); }
__device__ __forceinline__ int4 ld_stream4(const int* p) { return __ldcs(reinterpret_cast<const int4*>(p)); }
__device__ __forceinline__ void st_stream4(float* p, float4 v) { __stcs(reinterpret_cast<float4*>(p), v); }

__device__ __forceinline__ int row_degree(const GraphDev& g, int row) {
    int d = 0;
#pragma unroll 1
    for (int s = 0; s < g.S; ++s) {
        const int* off = g.offsets + (int64_t)s * (g.nrows + 1) + row;
        d += __ldg(off + 1) - __ldg(off);
    }
    return d;
}

// ---- output rows pushed to several GPUs (fused "compute + all-gather") ------------------------
// count == 0: plain local output.  Otherwise the row is stored either once through an NVLS
// multicast address (multimem.st: the NVSwitch replicates it into every GPU's buffer, the local
// one included) or, without multicast support, to each of the `count` peer-mapped bases.
constexpr int kMaxPeers = 8;
struct MultiOut {
    float* base[kMaxPeers];
    float* mc_base;
    const unsigned char* need;   // per-row bit mask (bit q: GPU q gathers this row) or nullptr = every GPU; peer stores only
    int count;
};

template <int VEC>
__device__ __forceinline__ void multimem_store(float* p, const float (&v)[VEC]) {
    if constexpr (VEC == 4)
        asm volatile("multimem.st.relaxed.sys.global.v4.f32 [%0], {%1, %2, %3, %4};" ::"l"(p), "f"(v[0]), "f"(v[1]),
                     "f"(v[2]), "f"(v[3])
                     : "memory");
    else if constexpr (VEC == 2)
        asm volatile("multimem.st.relaxed.sys.global.v2.f32 [%0], {%1, %2};" ::"l"(p), "f"(v[0]), "f"(v[1]) : "memory");
    else
        asm volatile("multimem.st.relaxed.sys.global.f32 [%0], %1;" ::"l"(p), "f"(v[0]) : "memory");
}

// store VEC consecutive floats of output row `row` at element offset `off` of every destination that gathers the row
template <int VEC>
__device__ __forceinline__ void multi_store(const MultiOut& mo, int64_t off, const Vec<VEC>& o, int64_t row) {
    if (mo.mc_base) {
        multimem_store<VEC>(mo.mc_base + off, o.v);
    } else {
        const unsigned need = mo.need ? (unsigned)__ldg(mo.need + row) : 0xffu;
#pragma unroll 1
        for (int q = 0; q < mo.count; ++q)
            if ((need >> q) & 1u) o.store(mo.base[q] + off);
    }
}

// ---- row -> warp / CTA assignment shared by every row-structured kernel -----------
// blockIdx.x <  n_hub : hub row hub_rows[blockIdx.x], its edge list split over the 8 warps
// blockIdx.x >= n_hub : 8 rows per CTA, one per warp, taken from row_order (degree-
//                       descending, hubs excluded) or, without a plan, natural order.
struct TaskParams {
    const int* __restrict__ hub_rows;
    const int* __restrict__ row_order;
    int n_hub;
    int n_ordered;
    int hub_threshold;
};

struct RowTask {
    int row, lo, hi;
    bool hub, valid;
};

__device__ __forceinline__ RowTask row_task(const GraphDev& g, const TaskParams& tp) {
    RowTask t;
    const int warp = threadIdx.x >> 5;
    t.hub = (int)blockIdx.x < tp.n_hub;
    t.valid = true;
    if (t.hub) {
        t.row = __ldg(tp.hub_rows + blockIdx.x);
        int deg = row_degree(g, t.row);
        int per = ((deg + kWarpsPerCta * 32 - 1) / (kWarpsPerCta * 32)) * 32;
        t.lo = warp * per;
        t.hi = min(deg, t.lo + per);
    } else {
        const int slot = ((int)blockIdx.x - tp.n_hub) * kWarpsPerCta + warp;
        t.row = 0;
        t.lo = 0;
        t.hi = 0x7fffffff;
        if (slot >= tp.n_ordered) {
            t.valid = false;
        } else if (tp.row_order) {
            t.row = __ldg(tp.row_order + slot);
        } else {
            t.row = slot;
            if (tp.n_hub > 0 && row_degree(g, t.row) > tp.hub_threshold) t.valid = false;
        }
    }
    return t;
}

}  // namespace gala
