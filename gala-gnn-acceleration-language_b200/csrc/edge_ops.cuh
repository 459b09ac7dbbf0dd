// edge_ops.cuh -- per-edge kernels of the GAT / normalisation path: row sums of edge
// values (K3), per-row scaling (K4), SDDVV add/mul (K5/K7), SDDMM dot (K6) and the
// fused edge-softmax forward / backward.  HBM-streaming kernels: every edge array is
// read and written once with coalesced 4-byte accesses, row scalars stay in registers.
//
// Same work decomposition as spmm.cuh: one warp per row, hub rows by a whole CTA.
#pragma once
#include "common.cuh"

namespace gala {

// Sum over the CTA's warps in warp order (hub rows) -- all threads get the total.
__device__ __forceinline__ float cta_sum_ordered(float warp_total) {
    __shared__ float s_part[kWarpsPerCta];
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    if (lane == 0) s_part[warp] = warp_total;
    __syncthreads();
    float t = 0.0f;
#pragma unroll
    for (int w = 0; w < kWarpsPerCta; ++w) t += s_part[w];
    __syncthreads();
    return t;
}

struct EdgeParams {
    GraphDev g;
    TaskParams t;
    const float* a;   // per-op meaning, see kernels
    const float* b;
    float* out;
    float seed;       // total seed added to a row sum (S * 1e-12f)
    float slope;
    int op;
    float* out2;
};

// ---- K3: out[row] = seed + sum vals -------------------------------------------------
template <bool V4>
__global__ void __launch_bounds__(kCtaThreads) edge_rowsum_kernel(const __grid_constant__ EdgeParams p) {
    RowTask t = row_task(p.g, p.t);
    if (!t.valid) return;
    const int lane = threadIdx.x & 31;
    float s = 0.0f;
#ifndef GALA_ROWSUM_UNROLL
#define GALA_ROWSUM_UNROLL 4     // 128-bit loads in flight per lane; measured on B200: 4 -> 0.111 ms, 8 -> 0.117, 16 -> 0.122
#endif
    for_each_chunk(p.g, t.row, t.lo, t.hi, [&](int e0, int e1) {
        warp_edges<V4, GALA_ROWSUM_UNROLL>(e0, e1, lane, [&](int e) { s += ld_stream(p.a + e); },
                       [&](int e) {
                           float4 v = ld_stream4(p.a + e);
                           s += (v.x + v.y) + (v.z + v.w);
                       });
    });
    s = warp_sum(s);
    if (t.hub) s = cta_sum_ordered(s);
    if (threadIdx.x == (t.hub ? 0 : (threadIdx.x & ~31))) p.out[t.row] = s + p.seed;
}

// ---- K4: vals[e] *= rowval[row] -------------------------------------------------------
template <bool V4>
__global__ void __launch_bounds__(kCtaThreads) edge_scale_kernel(const __grid_constant__ EdgeParams p) {
    RowTask t = row_task(p.g, p.t);
    if (!t.valid) return;
    const int lane = threadIdx.x & 31;
    const float r = __ldg(p.a + t.row);
    for_each_chunk(p.g, t.row, t.lo, t.hi, [&](int e0, int e1) {
        warp_edges<V4>(e0, e1, lane, [&](int e) { p.out[e] = p.out[e] * r; },
                       [&](int e) {
                           float4 v = *reinterpret_cast<const float4*>(p.out + e);
                           v.x *= r; v.y *= r; v.z *= r; v.w *= r;
                           *reinterpret_cast<float4*>(p.out + e) = v;
                       });
    });
}

// ---- K5 / K7: out[e] = A[row] (+|*) B[col[e]]  (optional fused LeakyReLU) ---------------
template <bool V4>
__global__ void __launch_bounds__(kCtaThreads) sddvv_kernel(const __grid_constant__ EdgeParams p) {
    RowTask t = row_task(p.g, p.t);
    if (!t.valid) return;
    const int lane = threadIdx.x & 31;
    const float ar = __ldg(p.a + t.row);
    const bool mul = p.op == GALA_SDDVV_MUL;
    const bool act = p.slope != 1.0f;
    auto f = [&](float bv) {
        float r = mul ? ar * bv : ar + bv;
        return act ? leaky(r, p.slope) : r;
    };
    for_each_chunk(p.g, t.row, t.lo, t.hi, [&](int e0, int e1) {
        warp_edges<V4>(e0, e1, lane, [&](int e) { st_stream(p.out + e, f(ld_keep(p.b + ld_stream(p.g.cols + e)))); },
                       [&](int e) {
                           int4 c = ld_stream4(p.g.cols + e);
                           float4 o;
                           o.x = ld_keep(p.b + c.x); o.y = ld_keep(p.b + c.y); o.z = ld_keep(p.b + c.z); o.w = ld_keep(p.b + c.w);
                           o.x = f(o.x); o.y = f(o.y); o.z = f(o.z); o.w = f(o.w);
                           st_stream4(p.out + e, o);
                       });
    });
}

// ---- edge-softmax forward: alpha = clamp(exp(x)) / (seed + sum_row clamp(exp(x))) --------
// (Measured alternatives, profiles/r02_variants_dot_pipeline_softmax_cache.txt, Reddit shape: keeping a warp's share of
//  the row in registers between the two passes -- one read, one exp per edge -- 0.29 ms; stashing the numerators in the
//  output and rescaling -- 0.26 ms; this two-pass form, second read from L2, exp evaluated twice -- 0.25 ms.)
template <bool V4>
__global__ void __launch_bounds__(kCtaThreads) edge_softmax_fwd_kernel(const __grid_constant__ EdgeParams p) {
    RowTask t = row_task(p.g, p.t);
    if (!t.valid) return;
    const int lane = threadIdx.x & 31;
    float s = 0.0f;
    for_each_chunk(p.g, t.row, t.lo, t.hi, [&](int e0, int e1) {
        warp_edges<V4>(e0, e1, lane, [&](int e) { s += softmax_num(p.a[e]); },
                       [&](int e) {
                           float4 v = *reinterpret_cast<const float4*>(p.a + e);
                           s += (softmax_num(v.x) + softmax_num(v.y)) + (softmax_num(v.z) + softmax_num(v.w));
                       });
    });
    s = warp_sum(s);
    if (t.hub) s = cta_sum_ordered(s);
    const float r = 1.0f / (s + p.seed);
    if (p.out2 && threadIdx.x == (t.hub ? 0 : (threadIdx.x & ~31))) p.out2[t.row] = r;
    // second pass: the row was just read, so it is served from L1/L2
    for_each_chunk(p.g, t.row, t.lo, t.hi, [&](int e0, int e1) {
        warp_edges<V4>(e0, e1, lane, [&](int e) { p.out[e] = softmax_num(p.a[e]) * r; },
                       [&](int e) {
                           float4 v = *reinterpret_cast<const float4*>(p.a + e);
                           v.x = softmax_num(v.x) * r; v.y = softmax_num(v.y) * r;
                           v.z = softmax_num(v.z) * r; v.w = softmax_num(v.w) * r;
                           *reinterpret_cast<float4*>(p.out + e) = v;
                       });
    });
}

// ---- edge-softmax backward: out = a*da - a * (seed + sum_row a*da) -----------------------
template <bool V4>
__global__ void __launch_bounds__(kCtaThreads) edge_softmax_bwd_kernel(const __grid_constant__ EdgeParams p) {
    RowTask t = row_task(p.g, p.t);
    if (!t.valid) return;
    const int lane = threadIdx.x & 31;
    float s = 0.0f;
    for_each_chunk(p.g, t.row, t.lo, t.hi, [&](int e0, int e1) {
        warp_edges<V4>(e0, e1, lane, [&](int e) { s += p.a[e] * p.b[e]; },
                       [&](int e) {
                           float4 a = *reinterpret_cast<const float4*>(p.a + e);
                           float4 b = *reinterpret_cast<const float4*>(p.b + e);
                           s += (a.x * b.x + a.y * b.y) + (a.z * b.z + a.w * b.w);
                       });
    });
    s = warp_sum(s);
    if (t.hub) s = cta_sum_ordered(s);
    const float tot = s + p.seed;
    for_each_chunk(p.g, t.row, t.lo, t.hi, [&](int e0, int e1) {
        warp_edges<V4>(e0, e1, lane,
                       [&](int e) {
                           float al = p.a[e];
                           p.out[e] = al * p.b[e] - al * tot;
                       },
                       [&](int e) {
                           float4 a = *reinterpret_cast<const float4*>(p.a + e);
                           float4 b = *reinterpret_cast<const float4*>(p.b + e);
                           float4 o;
                           o.x = a.x * b.x - a.x * tot; o.y = a.y * b.y - a.y * tot;
                           o.z = a.z * b.z - a.z * tot; o.w = a.w * b.w - a.w * tot;
                           *reinterpret_cast<float4*>(p.out + e) = o;
                       });
    });
}

// ---- GAT attention backward, edge side in one kernel ----------------------------------------
// What autograd runs for one GAT layer between d(alpha) and d(attenL)/d(attenR) in the generated
// program (common.h:791-799 softmax backward, LeakyReLU backward, common.h:630-675 edge-sum backward):
//   sds = alpha*dalpha;  ds = sds - alpha*(seed + sum_row sds);  de = ds * (aL[row]+aR[col] > 0 ? 1 : slope)
//   out[row] = seed + sum_row de            (the reference returns this row sum for BOTH attenL and attenR)
// e.a = alpha, e.b = dalpha, e.out = d_att; nothing of size E is written.
struct GatBwdParams {
    EdgeParams e;
    const float* __restrict__ aL;
    const float* __restrict__ aR;
};

template <bool V4>
__global__ void __launch_bounds__(kCtaThreads) gat_bwd_att_kernel(const __grid_constant__ GatBwdParams q) {
    const EdgeParams& p = q.e;
    RowTask t = row_task(p.g, p.t);
    if (!t.valid) return;
    const int lane = threadIdx.x & 31;
    float s = 0.0f;
    for_each_chunk(p.g, t.row, t.lo, t.hi, [&](int e0, int e1) {
        warp_edges<V4>(e0, e1, lane, [&](int e) { s += p.a[e] * p.b[e]; },
                       [&](int e) {
                           float4 a = *reinterpret_cast<const float4*>(p.a + e);
                           float4 b = *reinterpret_cast<const float4*>(p.b + e);
                           s += (a.x * b.x + a.y * b.y) + (a.z * b.z + a.w * b.w);
                       });
    });
    s = warp_sum(s);
    if (t.hub) s = cta_sum_ordered(s);
    const float tot = s + p.seed;
    const float al = __ldg(q.aL + t.row);
    const float slope = p.slope;
    float r = 0.0f;
    auto one = [&](float a, float b, int c) {
        const float ds = a * b - a * tot;
        return (al + ld_keep(q.aR + c)) > 0.0f ? ds : ds * slope;
    };
    for_each_chunk(p.g, t.row, t.lo, t.hi, [&](int e0, int e1) {
        warp_edges<V4>(e0, e1, lane, [&](int e) { r += one(p.a[e], p.b[e], ld_stream(p.g.cols + e)); },
                       [&](int e) {
                           float4 a = *reinterpret_cast<const float4*>(p.a + e);
                           float4 b = *reinterpret_cast<const float4*>(p.b + e);
                           int4 c = *reinterpret_cast<const int4*>(p.g.cols + e);
                           r += (one(a.x, b.x, c.x) + one(a.y, b.y, c.y)) + (one(a.z, b.z, c.z) + one(a.w, b.w, c.w));
                       });
    });
    r = warp_sum(r);
    if (t.hub) r = cta_sum_ordered(r);
    if (threadIdx.x == (t.hub ? 0 : (threadIdx.x & ~31))) p.out[t.row] = r + p.seed;
}

// ---- K6: out[e] = dot(A[row,:], B[col[e],:]) -----------------------------------------------
// LPR lanes cover one feature row; the A row lives in registers (ACC*VEC per lane) when
// K <= VEC*LPR*ACC, further feature tiles are re-read from L1.  Each lane keeps one partial
// dot product per edge of its group (LPR of them per 32-edge chunk); the partials are then
// reduced ACROSS the LPR lanes by a halving exchange (LPR-1 shuffles for LPR edges instead
// of LPR*log2(LPR)), which leaves lane `sub` holding the finished dot product of edge
// `sub` of its group; the 32 results go out as one coalesced 128-byte store.
struct SddmmParams {
    GraphDev g;
    TaskParams t;
    const float* __restrict__ A;
    const float* __restrict__ B;
    float* __restrict__ out;
    int K;
};

// EXACT: K == VEC*LPR*ACC (one feature tile, every lane's features valid): no per-lane predicates, no
// zero-fill of the gathered vectors, no remainder-tile loop.
template <int VEC, int LPR, int ACC, bool EXACT = false>
__global__ void __launch_bounds__(kCtaThreads) sddmm_kernel(const __grid_constant__ SddmmParams p) {
    constexpr int EPI = 32 / LPR;
    constexpr int TW = VEC * LPR * ACC;
    RowTask t = row_task(p.g, p.t);
    if (!t.valid) return;
    const int lane = threadIdx.x & 31;
    const int sub = lane % LPR, grp = lane / LPR;
    const int ntiles = EXACT ? 1 : (p.K + TW - 1) / TW;
    const uint32_t row_bytes = (uint32_t)p.K * 4u;

    float areg[ACC][VEC];
    bool fvalid[ACC];
    const float* arow = p.A + (int64_t)t.row * p.K;
#pragma unroll
    for (int a = 0; a < ACC; ++a) {
        const int f = (a * LPR + sub) * VEC;
        fvalid[a] = EXACT || f < p.K;
        Vec<VEC> x;
#pragma unroll
        for (int v = 0; v < VEC; ++v) x.v[v] = 0.0f;
        if (fvalid[a]) x.load(arow + f);
#pragma unroll
        for (int v = 0; v < VEC; ++v) areg[a][v] = x.v[v];
    }
    // lanes past K read the row start (A is zero there, so the product vanishes)
    const char* blane = reinterpret_cast<const char*>(p.B + ((EXACT || sub * VEC < p.K) ? sub * VEC : 0));
    // after the halving exchange lane `sub` owns edge j = sub of its group, i.e. chunk
    // position sub*EPI + grp
    const int my_pos = sub * EPI + grp;

    for_each_chunk(p.g, t.row, t.lo, t.hi, [&](int e0, int e1) {
        int c_nxt = (e0 + lane < e1) ? ld_stream(p.g.cols + e0 + lane) : -1;
        for (int base = e0; base < e1; base += 32) {
            const int c = c_nxt;
            if (base + 32 < e1) c_nxt = (base + 32 + lane < e1) ? ld_stream(p.g.cols + base + 32 + lane) : -1;
            float d[LPR];
            constexpr int UNR = LPR < 8 ? LPR : 8;
#pragma unroll
            for (int j0 = 0; j0 < LPR; j0 += UNR) {
                Vec<VEC> x[UNR][ACC];
#pragma unroll
                for (int u = 0; u < UNR; ++u) {
                    const int cj = __shfl_sync(kFull, c, (j0 + u) * EPI + grp);
                    const char* br = blane + (uint64_t)(uint32_t)max(cj, 0) * row_bytes;
#pragma unroll
                    for (int a = 0; a < ACC; ++a) {
                        if (EXACT) {
                            x[u][a].load(reinterpret_cast<const float*>(br) + a * LPR * VEC);
                        } else {
#pragma unroll
                            for (int v = 0; v < VEC; ++v) x[u][a].v[v] = 0.0f;
                            if (fvalid[a]) x[u][a].load(reinterpret_cast<const float*>(br) + a * LPR * VEC);
                        }
                    }
                }
#pragma unroll
                for (int u = 0; u < UNR; ++u) {
                    float s = 0.0f;
#pragma unroll
                    for (int a = 0; a < ACC; ++a)
#pragma unroll
                        for (int v = 0; v < VEC; ++v) s = fmaf(areg[a][v], x[u][a].v[v], s);
                    d[j0 + u] = s;
                }
            }
            if (!EXACT && ntiles > 1) {  // K > TW: remaining feature tiles (A row from L1)
#pragma unroll 1
                for (int j = 0; j < LPR; ++j) {
                    const int cj = __shfl_sync(kFull, c, j * EPI + grp);
                    const float* br = p.B + (uint64_t)(uint32_t)max(cj, 0) * (uint32_t)p.K;
                    float s = 0.0f;
                    for (int tl = 1; tl < ntiles; ++tl) {
#pragma unroll
                        for (int a = 0; a < ACC; ++a) {
                            const int f = tl * TW + (a * LPR + sub) * VEC;
                            if (f < p.K) {
                                Vec<VEC> x, y;
                                x.load(br + f);
                                y.load(arow + f);
#pragma unroll
                                for (int v = 0; v < VEC; ++v) s = fmaf(y.v[v], x.v[v], s);
                            }
                        }
                    }
#pragma unroll
                    for (int jj = 0; jj < LPR; ++jj)
                        if (jj == j) d[jj] += s;
                }
            }
            // halving exchange: after the step with offset o each lane keeps the half of its
            // partials whose edge index has bit o equal to the lane's own bit o
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
            if (base + my_pos < e1) st_stream(p.out + base + my_pos, d[0]);
        }
    });
}

}  // namespace gala
