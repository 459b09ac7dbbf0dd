// edge_ops.cuh -- per-edge kernels of the GAT / normalisation path: row sums of edge
// values (K3), per-row scaling (K4), SDDVV add/mul (K5/K7), SDDMM dot (K6) and the
// fused edge-softmax forward / backward.  HBM-streaming kernels: every edge array is
// read and written once with coalesced 4-byte accesses, row scalars stay in registers.
//
// Same work decomposition as spmm.cuh: one warp per row, hub rows by a whole CTA.
#pragma once
#include "common.cuh"

namespace gala {

struct RowTask {
    int row, lo, hi;
    bool hub, valid;
};

__device__ __forceinline__ RowTask row_task(const GraphDev& g, const int* hub_rows, int n_hub,
                                            int hub_threshold) {
    RowTask t;
    const int warp = threadIdx.x >> 5;
    t.hub = (int)blockIdx.x < n_hub;
    t.valid = true;
    if (t.hub) {
        t.row = __ldg(hub_rows + blockIdx.x);
        int deg = row_degree(g, t.row);
        int per = ((deg + kWarpsPerCta * 32 - 1) / (kWarpsPerCta * 32)) * 32;
        t.lo = warp * per;
        t.hi = min(deg, t.lo + per);
    } else {
        t.row = ((int)blockIdx.x - n_hub) * kWarpsPerCta + warp;
        t.lo = 0;
        t.hi = 0x7fffffff;
        if (t.row >= g.nrows) t.valid = false;
        else if (n_hub > 0 && row_degree(g, t.row) > hub_threshold) t.valid = false;
    }
    return t;
}

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
    const int* __restrict__ hub_rows;
    int n_hub;
    int hub_threshold;
    const float* a;   // per-op meaning, see kernels
    const float* b;
    float* out;
    float seed;       // total seed added to a row sum (S * 1e-12f)
    float slope;
    int op;
    float* out2;
};

// ---- K3: out[row] = seed + sum vals -------------------------------------------------
__global__ void __launch_bounds__(kCtaThreads) edge_rowsum_kernel(const __grid_constant__ EdgeParams p) {
    RowTask t = row_task(p.g, p.hub_rows, p.n_hub, p.hub_threshold);
    if (!t.valid) return;
    const int lane = threadIdx.x & 31;
    float s = 0.0f;
    for_each_chunk(p.g, t.row, t.lo, t.hi, [&](int e0, int e1) {
#pragma unroll 4
        for (int e = e0 + lane; e < e1; e += 32) s += ld_stream(p.a + e);
    });
    s = warp_sum(s);
    if (t.hub) s = cta_sum_ordered(s);
    if (threadIdx.x == (t.hub ? 0 : (threadIdx.x & ~31))) p.out[t.row] = s + p.seed;
}

// ---- K4: vals[e] *= rowval[row] -------------------------------------------------------
__global__ void __launch_bounds__(kCtaThreads) edge_scale_kernel(const __grid_constant__ EdgeParams p) {
    RowTask t = row_task(p.g, p.hub_rows, p.n_hub, p.hub_threshold);
    if (!t.valid) return;
    const int lane = threadIdx.x & 31;
    const float r = __ldg(p.a + t.row);
    for_each_chunk(p.g, t.row, t.lo, t.hi, [&](int e0, int e1) {
#pragma unroll 4
        for (int e = e0 + lane; e < e1; e += 32) p.out[e] = p.out[e] * r;
    });
}

// ---- K5 / K7: out[e] = A[row] (+|*) B[col[e]]  (optional fused LeakyReLU) ---------------
__global__ void __launch_bounds__(kCtaThreads) sddvv_kernel(const __grid_constant__ EdgeParams p) {
    RowTask t = row_task(p.g, p.hub_rows, p.n_hub, p.hub_threshold);
    if (!t.valid) return;
    const int lane = threadIdx.x & 31;
    const float ar = __ldg(p.a + t.row);
    const bool mul = p.op == GALA_SDDVV_MUL;
    const bool act = p.slope != 1.0f;
    for_each_chunk(p.g, t.row, t.lo, t.hi, [&](int e0, int e1) {
#pragma unroll 4
        for (int e = e0 + lane; e < e1; e += 32) {
            float bv = __ldg(p.b + ld_stream(p.g.cols + e));
            float r = mul ? ar * bv : ar + bv;
            if (act) r = leaky(r, p.slope);
            st_stream(p.out + e, r);
        }
    });
}

// ---- edge-softmax forward: alpha = clamp(exp(x)) / (seed + sum_row clamp(exp(x))) --------
__global__ void __launch_bounds__(kCtaThreads) edge_softmax_fwd_kernel(const __grid_constant__ EdgeParams p) {
    RowTask t = row_task(p.g, p.hub_rows, p.n_hub, p.hub_threshold);
    if (!t.valid) return;
    const int lane = threadIdx.x & 31;
    float s = 0.0f;
    for_each_chunk(p.g, t.row, t.lo, t.hi, [&](int e0, int e1) {
#pragma unroll 4
        for (int e = e0 + lane; e < e1; e += 32) s += softmax_num(p.a[e]);
    });
    s = warp_sum(s);
    if (t.hub) s = cta_sum_ordered(s);
    const float r = 1.0f / (s + p.seed);
    if (p.out2 && threadIdx.x == (t.hub ? 0 : (threadIdx.x & ~31))) p.out2[t.row] = r;
    // second pass: the row was just read, so it is served from L1/L2
    for_each_chunk(p.g, t.row, t.lo, t.hi, [&](int e0, int e1) {
#pragma unroll 4
        for (int e = e0 + lane; e < e1; e += 32) p.out[e] = softmax_num(p.a[e]) * r;
    });
}

// ---- edge-softmax backward: out = a*da - a * (seed + sum_row a*da) -----------------------
__global__ void __launch_bounds__(kCtaThreads) edge_softmax_bwd_kernel(const __grid_constant__ EdgeParams p) {
    RowTask t = row_task(p.g, p.hub_rows, p.n_hub, p.hub_threshold);
    if (!t.valid) return;
    const int lane = threadIdx.x & 31;
    float s = 0.0f;
    for_each_chunk(p.g, t.row, t.lo, t.hi, [&](int e0, int e1) {
#pragma unroll 4
        for (int e = e0 + lane; e < e1; e += 32) s += p.a[e] * p.b[e];
    });
    s = warp_sum(s);
    if (t.hub) s = cta_sum_ordered(s);
    const float tot = s + p.seed;
    for_each_chunk(p.g, t.row, t.lo, t.hi, [&](int e0, int e1) {
#pragma unroll 4
        for (int e = e0 + lane; e < e1; e += 32) {
            float al = p.a[e];
            float sds = al * p.b[e];
            p.out[e] = sds - al * tot;
        }
    });
}

// ---- K6: out[e] = dot(A[row,:], B[col[e],:]) -----------------------------------------------
// LPR lanes cover one feature row; the A row lives in registers (ACC*VEC per lane) when
// K <= VEC*LPR*ACC, otherwise the kernel loops over feature tiles (GENERIC).
struct SddmmParams {
    GraphDev g;
    const int* __restrict__ hub_rows;
    int n_hub;
    int hub_threshold;
    const float* __restrict__ A;
    const float* __restrict__ B;
    float* __restrict__ out;
    int K;
};

template <int VEC, int LPR, int ACC>
__global__ void __launch_bounds__(kCtaThreads) sddmm_kernel(const __grid_constant__ SddmmParams p) {
    constexpr int EPI = 32 / LPR;
    constexpr int TW = VEC * LPR * ACC;
    RowTask t = row_task(p.g, p.hub_rows, p.n_hub, p.hub_threshold);
    if (!t.valid) return;
    const int lane = threadIdx.x & 31;
    const int sub = lane % LPR, grp = lane / LPR;
    const int ntiles = (p.K + TW - 1) / TW;

    float areg[ACC][VEC];
    bool fvalid[ACC];
    const float* arow = p.A + (int64_t)t.row * p.K;
#pragma unroll
    for (int a = 0; a < ACC; ++a) {
        const int f = (a * LPR + sub) * VEC;
        fvalid[a] = f < p.K;
        Vec<VEC> x;
#pragma unroll
        for (int v = 0; v < VEC; ++v) x.v[v] = 0.0f;
        if (fvalid[a]) x.load(arow + f);
#pragma unroll
        for (int v = 0; v < VEC; ++v) areg[a][v] = x.v[v];
    }

    for_each_chunk(p.g, t.row, t.lo, t.hi, [&](int e0, int e1) {
        for (int base = e0; base < e1; base += 32) {
            const int idx = base + lane;
            const int c = idx < e1 ? ld_stream(p.g.cols + idx) : -1;
            float mine = 0.0f;
#pragma unroll
            for (int j = 0; j < LPR; ++j) {
                const int cj = __shfl_sync(kFull, c, j * EPI + grp);
                float d = 0.0f;
                if (cj >= 0) {
                    const float* br = p.B + (int64_t)cj * p.K;
#pragma unroll
                    for (int a = 0; a < ACC; ++a) {
                        if (fvalid[a]) {
                            Vec<VEC> x;
                            x.load(br + (a * LPR + sub) * VEC);
#pragma unroll
                            for (int v = 0; v < VEC; ++v) d = fmaf(areg[a][v], x.v[v], d);
                        }
                    }
                    for (int tl = 1; tl < ntiles; ++tl) {  // K > TW: remaining tiles from L1/L2
#pragma unroll
                        for (int a = 0; a < ACC; ++a) {
                            const int f = tl * TW + (a * LPR + sub) * VEC;
                            if (f < p.K) {
                                Vec<VEC> x, y;
                                x.load(br + f);
                                y.load(arow + f);
#pragma unroll
                                for (int v = 0; v < VEC; ++v) d = fmaf(y.v[v], x.v[v], d);
                            }
                        }
                    }
                }
#pragma unroll
                for (int o = 1; o < LPR; o <<= 1) d += __shfl_xor_sync(kFull, d, o);
                // lane L owns edge L of the chunk = iteration L / EPI, group L % EPI
                const float got = __shfl_sync(kFull, d, (lane % EPI) * LPR);
                if (lane / EPI == j) mine = got;
            }
            if (idx < e1) st_stream(p.out + idx, mine);
        }
    });
}

}  // namespace gala
