// plan.cuh -- per-graph load-balance plan, built on the device.
//   hub_rows : rows whose total degree exceeds hub_threshold (one CTA each)
//   row_order: every other row, longest first (bucketed by degree, 4096 buckets), so
//              that the 8 warps of a CTA carry rows of near-equal length and the long
//              rows are scheduled first (LPT order).  The order inside a bucket depends
//              on atomics, but it only decides WHICH warp computes a row, never how a
//              row is summed, so results do not depend on it.
#pragma once
#include "common.cuh"

namespace gala {

constexpr int kPlanBins = 4096;

__device__ __forceinline__ int plan_bin(int deg) { return kPlanBins - 1 - min(deg, kPlanBins - 1); }

__global__ void __launch_bounds__(256) plan_hist_kernel(GraphDev g, int thr, int* hub_count, int* hub_rows,
                                                        int* hist) {
    __shared__ int s_hist[kPlanBins];
    for (int i = threadIdx.x; i < kPlanBins; i += blockDim.x) s_hist[i] = 0;
    __syncthreads();
    for (int row = blockIdx.x * blockDim.x + threadIdx.x; row < g.nrows; row += gridDim.x * blockDim.x) {
        int deg = row_degree(g, row);
        if (deg > thr) hub_rows[atomicAdd(hub_count, 1)] = row;
        else atomicAdd(&s_hist[plan_bin(deg)], 1);
    }
    __syncthreads();
    for (int i = threadIdx.x; i < kPlanBins; i += blockDim.x)
        if (s_hist[i]) atomicAdd(&hist[i], s_hist[i]);
}

// exclusive scan of the 4096 bucket counts -> cursor (single CTA, 1024 threads x 4 bins)
__global__ void __launch_bounds__(1024) plan_scan_kernel(const int* hist, int* cursor) {
    __shared__ int s_sum[1024];
    const int t = threadIdx.x;
    int v[4], run = 0;
#pragma unroll
    for (int i = 0; i < 4; ++i) {
        v[i] = hist[t * 4 + i];
        run += v[i];
    }
    s_sum[t] = run;
    __syncthreads();
    for (int o = 1; o < 1024; o <<= 1) {
        int x = t >= o ? s_sum[t - o] : 0;
        __syncthreads();
        s_sum[t] += x;
        __syncthreads();
    }
    int base = s_sum[t] - run;
#pragma unroll
    for (int i = 0; i < 4; ++i) {
        cursor[t * 4 + i] = base;
        base += v[i];
    }
}

__global__ void __launch_bounds__(256) plan_scatter_kernel(GraphDev g, int thr, int* cursor, int* row_order) {
    __shared__ int s_cnt[kPlanBins];
    __shared__ int s_base[kPlanBins];
    // rows are visited in the same grid-stride pattern as plan_hist_kernel
    for (int i = threadIdx.x; i < kPlanBins; i += blockDim.x) s_cnt[i] = 0;
    __syncthreads();
    for (int row = blockIdx.x * blockDim.x + threadIdx.x; row < g.nrows; row += gridDim.x * blockDim.x) {
        int deg = row_degree(g, row);
        if (deg <= thr) atomicAdd(&s_cnt[plan_bin(deg)], 1);
    }
    __syncthreads();
    for (int i = threadIdx.x; i < kPlanBins; i += blockDim.x) {
        int c = s_cnt[i];
        s_base[i] = c ? atomicAdd(&cursor[i], c) : 0;
        s_cnt[i] = 0;
    }
    __syncthreads();
    for (int row = blockIdx.x * blockDim.x + threadIdx.x; row < g.nrows; row += gridDim.x * blockDim.x) {
        int deg = row_degree(g, row);
        if (deg <= thr) {
            int b = plan_bin(deg);
            row_order[s_base[b] + atomicAdd(&s_cnt[b], 1)] = row;
        }
    }
}

}  // namespace gala
