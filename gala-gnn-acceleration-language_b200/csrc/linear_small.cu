// linear_small.cu -- Y[M,N] = X[M,K] * W[N,K]^T + b for the narrow transforms that follow an
// aggregation in the generated models (classifier Linear(32, classes), the two Linear(h,1)
// attention projections; reference src/codegen/common.h:1185-1281).  K <= 64, N <= 64.
//
// These are pure streaming kernels (30 MB in, <= 38 MB out on the Reddit shape): one warp per
// row, the weights live in registers for the whole kernel (lane n holds W[n,:] and W[n+32,:]),
// the input row is loaded once, coalesced, and broadcast with shuffles; exact fp32 FMA in the
// reference's k order.  cuBLAS spends a GEMM + a bias kernel (~95 us) on the classifier shape.
#include <algorithm>
#include <cstring>

#include "common.cuh"

using namespace gala;

namespace {

struct SmallParams {
    const float* __restrict__ X;
    const float* __restrict__ W;
    const float* __restrict__ bias;
    float* __restrict__ Y;
    int64_t M;
    int K, N, relu, transpose_out;
};

template <int KP>   // KP = K rounded up to 32 or 64
__global__ void __launch_bounds__(256) linear_small_kernel(const __grid_constant__ SmallParams p) {
    const int lane = threadIdx.x & 31;
    const int64_t warp_global = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    const int64_t nwarps = ((int64_t)gridDim.x * blockDim.x) >> 5;
    float w0[KP], w1[KP];
#pragma unroll
    for (int k = 0; k < KP; ++k) {
        w0[k] = (lane < p.N && k < p.K) ? __ldg(p.W + (int64_t)lane * p.K + k) : 0.0f;
        w1[k] = (lane + 32 < p.N && k < p.K) ? __ldg(p.W + (int64_t)(lane + 32) * p.K + k) : 0.0f;
    }
    const float b0 = (p.bias && lane < p.N) ? __ldg(p.bias + lane) : 0.0f;
    const float b1 = (p.bias && lane + 32 < p.N) ? __ldg(p.bias + lane + 32) : 0.0f;
    const bool two = p.N > 32;
    for (int64_t row = warp_global; row < p.M; row += nwarps) {
        const float* xr = p.X + row * p.K;
        float x0 = lane < p.K ? ld_stream(xr + lane) : 0.0f;
        float x1 = 0.0f;
        if (KP > 32) x1 = lane + 32 < p.K ? ld_stream(xr + lane + 32) : 0.0f;
        float o0 = b0, o1 = b1;
#pragma unroll
        for (int k = 0; k < KP; ++k) {
            const float xk = __shfl_sync(kFull, k < 32 ? x0 : x1, k & 31);
            o0 = fmaf(xk, w0[k], o0);
            if (two) o1 = fmaf(xk, w1[k], o1);
        }
        if (p.relu) {
            o0 = fmaxf(o0, 0.0f);
            o1 = fmaxf(o1, 0.0f);
        }
        if (p.transpose_out) {   // Y laid out [N, M]: attention projections as two contiguous vectors
            if (lane < p.N) p.Y[(int64_t)lane * p.M + row] = o0;
            if (lane + 32 < p.N) p.Y[(int64_t)(lane + 32) * p.M + row] = o1;
        } else {
            if (lane < p.N) st_stream(p.Y + row * p.N + lane, o0);
            if (lane + 32 < p.N) st_stream(p.Y + row * p.N + lane + 32, o1);
        }
    }
}

}  // namespace

extern "C" int gala_linear_small_f32(const float* X, int64_t M, int32_t K, const float* W, const float* bias, int32_t N,
                                     float* Y, int32_t relu, int32_t transpose_out, gala_stream_t stream) {
    if (M < 0 || K <= 0 || N <= 0) return GALA_ERR_BAD_SHAPE;
    if (K > 64 || N > 64) return GALA_ERR_UNSUPPORTED;
    if (M == 0) return GALA_OK;
    if (!X || !W || !Y) return GALA_ERR_NULL_POINTER;
    SmallParams p;
    std::memset(&p, 0, sizeof(p));
    p.X = X;
    p.W = W;
    p.bias = bias;
    p.Y = Y;
    p.M = M;
    p.K = K;
    p.N = N;
    p.relu = relu;
    p.transpose_out = transpose_out;
    cudaStream_t st = reinterpret_cast<cudaStream_t>(stream);
    const int64_t warps_needed = M;
    const unsigned grid = (unsigned)std::min<int64_t>((warps_needed + 7) / 8, 148 * 8);
    if (K <= 32) linear_small_kernel<32><<<grid, 256, 0, st>>>(p);
    else linear_small_kernel<64><<<grid, 256, 0, st>>>(p);
    cudaError_t e = cudaGetLastError();
    return e == cudaSuccess ? GALA_OK : (int)e;
}
