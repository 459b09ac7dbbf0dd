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
    // W -> shared memory with coalesced loads (row stride K+1: the per-lane column reads below are
    // conflict-free), then each lane keeps its one or two output columns in registers.
    __shared__ float sW[64 * 65];
    for (int i = threadIdx.x; i < p.N * p.K; i += blockDim.x) sW[(i / p.K) * (p.K + 1) + (i % p.K)] = __ldg(p.W + i);
    __syncthreads();
    float w0[KP], w1[KP];
#pragma unroll
    for (int k = 0; k < KP; ++k) {
        w0[k] = (lane < p.N && k < p.K) ? sW[lane * (p.K + 1) + k] : 0.0f;
        w1[k] = (lane + 32 < p.N && k < p.K) ? sW[(lane + 32) * (p.K + 1) + k] : 0.0f;
    }
    const float b0 = (p.bias && lane < p.N) ? __ldg(p.bias + lane) : 0.0f;
    const float b1 = (p.bias && lane + 32 < p.N) ? __ldg(p.bias + lane + 32) : 0.0f;
    const bool two = p.N > 32;
    constexpr int R = 4;   // rows per warp iteration: R independent loads in flight, R independent FMA chains
    for (int64_t row0 = warp_global * R; row0 < p.M; row0 += nwarps * R) {
        float x0[R], x1[R], o0[R], o1[R];
#pragma unroll
        for (int r = 0; r < R; ++r) {
            const int64_t row = row0 + r;
            const float* xr = p.X + row * p.K;
            x0[r] = (row < p.M && lane < p.K) ? ld_stream(xr + lane) : 0.0f;
            x1[r] = 0.0f;
            if (KP > 32) x1[r] = (row < p.M && lane + 32 < p.K) ? ld_stream(xr + lane + 32) : 0.0f;
            o0[r] = b0;
            o1[r] = b1;
        }
#pragma unroll
        for (int k = 0; k < KP; ++k) {
#pragma unroll
            for (int r = 0; r < R; ++r) {
                const float xk = __shfl_sync(kFull, k < 32 ? x0[r] : x1[r], k & 31);
                o0[r] = fmaf(xk, w0[k], o0[r]);
                if (two) o1[r] = fmaf(xk, w1[k], o1[r]);
            }
        }
#pragma unroll
        for (int r = 0; r < R; ++r) {
            const int64_t row = row0 + r;
            if (row >= p.M) break;
            float a = o0[r], c = o1[r];
            if (p.relu) {
                a = fmaxf(a, 0.0f);
                c = fmaxf(c, 0.0f);
            }
            if (p.transpose_out) {   // Y laid out [N, M]: attention projections as two contiguous vectors
                if (lane < p.N) p.Y[(int64_t)lane * p.M + row] = a;
                if (lane + 32 < p.N) p.Y[(int64_t)(lane + 32) * p.M + row] = c;
            } else {
                if (lane < p.N) st_stream(p.Y + row * p.N + lane, a);
                if (lane + 32 < p.N) st_stream(p.Y + row * p.N + lane + 32, c);
            }
        }
    }
}

// N <= 4 (the attention projections): lane l multiplies its input element with W[n, l] and the
// warp reduces -- 5 shuffles per output instead of one broadcast shuffle per input element.
template <int NT>
__global__ void __launch_bounds__(256) linear_tiny_kernel(const __grid_constant__ SmallParams p) {
    const int lane = threadIdx.x & 31;
    const int64_t warp_global = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    const int64_t nwarps = ((int64_t)gridDim.x * blockDim.x) >> 5;
    float w0[NT], w1[NT], b[NT];
#pragma unroll
    for (int n = 0; n < NT; ++n) {
        w0[n] = (n < p.N && lane < p.K) ? __ldg(p.W + n * p.K + lane) : 0.0f;
        w1[n] = (n < p.N && lane + 32 < p.K) ? __ldg(p.W + n * p.K + lane + 32) : 0.0f;
        b[n] = (p.bias && n < p.N) ? __ldg(p.bias + n) : 0.0f;
    }
    constexpr int R = 8;
    for (int64_t row0 = warp_global * R; row0 < p.M; row0 += nwarps * R) {
        float x0[R], x1[R];
#pragma unroll
        for (int r = 0; r < R; ++r) {
            const int64_t row = row0 + r;
            const float* xr = p.X + row * p.K;
            x0[r] = (row < p.M && lane < p.K) ? ld_stream(xr + lane) : 0.0f;
            x1[r] = (row < p.M && lane + 32 < p.K) ? ld_stream(xr + lane + 32) : 0.0f;
        }
#pragma unroll
        for (int r = 0; r < R; ++r) {
            const int64_t row = row0 + r;
#pragma unroll
            for (int n = 0; n < NT; ++n) {
                float s = warp_sum(fmaf(x0[r], w0[n], x1[r] * w1[n])) + b[n];
                if (p.relu) s = fmaxf(s, 0.0f);
                if (lane == 0 && n < p.N && row < p.M) {
                    if (p.transpose_out) p.Y[(int64_t)n * p.M + row] = s;
                    else p.Y[row * p.N + n] = s;
                }
            }
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
    const int64_t warps_needed = (M + 3) / 4;
    const unsigned grid = (unsigned)std::min<int64_t>((warps_needed + 7) / 8, 148 * 2);   // persistent: weights are loaded once per CTA
    if (N <= 2) linear_tiny_kernel<2><<<(unsigned)std::min<int64_t>((M + 63) / 64, 148 * 8), 256, 0, st>>>(p);
    else if (N <= 4) linear_tiny_kernel<4><<<(unsigned)std::min<int64_t>((M + 63) / 64, 148 * 8), 256, 0, st>>>(p);
    else if (K <= 32) linear_small_kernel<32><<<grid, 256, 0, st>>>(p);
    else linear_small_kernel<64><<<grid, 256, 0, st>>>(p);
    cudaError_t e = cudaGetLastError();
    return e == cudaSuccess ? GALA_OK : (int)e;
}
