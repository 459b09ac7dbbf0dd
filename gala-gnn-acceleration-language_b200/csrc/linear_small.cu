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

// Row-per-thread form of the same transform for N > 4, row-major output (the classifier that follows the last
// aggregation: [233K x 32] x [32 x 41] on the Reddit shape).  The warp-per-row kernel above keeps 4 rows = 512 bytes in
// flight per warp -- latency-bound at ~1.1 TB/s.  Here a warp owns 32 consecutive rows: their K-float inputs are one
// contiguous span (32*K*4 bytes: 8 independent 128-bit loads per lane, all in flight) staged through shared memory,
// lane l then holds row l in registers and walks the N outputs with the weights broadcast from shared memory
// (LDS.128, 4 weights per instruction), and the 32 x N outputs leave through the same staging tile as one contiguous,
// fully coalesced span.  Same arithmetic as above: one accumulator per output, bias first, k ascending.
template <int KP>   // K rounded up to 32 or 64
__global__ void __launch_bounds__(128) linear_rows_kernel(const __grid_constant__ SmallParams p) {
    extern __shared__ __align__(16) float smem[];
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const int N4 = (p.N + 3) & ~3;                  // weight rows, zero-padded to a multiple of 4 outputs
    const int NP = p.N | 1;                          // odd output pitch of the staging tile: conflict-free per-lane writes
    constexpr int XP = KP + 4;                       // input pitch: 16-byte aligned rows, conflict-free 128-bit reads
    const int TS = 32 * (XP > NP ? XP : NP);
    float* sW = smem;                                // [N4][KP]
    float* sB = smem + N4 * KP;                      // [N4]
    float* tile = sB + N4 + warp * TS;
    for (int i = threadIdx.x; i < N4 * KP; i += blockDim.x) {
        const int n = i / KP, k = i % KP;
        sW[i] = (n < p.N && k < p.K) ? __ldg(p.W + n * p.K + k) : 0.0f;
    }
    for (int i = threadIdx.x; i < N4; i += blockDim.x) sB[i] = (p.bias && i < p.N) ? __ldg(p.bias + i) : 0.0f;
    __syncthreads();
    const int64_t ntiles = (p.M + 31) / 32;
    const bool vec_in = p.K == KP && (reinterpret_cast<uintptr_t>(p.X) & 15) == 0;
    for (int64_t t = (int64_t)blockIdx.x * 4 + warp; t < ntiles; t += (int64_t)gridDim.x * 4) {
        const int64_t row0 = t * 32;
        const int rows = (int)(p.M - row0 < 32 ? p.M - row0 : 32);
        if (vec_in) {
            const float4* src = reinterpret_cast<const float4*>(p.X + row0 * KP);
            constexpr int C4 = KP / 4;
#pragma unroll
            for (int j = 0; j < C4; ++j) {           // 32 * C4 float4 of the tile, 32 per step
                const int i = j * 32 + lane;
                const int r = i / C4, c4 = i % C4;
                if (r < rows) *reinterpret_cast<float4*>(tile + r * XP + c4 * 4) = __ldcs(src + i);
            }
        } else {
            const float* src = p.X + row0 * p.K;
            for (int r = 0; r < rows; ++r)
                for (int c = lane; c < KP; c += 32) tile[r * XP + c] = c < p.K ? ld_stream(src + r * p.K + c) : 0.0f;
        }
        __syncwarp();
        float x[KP];
        if (lane < rows) {
#pragma unroll
            for (int c4 = 0; c4 < KP / 4; ++c4) {
                const float4 v = *reinterpret_cast<const float4*>(tile + lane * XP + c4 * 4);
                x[c4 * 4] = v.x; x[c4 * 4 + 1] = v.y; x[c4 * 4 + 2] = v.z; x[c4 * 4 + 3] = v.w;
            }
        } else {
#pragma unroll
            for (int k = 0; k < KP; ++k) x[k] = 0.0f;
        }
        __syncwarp();                                // the tile now takes the outputs
        for (int n0 = 0; n0 < N4; n0 += 4) {
            float o[4];
#pragma unroll
            for (int j = 0; j < 4; ++j) o[j] = sB[n0 + j];
#pragma unroll
            for (int k4 = 0; k4 < KP / 4; ++k4) {
#pragma unroll
                for (int j = 0; j < 4; ++j) {
                    const float4 w = *reinterpret_cast<const float4*>(sW + (n0 + j) * KP + k4 * 4);
                    o[j] = fmaf(x[k4 * 4], w.x, o[j]);
                    o[j] = fmaf(x[k4 * 4 + 1], w.y, o[j]);
                    o[j] = fmaf(x[k4 * 4 + 2], w.z, o[j]);
                    o[j] = fmaf(x[k4 * 4 + 3], w.w, o[j]);
                }
            }
#pragma unroll
            for (int j = 0; j < 4; ++j)
                if (n0 + j < p.N) tile[lane * NP + n0 + j] = p.relu ? fmaxf(o[j], 0.0f) : o[j];
        }
        __syncwarp();
        float* dst = p.Y + row0 * p.N;
        const int total = rows * p.N;
        int r = 0, c = lane;
        for (int i = lane; i < total; i += 32) {
            while (c >= p.N) { c -= p.N; ++r; }
            st_stream(dst + i, tile[r * NP + c]);
            c += 32;
        }
        __syncwarp();                                // before the next tile's inputs overwrite it
    }
}

static size_t rows_kernel_smem(int N, int KP) {
    const int N4 = (N + 3) & ~3, NP = N | 1, XP = KP + 4;
    return (size_t)(N4 * KP + N4 + 4 * 32 * std::max(XP, NP)) * sizeof(float);
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

// The same warp-per-row transform with the row epilogues of gala_linear_f32 (row scale, the two attention
// projections, rows and right-hand attention scalars pushed to every GPU that gathers them).  A LIGHT kernel on
// purpose -- 256-thread CTAs, < 64 registers, a capped grid: the row-block pipeline of the partitioned runners
// (dist_gat) runs it on a side stream NEXT TO the gather kernel, and while its peer stores wait on NVLink it must
// not hold the registers and shared memory the gather needs (the persistent tcgen05 kernel takes 53 K of an SM's
// 64 K registers).
struct SmallExParams {
    SmallParams s;
    const float* __restrict__ row_scale;
    const float* __restrict__ att_w;   // [2, N] or nullptr
    float att_b0, att_b1;
    float* __restrict__ att_out;       // [2, M]
    MultiOut mo, att_mo;
};

template <int KP, bool TWO>   // TWO: N > 32 (a second output column per lane)
__global__ void __launch_bounds__(256) linear_small_ex_kernel(const __grid_constant__ SmallExParams q) {
    const SmallParams& p = q.s;
    const int lane = threadIdx.x & 31;
    const int64_t warp_global = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    const int64_t nwarps = ((int64_t)gridDim.x * blockDim.x) >> 5;
    __shared__ float sW[64 * 65];
    for (int i = threadIdx.x; i < p.N * p.K; i += blockDim.x) sW[(i / p.K) * (p.K + 1) + (i % p.K)] = __ldg(p.W + i);
    __syncthreads();
    float w0[KP], w1[TWO ? KP : 1];
#pragma unroll
    for (int k = 0; k < KP; ++k) {
        w0[k] = (lane < p.N && k < p.K) ? sW[lane * (p.K + 1) + k] : 0.0f;
        if (TWO) w1[k] = (lane + 32 < p.N && k < p.K) ? sW[(lane + 32) * (p.K + 1) + k] : 0.0f;
    }
    const float b0 = (p.bias && lane < p.N) ? __ldg(p.bias + lane) : 0.0f;
    const float b1 = (p.bias && lane + 32 < p.N) ? __ldg(p.bias + lane + 32) : 0.0f;
    float al0 = 0.0f, al1 = 0.0f, ar0 = 0.0f, ar1 = 0.0f;      // this lane's columns of the two attention vectors
    if (q.att_w) {
        if (lane < p.N) { al0 = __ldg(q.att_w + lane); ar0 = __ldg(q.att_w + p.N + lane); }
        if (lane + 32 < p.N) { al1 = __ldg(q.att_w + lane + 32); ar1 = __ldg(q.att_w + p.N + lane + 32); }
    }
    constexpr int R = 4;
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
                if (TWO) o1[r] = fmaf(xk, w1[k], o1[r]);
            }
        }
#pragma unroll
        for (int r = 0; r < R; ++r) {
            const int64_t row = row0 + r;
            if (row >= p.M) break;                                   // (warp-uniform)
            float a = o0[r], c = o1[r];
            if (q.att_w) {   // projections of the pre-activation, unscaled row, as in gala_linear_f32
                const float aL = warp_sum(fmaf(a, al0, c * al1)) + q.att_b0;
                const float aR = warp_sum(fmaf(a, ar0, c * ar1)) + q.att_b1;
                if (lane == 0) {
                    q.att_out[row] = aL;
                    q.att_out[p.M + row] = aR;
                    if (q.att_mo.count > 0) {
                        Vec<1> o;
                        o.v[0] = aR;
                        multi_store<1>(q.att_mo, row, o, row);
                    }
                }
            }
            if (q.row_scale) {
                const float rs = __ldg(q.row_scale + row);
                a *= rs;
                c *= rs;
            }
            if (p.relu) {
                a = fmaxf(a, 0.0f);
                c = fmaxf(c, 0.0f);
            }
            Vec<1> va, vc;
            va.v[0] = a;
            vc.v[0] = c;
            if (q.mo.count > 0) {                                    // 32 lanes x 4 bytes: one 128-byte line per peer
                if (lane < p.N) multi_store<1>(q.mo, row * p.N + lane, va, row);
                if (lane + 32 < p.N) multi_store<1>(q.mo, row * p.N + lane + 32, vc, row);
            } else {
                if (lane < p.N) st_stream(p.Y + row * p.N + lane, a);
                if (lane + 32 < p.N) st_stream(p.Y + row * p.N + lane + 32, c);
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
    else if (!transpose_out) {
        // row-per-thread kernel: 128-thread CTAs, up to 8 per SM (26 KB of shared memory each on the classifier shape)
        const int KP = K <= 32 ? 32 : 64;
        const size_t smem = rows_kernel_smem(N, KP);
        const int64_t tiles = (M + 31) / 32;
        // equal tile counts per warp: as many rounds as one resident wave (7 CTAs per SM at 67 registers) needs
        const int64_t ctas = (tiles + 3) / 4, wave = 148 * 7, rounds = (ctas + wave - 1) / wave;
        const unsigned g2 = (unsigned)((ctas + rounds - 1) / rounds);
        if (KP == 32) {
            cudaFuncSetAttribute(linear_rows_kernel<32>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
            linear_rows_kernel<32><<<g2, 128, smem, st>>>(p);
        } else {
            cudaFuncSetAttribute(linear_rows_kernel<64>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
            linear_rows_kernel<64><<<g2, 128, smem, st>>>(p);
        }
    } else if (K <= 32) linear_small_kernel<32><<<grid, 256, 0, st>>>(p);
    else linear_small_kernel<64><<<grid, 256, 0, st>>>(p);
    cudaError_t e = cudaGetLastError();
    return e == cudaSuccess ? GALA_OK : (int)e;
}


extern "C" int gala_linear_small_ex_f32(const float* X, int64_t M, int32_t K, const float* W, const float* bias, int32_t N,
                                        float* Y, const float* row_scale, int32_t relu, const float* att_w,
                                        const float* att_b, float* att_out, const gala_multi_out_t* multi_out,
                                        const gala_multi_out_t* att_multi_out, int32_t max_ctas, gala_stream_t stream) {
    if (M < 0 || K <= 0 || N <= 0) return GALA_ERR_BAD_SHAPE;
    if (K > 64 || N > 64) return GALA_ERR_UNSUPPORTED;
    if (M == 0) return GALA_OK;
    const bool multi = multi_out && multi_out->count > 0;
    if (!X || !W || (!Y && !multi) || (att_w && (!att_b || !att_out))) return GALA_ERR_NULL_POINTER;
    if (multi && multi_out->count > kMaxPeers) return GALA_ERR_UNSUPPORTED;
    SmallExParams q;
    std::memset(&q, 0, sizeof(q));
    q.s.X = X;
    q.s.W = W;
    q.s.bias = bias;
    q.s.Y = Y;
    q.s.M = M;
    q.s.K = K;
    q.s.N = N;
    q.s.relu = relu;
    q.row_scale = row_scale;
    q.att_w = att_w;
    q.att_out = att_out;
    if (att_w) {
        q.att_b0 = att_b[0];
        q.att_b1 = att_b[1];
    }
    if (multi) {
        q.mo.count = multi_out->count;
        q.mo.mc_base = multi_out->multicast_base;
        q.mo.need = multi_out->need_mask;
        for (int i = 0; i < multi_out->count; ++i) q.mo.base[i] = multi_out->base[i];
    }
    if (att_multi_out && att_multi_out->count > 0) {
        if (!att_w || att_multi_out->count > kMaxPeers) return GALA_ERR_UNSUPPORTED;
        q.att_mo.count = att_multi_out->count;
        q.att_mo.mc_base = att_multi_out->multicast_base;
        q.att_mo.need = att_multi_out->need_mask;
        for (int i = 0; i < att_multi_out->count; ++i) q.att_mo.base[i] = att_multi_out->base[i];
    }
    cudaStream_t st = reinterpret_cast<cudaStream_t>(stream);
    const int64_t warps_needed = (M + 3) / 4;
    int64_t cap = max_ctas > 0 ? max_ctas : 148 * 2;
    const unsigned grid = (unsigned)std::max<int64_t>(1, std::min<int64_t>((warps_needed + 7) / 8, cap));
    if (K <= 32 && N <= 32) linear_small_ex_kernel<32, false><<<grid, 256, 0, st>>>(q);
    else if (K <= 32) linear_small_ex_kernel<32, true><<<grid, 256, 0, st>>>(q);
    else if (N <= 32) linear_small_ex_kernel<64, false><<<grid, 256, 0, st>>>(q);
    else linear_small_ex_kernel<64, true><<<grid, 256, 0, st>>>(q);
    cudaError_t e = cudaGetLastError();
    return e == cudaSuccess ? GALA_OK : (int)e;
}
