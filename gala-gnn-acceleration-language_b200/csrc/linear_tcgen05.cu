// linear_tcgen05.cu -- the dense feature transform Y = X * W^T + b of the generated models
// (torch::nn::Linear, reference src/codegen/common.h:1185-1281) on the 5th-generation tensor
// cores: tcgen05.mma kind::tf32 with the accumulator in TMEM, error-compensated to fp32 accuracy
// ("3xTF32": x = hi + lo with hi = the 19 significant bits the tensor core keeps,
//  X*W ~= Xhi*Whi + Xhi*Wlo + Xlo*Whi; the dropped Xlo*Wlo term is 2^-22 relative).
//
// Why a hand-written kernel: [233K x 602] * [602 x 32] is 9 GFLOP over 561 MB of X -- 86 us at
// the HBM roofline, out of reach of the fp32 SIMT pipes (cuBLAS' fp32 path takes ~240 us on
// B200) and plain TF32 misses the 1e-5 parity bound.  The tensor pipe makes it HBM-bound.
//
// Shape of the kernel (one CTA = 128 rows of X, all N <= 64 output columns):
//   warps 0-3  producers: coalesced 128-byte row loads (X's row pitch K*4 is not a multiple of
//              16 bytes for K = 602, so TMA cannot describe it), hi/lo split in registers,
//              stores into the 128B-swizzled K-major layout UMMA expects; later the epilogue
//              (tcgen05.ld: one accumulator row per thread, bias / ReLU / attention projections).
//   warp 4     one elected thread issues 3 x 4 tcgen05.mma per 32-wide K chunk and commits to
//              the stage's "empty" mbarrier; TMEM allocation / release.
//   2-stage shared-memory ring (A hi/lo 2 x 16 KB + B hi/lo 2 x NPAD*128 B per stage), several
//   CTAs per SM so that one CTA's epilogue overlaps another's loads.
#include <algorithm>
#include <cstring>

#include "common.cuh"

using namespace gala;

namespace {

constexpr int kBM = 128;          // rows per CTA = UMMA_M
constexpr int kBK = 32;           // fp32 per K chunk = one 128-byte swizzle row
constexpr int kStages = 2;
constexpr int kProducerThreads = 128;
constexpr int kThreads = 160;     // 4 producer/epilogue warps + 1 MMA warp

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }

__device__ __forceinline__ void mbar_init(uint64_t* bar, uint32_t count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count));
}
__device__ __forceinline__ void mbar_arrive(uint64_t* bar) {
    asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ void mbar_wait(uint64_t* bar, uint32_t parity) {
    asm volatile(
        "{\n"
        ".reg .pred p;\n"
        "WAIT_LOOP:\n"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n"
        "@p bra DONE;\n"
        "bra WAIT_LOOP;\n"
        "DONE:\n"
        "}\n" ::"r"(smem_u32(bar)),
        "r"(parity)
        : "memory");
}

// K-major, SWIZZLE_128B shared-memory matrix descriptor (cute::UMMA::SmemDescriptor layout):
// start address >> 4 in [0,14), LBO >> 4 in [16,30) (unused for swizzled K-major), SBO >> 4 in
// [32,46) = 1024 B between 8-row groups, version 1 in [46,48), layout type 2 (128B) in [61,64).
__device__ __forceinline__ uint64_t make_desc(uint32_t smem_addr) {
    uint64_t d = 0;
    d |= (uint64_t)((smem_addr & 0x3FFFF) >> 4);
    d |= (uint64_t)(1024 >> 4) << 32;
    d |= (uint64_t)1 << 46;
    d |= (uint64_t)2 << 61;
    return d;
}

// kind::tf32 instruction descriptor (cute::UMMA::InstrDescriptor): D fp32, A/B tf32, both K-major.
__host__ __device__ constexpr uint32_t make_idesc(int n) {
    return (1u << 4) | (2u << 7) | (2u << 10) | ((uint32_t)(n >> 3) << 17) | ((uint32_t)(kBM >> 4) << 24);
}

__device__ __forceinline__ void umma_tf32(uint32_t tmem_d, uint64_t a_desc, uint64_t b_desc, uint32_t idesc, uint32_t acc) {
    asm volatile(
        "{\n"
        ".reg .pred p;\n"
        "setp.ne.b32 p, %4, 0;\n"
        "tcgen05.mma.cta_group::1.kind::tf32 [%0], %1, %2, %3, p;\n"
        "}\n" ::"r"(tmem_d),
        "l"(a_desc), "l"(b_desc), "r"(idesc), "r"(acc)
        : "memory");
}

__device__ __forceinline__ void umma_commit(uint64_t* bar) {
    asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(bar))
                 : "memory");
}

// byte offset of element (row r, 4-byte column c) inside a [rows x 32 fp32] K-major SW128 tile
[[maybe_unused]] __device__ __forceinline__ uint32_t sw128(int r, int c) {   // reference form of the swizzle used below
    return (uint32_t)(r * 128 + ((((c >> 2) ^ (r & 7)) << 4) | ((c & 3) << 2)));
}

__device__ __forceinline__ void split_tf32(float x, float& hi, float& lo) {
    hi = __uint_as_float(__float_as_uint(x) & 0xFFFFE000u);   // exactly representable in tf32
    lo = x - hi;                                              // exact in fp32; tf32 keeps its top bits
}

struct LinearParams {
    const float* __restrict__ X;      // [M, K]
    const float* __restrict__ W;      // [N, K]
    const float* __restrict__ bias;   // [N] or nullptr
    float* __restrict__ Y;            // [M, N]
    const float* __restrict__ row_scale;  // [M] or nullptr: Y[r,:] *= row_scale[r] (before ReLU)
    const float* __restrict__ att_w;  // [2, N] or nullptr
    const float* __restrict__ att_b_dev;  // [2] on the device, or nullptr: att_b0 / att_b1 below
    float att_b0, att_b1;
    float* __restrict__ att_out;      // [2, M]
    int64_t M;
    int K, N, relu;
    int round_robin;                  // v2 kernel: tiles dealt round-robin (1) or one contiguous row range per CTA (0)
    MultiOut mo;                      // count > 0: output rows pushed to every GPU instead of Y
    MultiOut att_mo;                  // count > 0: the second projection (attenR, one float per row) is ALSO stored
                                      // at element `row` of every GPU's gathered vector
};

// NPAD <= 64: the shapes of the GNN layers (hidden / class widths), 2 CTAs per SM, fused row epilogues.
// NPAD in (64, 256] ("wide": the 172-class classifier of the Papers shape): the same pipeline with one MMA covering all
// NPAD columns (UMMA_N up to 256), fewer accumulators (TMEM has 512 columns), W loaded in row blocks, and an epilogue
// that walks the accumulator row 16 columns at a time instead of holding it in registers.
template <int NPAD>
__global__ void __launch_bounds__(kThreads, (NPAD > 64 ? 1 : 2)) linear_tf32x3_kernel(const __grid_constant__ LinearParams p) {
    constexpr uint32_t kABytes = kBM * 128;            // one A tile (hi or lo)
    constexpr uint32_t kBBytes = NPAD * 128;           // one B tile (hi or lo)
    constexpr uint32_t kStageBytes = 2 * kABytes + 2 * kBBytes;
    // The tensor core truncates when it aligns addends into the fp32 accumulator, so one long
    // accumulation chain drifts (4e-6 relative after 228 MMAs, measured).  The hi*hi products are
    // therefore spread round-robin over kMain accumulators, the small correction terms go to their
    // own accumulator, and the epilogue adds the kMain + 1 partials with IEEE fp32 adds.
    constexpr bool kWide = NPAD > 64;
    constexpr int kMain = NPAD > 128 ? 1 : NPAD > 64 ? 2 : NPAD == 64 ? 3 : 4;
    constexpr uint32_t kAccCols = (kMain + 1) * NPAD;
    constexpr uint32_t kTmemCols = kAccCols <= 32 ? 32 : kAccCols <= 64 ? 64 : kAccCols <= 128 ? 128 : kAccCols <= 256 ? 256 : 512;
    static_assert(kAccCols <= 512, "TMEM has 512 columns");
    extern __shared__ __align__(1024) uint8_t smem_raw[];
    uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
    __shared__ __align__(8) uint64_t full_bar[kStages], empty_bar[kStages], acc_bar;
    __shared__ uint32_t tmem_base_smem;

    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    const int64_t row0 = (int64_t)blockIdx.x * kBM;
    const int nchunks = (p.K + kBK - 1) / kBK;

    if (tid == 0) {
        for (int s = 0; s < kStages; ++s) {
            mbar_init(&full_bar[s], kProducerThreads);
            mbar_init(&empty_bar[s], 1);
        }
        mbar_init(&acc_bar, 1);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    if (warp == 4) {   // TMEM allocation is warp-collective
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(&tmem_base_smem)),
                     "r"(kTmemCols)
                     : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
    }
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    const uint32_t tmem_base = tmem_base_smem;

    if (warp < 4) {
        // ---------------- producers: 4 warps x 32 rows of the tile each, one row per step -------------
        // Software-pipelined: the 32 + NPAD/4 global loads of chunk ch+1 are issued before chunk ch is
        // split and stored, so every producer warp keeps two chunks (8 KB) of HBM requests in flight.
        constexpr int kWRows = NPAD / 4;
        constexpr int kWReg = kWide ? 1 : kWRows;     // wide: W is not staged in registers (row blocks at store time)
        float va[32], wa[kWReg], vb[32], wb[kWReg];
        // Addressing is hoisted so that the steady state costs one IMAD + LDG per element on the load
        // side and LOP3 + FADD + 2 STS (immediate offsets) on the store side.
        const int64_t wrow0 = row0 + warp * 32;
        const char* xlane = reinterpret_cast<const char*>(p.X + wrow0 * p.K + lane);
        const char* wlane = reinterpret_cast<const char*>(p.W + (int64_t)(warp * kWRows) * p.K + lane);
        const uint32_t pitch = (uint32_t)p.K * 4u;
        const bool rows_full = wrow0 + 32 <= p.M;
        const bool wrows_full = warp * kWRows + kWRows <= p.N;
        uint32_t sw[8];   // swizzled 16-byte chunk of this lane's column for row-in-atom j
#pragma unroll
        for (int j = 0; j < 8; ++j) sw[j] = (uint32_t)((((lane >> 2) ^ j) << 4) | ((lane & 3) << 2));
        auto load_chunk = [&](int ch, float (&v)[32], float (&w)[kWReg]) {
            const bool kok = ch * kBK + lane < p.K;
            const char* xc = xlane + ch * (kBK * 4);
            const char* wc = wlane + ch * (kBK * 4);
            if (rows_full && kok) {
#pragma unroll
                for (int i = 0; i < 32; ++i) v[i] = __ldg(reinterpret_cast<const float*>(xc + (uint64_t)i * pitch));
            } else {
#pragma unroll
                for (int i = 0; i < 32; ++i)
                    v[i] = (kok && wrow0 + i < p.M) ? __ldg(reinterpret_cast<const float*>(xc + (uint64_t)i * pitch)) : 0.0f;
            }
            if constexpr (!kWide) {
                if (wrows_full && kok) {
#pragma unroll
                    for (int i = 0; i < kWRows; ++i) w[i] = __ldg(reinterpret_cast<const float*>(wc + (uint64_t)i * pitch));
                } else {
#pragma unroll
                    for (int i = 0; i < kWRows; ++i)   // rows >= N are zero
                        w[i] = (kok && warp * kWRows + i < p.N) ? __ldg(reinterpret_cast<const float*>(wc + (uint64_t)i * pitch)) : 0.0f;
                }
            }
        };
        auto store_chunk = [&](int ch, const float (&v)[32], const float (&w)[kWReg]) {
            const int s = ch % kStages;
            const uint32_t ph = (uint32_t)((ch / kStages) & 1);
            if (ch >= kStages) mbar_wait(&empty_bar[s], ph ^ 1);
            uint8_t* a_hi = smem + (size_t)s * kStageBytes + warp * 4096;
            uint8_t* a_lo = a_hi + kABytes;
            uint8_t* b_hi = smem + (size_t)s * kStageBytes + 2 * kABytes + warp * (kWRows * 128);
            uint8_t* b_lo = b_hi + kBBytes;
#pragma unroll
            for (int i = 0; i < 32; ++i) {
                float hi, lo;
                split_tf32(v[i], hi, lo);
                *reinterpret_cast<float*>(a_hi + i * 128 + sw[i & 7]) = hi;
                *reinterpret_cast<float*>(a_lo + i * 128 + sw[i & 7]) = lo;
            }
            if constexpr (!kWide) {
#pragma unroll
                for (int i = 0; i < kWRows; ++i) {   // warp * kWRows is a multiple of 4; row-in-atom = (warp*kWRows + i) & 7
                    float hi, lo;
                    split_tf32(w[i], hi, lo);
                    const uint32_t o = (uint32_t)(i * 128) + sw[(warp * kWRows + i) & 7];
                    *reinterpret_cast<float*>(b_hi + o) = hi;
                    *reinterpret_cast<float*>(b_lo + o) = lo;
                }
            } else {
                // W (L2-resident, N*K*4 bytes) in blocks of 8 rows, loaded right here: kWRows is a multiple of 4
                const bool kok = ch * kBK + lane < p.K;
                const char* wc = wlane + ch * (kBK * 4);
#pragma unroll 1
                for (int i0 = 0; i0 < kWRows; i0 += 8) {
                    float w8[8];
#pragma unroll
                    for (int i = 0; i < 8; ++i)
                        w8[i] = (i0 + i < kWRows && kok && warp * kWRows + i0 + i < p.N)
                                    ? __ldg(reinterpret_cast<const float*>(wc + (uint64_t)(i0 + i) * pitch)) : 0.0f;
#pragma unroll
                    for (int i = 0; i < 8; ++i) {
                        if (i0 + i < kWRows) {
                            float hi, lo;
                            split_tf32(w8[i], hi, lo);
                            const uint32_t o = (uint32_t)((i0 + i) * 128) + sw[(warp * kWRows + i0 + i) & 7];
                            *reinterpret_cast<float*>(b_hi + o) = hi;
                            *reinterpret_cast<float*>(b_lo + o) = lo;
                        }
                    }
                }
            }
            asm volatile("fence.proxy.async.shared::cta;" ::: "memory");   // generic-proxy stores -> async proxy (UMMA)
            mbar_arrive(&full_bar[s]);
        };
        load_chunk(0, va, wa);
        for (int ch = 0; ch < nchunks; ch += 2) {
            if (ch + 1 < nchunks) load_chunk(ch + 1, vb, wb);
            store_chunk(ch, va, wa);
            if (ch + 1 < nchunks) {
                if (ch + 2 < nchunks) load_chunk(ch + 2, va, wa);
                store_chunk(ch + 1, vb, wb);
            }
        }
        // ---------------- epilogue: thread t owns accumulator row t (TMEM lane t) ----------------------
        mbar_wait(&acc_bar, 0);
        asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
        const int64_t r = row0 + tid;
      if constexpr (kWide) {
        // 16 columns at a time: sum the kMain + 1 partial accumulators, bias / row scale / ReLU, 64-byte store
        const float rscale = (p.row_scale && r < p.M) ? __ldg(p.row_scale + r) : 1.0f;
        float* yrow = p.Y + r * p.N;
        const bool y16 = (p.N & 3) == 0 && (reinterpret_cast<uintptr_t>(p.Y) & 15) == 0;
#pragma unroll 1
        for (int c0 = 0; c0 < NPAD; c0 += 16) {
            if (c0 >= p.N) break;               // warp-uniform: the tcgen05.ld below is warp-collective
            float o[16];
#pragma unroll
            for (int j = 0; j < 16; ++j) o[j] = 0.0f;
#pragma unroll
            for (int a = 0; a <= kMain; ++a) {
                if (a < kMain && a >= nchunks) continue;    // main accumulator a is written only if there are more than a chunks
                uint32_t u[16];
                const uint32_t taddr = tmem_base + ((uint32_t)(warp * 32) << 16) + (uint32_t)(a * NPAD + c0);
                asm volatile(
                    "tcgen05.ld.sync.aligned.32x32b.x16.b32 {%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15}, [%16];"
                    : "=r"(u[0]), "=r"(u[1]), "=r"(u[2]), "=r"(u[3]), "=r"(u[4]), "=r"(u[5]), "=r"(u[6]), "=r"(u[7]),
                      "=r"(u[8]), "=r"(u[9]), "=r"(u[10]), "=r"(u[11]), "=r"(u[12]), "=r"(u[13]), "=r"(u[14]), "=r"(u[15])
                    : "r"(taddr));
                asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
#pragma unroll
                for (int j = 0; j < 16; ++j) o[j] += __uint_as_float(u[j]);
            }
            if (r < p.M) {
#pragma unroll
                for (int j = 0; j < 16; ++j) {
                    float y = o[j] + ((p.bias && c0 + j < p.N) ? __ldg(p.bias + c0 + j) : 0.0f);
                    y *= rscale;
                    if (p.relu) y = fmaxf(y, 0.0f);
                    o[j] = y;
                }
                if (y16) {
#pragma unroll
                    for (int j = 0; j < 16; j += 4)
                        if (c0 + j < p.N) __stcs(reinterpret_cast<float4*>(yrow + c0 + j), make_float4(o[j], o[j + 1], o[j + 2], o[j + 3]));
                } else {
#pragma unroll
                    for (int j = 0; j < 16; ++j)
                        if (c0 + j < p.N) yrow[c0 + j] = o[j];
                }
            }
        }
      } else {
        float acc[NPAD];
#pragma unroll
        for (int n = 0; n < NPAD; ++n) acc[n] = 0.0f;
#pragma unroll
        for (int c0 = 0; c0 < (int)kAccCols; c0 += 16) {
            // main accumulator a is written only if the tile has more than a K chunks
            if (c0 / NPAD < kMain && c0 / NPAD >= nchunks) continue;
            uint32_t u[16];
            const uint32_t taddr = tmem_base + ((uint32_t)(warp * 32) << 16) + (uint32_t)c0;
            asm volatile(
                "tcgen05.ld.sync.aligned.32x32b.x16.b32 {%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15}, [%16];"
                : "=r"(u[0]), "=r"(u[1]), "=r"(u[2]), "=r"(u[3]), "=r"(u[4]), "=r"(u[5]), "=r"(u[6]), "=r"(u[7]),
                  "=r"(u[8]), "=r"(u[9]), "=r"(u[10]), "=r"(u[11]), "=r"(u[12]), "=r"(u[13]), "=r"(u[14]), "=r"(u[15])
                : "r"(taddr));
            asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
#pragma unroll
            for (int j = 0; j < 16; ++j) acc[(c0 + j) % NPAD] += __uint_as_float(u[j]);
        }
        if (r < p.M) {
            float a0 = p.att_b_dev ? __ldg(p.att_b_dev) : p.att_b0, a1 = p.att_b_dev ? __ldg(p.att_b_dev + 1) : p.att_b1;
            const float rscale = p.row_scale ? __ldg(p.row_scale + r) : 1.0f;
#pragma unroll
            for (int n = 0; n < NPAD; ++n) {
                if (n < p.N) {
                    float y = acc[n] + (p.bias ? __ldg(p.bias + n) : 0.0f);
                    if (p.att_w) {   // projections of the (pre-activation) output row: aL = y.wl + bl, aR = y.wr + br
                        a0 = fmaf(y, __ldg(p.att_w + n), a0);
                        a1 = fmaf(y, __ldg(p.att_w + p.N + n), a1);
                    }
                    y *= rscale;
                    if (p.relu) y = fmaxf(y, 0.0f);
                    acc[n] = y;
                }
            }
            float* yrow = p.Y + r * p.N;
            if (p.mo.count > 0) {
#pragma unroll
                for (int n = 0; n < NPAD; n += 4)
                    if (n < p.N) {
                        Vec<4> o;
                        o.v[0] = acc[n]; o.v[1] = acc[n + 1]; o.v[2] = acc[n + 2]; o.v[3] = acc[n + 3];
                        multi_store<4>(p.mo, r * p.N + n, o, r);
                    }
            } else if ((p.N & 3) == 0 && (reinterpret_cast<uintptr_t>(yrow) & 15) == 0) {
#pragma unroll
                for (int n = 0; n < NPAD; n += 4)
                    if (n < p.N) *reinterpret_cast<float4*>(yrow + n) = make_float4(acc[n], acc[n + 1], acc[n + 2], acc[n + 3]);
            } else {
#pragma unroll
                for (int n = 0; n < NPAD; ++n)
                    if (n < p.N) yrow[n] = acc[n];
            }
            if (p.att_w) {
                p.att_out[r] = a0;
                p.att_out[p.M + r] = a1;
                if (p.att_mo.count > 0) {
                    Vec<1> o1;
                    o1.v[0] = a1;
                    multi_store<1>(p.att_mo, r, o1, r);
                }
            }
        }
      }
        asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    } else if (lane == 0) {
        // ---------------- MMA issuer (one thread) ---------------------------------------------------
        constexpr uint32_t idesc = make_idesc(NPAD);
        for (int ch = 0; ch < nchunks; ++ch) {
            const int s = ch % kStages;
            const uint32_t ph = (uint32_t)((ch / kStages) & 1);
            mbar_wait(&full_bar[s], ph);
            asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
            const uint32_t a_hi = smem_u32(smem + (size_t)s * kStageBytes);
            const uint32_t a_lo = a_hi + kABytes;
            const uint32_t b_hi = a_lo + kABytes;
            const uint32_t b_lo = b_hi + kBBytes;
            const uint32_t d_main = tmem_base + (uint32_t)((ch % kMain) * NPAD);
            const uint32_t d_corr = tmem_base + (uint32_t)(kMain * NPAD);
#pragma unroll
            for (int k8 = 0; k8 < kBK / 8; ++k8) {   // UMMA_K = 8 tf32 = 32 bytes along the swizzled row
                const uint32_t ko = (uint32_t)k8 * 32;
                umma_tf32(d_main, make_desc(a_hi + ko), make_desc(b_hi + ko), idesc, (ch >= kMain) | (k8 != 0));
                umma_tf32(d_corr, make_desc(a_hi + ko), make_desc(b_lo + ko), idesc, (ch | k8) != 0);
                umma_tf32(d_corr, make_desc(a_lo + ko), make_desc(b_hi + ko), idesc, 1);
            }
            umma_commit(&empty_bar[s]);            // frees the stage when these MMAs have read it
        }
        umma_commit(&acc_bar);                     // accumulator complete
    }
    __syncthreads();
    if (warp == 4) {
        asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
        asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(kTmemCols) : "memory");
    }
}

// ---------------------------------------------------------------------------------------------
// v2: persistent, warp-specialised (one CTA per SM).
//   warps 0-7   producers: 16 rows each, one 8-byte load per lane per row = 256 contiguous bytes of a
//               row per instruction (two 32-float swizzle atoms), next super-stage prefetched in
//               registers while the current one is split and stored;
//   warp  8     MMA issuer (one thread), 2 shared-memory super-stages, 2 TMEM accumulator sets;
//   warps 9-12  epilogue of tile i while the producers / tensor core already work on tile i+1.
// Requires even K and 8-byte aligned X / W (the 4-byte kernel above covers the rest).
constexpr int kV2Producers = 8;
constexpr int kV2Threads = (kV2Producers + 1 + 4) * 32;   // 416

template <int NPAD>
__global__ void __launch_bounds__(kV2Threads, 1) linear_tf32x3_v2_kernel(const __grid_constant__ LinearParams p) {
    constexpr int kSK = 64;                                  // floats of K per super-stage (2 atoms)
    constexpr uint32_t kAtomA = kBM * 128;                   // 16 KB: [128 rows x 32 fp32]
    constexpr uint32_t kAtomB = NPAD * 128;
    constexpr uint32_t kStageBytes = 4 * kAtomA + 4 * kAtomB;   // A hi[2] lo[2], B hi[2] lo[2]
    constexpr int kMain = NPAD == 64 ? 3 : 4;
    constexpr uint32_t kAccCols = (kMain + 1) * NPAD;
    constexpr uint32_t kTmemCols = 2 * kAccCols <= 256 ? 256 : 512;
    constexpr int kWRows = NPAD / kV2Producers;
    extern __shared__ __align__(1024) uint8_t smem_raw[];
    uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
    __shared__ __align__(8) uint64_t full_bar[2], empty_bar[2], accf_bar[2], acce_bar[2];
    __shared__ uint32_t tmem_base_smem;
    // epilogue staging: each epilogue warp transposes its 32 rows through shared memory so that a
    // row leaves as one contiguous store (full 128-byte NVLink / HBM writes instead of 16-byte pieces)
    __shared__ __align__(16) float stage_out[4][32][NPAD + 4];

    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    const int nss = (p.K + kSK - 1) / kSK;
    // Every CTA owns one contiguous, equally long range of rows and walks it in 128-row tiles (the last one partial):
    // the bytes per CTA are equal whatever M is.  (Round-robin over fixed 128-row tiles left the last wave mostly empty
    // when a rank holds only ~1.5 tiles per SM: 228 tiles on 148 CTAs at 8 GPUs.)
    const int64_t rows_per_cta = ((p.M + gridDim.x - 1) / gridDim.x + 7) & ~int64_t(7);
    const int64_t row_lo = min(p.M, (int64_t)blockIdx.x * rows_per_cta);
    // Measured on B200 (profiles/r02_linear_tile_assignment.txt): the contiguous ranges win when a CTA holds few tiles and at
    // small K (Products K = 100: 0.695 -> 0.654 ms), round-robin tiles win at large K with many tiles per CTA (Reddit
    // K = 602: 0.239 vs 0.220 ms -- every CTA's last, partial tile pays a full tile's MMA chain and epilogue).
    const int64_t ntiles_all = (p.M + kBM - 1) / kBM;
    const int64_t row_end = p.round_robin ? p.M : min(p.M, row_lo + rows_per_cta);
    const uint32_t my_tiles = p.round_robin
        ? ((int64_t)blockIdx.x < ntiles_all ? (uint32_t)((ntiles_all - blockIdx.x + gridDim.x - 1) / gridDim.x) : 0u)
        : (uint32_t)((row_end - row_lo + kBM - 1) / kBM);
    auto tile_row = [&](uint32_t it) -> int64_t {
        return p.round_robin ? ((int64_t)blockIdx.x + (int64_t)it * gridDim.x) * kBM : row_lo + (int64_t)it * kBM;
    };

    if (tid == 0) {
        for (int s = 0; s < 2; ++s) {
            mbar_init(&full_bar[s], kV2Producers * 32);
            mbar_init(&empty_bar[s], 1);
            mbar_init(&accf_bar[s], 1);
            mbar_init(&acce_bar[s], 128);
        }
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    if (warp == kV2Producers) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(&tmem_base_smem)),
                     "r"(kTmemCols)
                     : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
    }
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    const uint32_t tmem_base = tmem_base_smem;

    if (warp < kV2Producers) {
        // ------------------------------- producers ------------------------------------------------
        const uint32_t pitch = (uint32_t)p.K * 4u;
        const int atom = lane >> 4, col = (2 * lane) & 31;
        uint32_t swa[8], swb[kWRows];
#pragma unroll
        for (int j = 0; j < 8; ++j)
            swa[j] = (uint32_t)atom * kAtomA + (uint32_t)(((((col >> 2) ^ j)) << 4) | ((col & 3) << 2));
#pragma unroll
        for (int i = 0; i < kWRows; ++i) {
            const int n = warp * kWRows + i;
            swb[i] = (uint32_t)atom * kAtomB + (uint32_t)(n * 128) + (uint32_t)(((((col >> 2) ^ (n & 7))) << 4) | ((col & 3) << 2));
        }
        const char* wlane = reinterpret_cast<const char*>(p.W + (int64_t)(warp * kWRows) * p.K + 2 * lane);
        const bool wrows_full = warp * kWRows + kWRows <= p.N;
        float2 va[16], vb[16];
        const float2 zero2 = make_float2(0.0f, 0.0f);

        auto load_ss = [&](const char* xlane, bool rows_full, int64_t wrow0, int ss, float2 (&v)[16]) {
            const bool kok = ss * kSK + 2 * lane < p.K;   // K even: the pair is in or out together
            const char* xc = xlane + ss * (kSK * 4);
            if (rows_full && kok) {
#pragma unroll
                for (int i = 0; i < 16; ++i) v[i] = __ldg(reinterpret_cast<const float2*>(xc + (uint64_t)i * pitch));
            } else {
#pragma unroll
                for (int i = 0; i < 16; ++i)
                    v[i] = (kok && wrow0 + i < row_end) ? __ldg(reinterpret_cast<const float2*>(xc + (uint64_t)i * pitch)) : zero2;
            }
        };
        auto store_ss = [&](uint32_t g, const float2 (&v)[16]) {
            const uint32_t s = g & 1;
            // this warp's rows of W for the stage (L2-resident, a few hundred bytes): requested here, consumed after the
            // A rows have been split, so their latency hides behind the wait for the stage and the A stores
            const int ss = (int)(g % (uint32_t)nss);
            const bool kok = ss * kSK + 2 * lane < p.K;
            const char* wc = wlane + ss * (kSK * 4);
            float2 w[kWRows];
#pragma unroll
            for (int i = 0; i < kWRows; ++i)
                w[i] = (kok && (wrows_full || warp * kWRows + i < p.N))
                           ? __ldg(reinterpret_cast<const float2*>(wc + (uint64_t)i * pitch)) : zero2;
            if (g >= 2) mbar_wait(&empty_bar[s], ((g >> 1) & 1) ^ 1);
            uint8_t* a_hi = smem + (size_t)s * kStageBytes + warp * (16 * 128);
            uint8_t* a_lo = a_hi + 2 * kAtomA;
            uint8_t* b_hi = smem + (size_t)s * kStageBytes + 4 * kAtomA;
            uint8_t* b_lo = b_hi + 2 * kAtomB;
#pragma unroll
            for (int i = 0; i < 16; ++i) {
                float2 hi, lo;
                split_tf32(v[i].x, hi.x, lo.x);
                split_tf32(v[i].y, hi.y, lo.y);
                *reinterpret_cast<float2*>(a_hi + i * 128 + swa[i & 7]) = hi;
                *reinterpret_cast<float2*>(a_lo + i * 128 + swa[i & 7]) = lo;
            }
#pragma unroll
            for (int i = 0; i < kWRows; ++i) {
                float2 hi, lo;
                split_tf32(w[i].x, hi.x, lo.x);
                split_tf32(w[i].y, hi.y, lo.y);
                *reinterpret_cast<float2*>(b_hi + swb[i]) = hi;
                *reinterpret_cast<float2*>(b_lo + swb[i]) = lo;
            }
            asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
            mbar_arrive(&full_bar[s]);
        };

        // The (tile, super-stage) pairs this CTA processes form one flat sequence g = 0, 1, ...: the register pipeline
        // below never drains at a tile boundary.  GALA_LINEAR_REGBUF super-stages are in flight per lane (x 16 rows x
        // 8 bytes = 64 KB per SM at 2).  A third buffer (96 KB, above the ~75 KB the HBM latency-bandwidth product
        // asks of one SM) does not fit: 13 warps put 4 on one SM sub-partition, whose 16 K registers cap a thread at
        // 128, and the third buffer spills (156 bytes at N = 32).
#ifndef GALA_LINEAR_REGBUF
#define GALA_LINEAR_REGBUF 2
#endif
        const uint32_t total = my_tiles * (uint32_t)nss;
        auto load_g = [&](uint32_t g, float2 (&v)[16]) {
            if (g >= total) return;
            const uint32_t it = g / (uint32_t)nss;
            const int ss = (int)(g - it * (uint32_t)nss);
            const int64_t wrow0 = tile_row(it) + warp * 16;
            const char* xlane = reinterpret_cast<const char*>(p.X + wrow0 * p.K + 2 * lane);
            load_ss(xlane, wrow0 + 16 <= row_end, wrow0, ss, v);
        };
#if GALA_LINEAR_REGBUF == 3
        float2 vc[16];
        load_g(0, va);
        load_g(1, vb);
        for (uint32_t g = 0; g < total; g += 3) {
            load_g(g + 2, vc);
            store_ss(g, va);
            if (g + 1 >= total) break;
            load_g(g + 3, va);
            store_ss(g + 1, vb);
            if (g + 2 >= total) break;
            load_g(g + 4, vb);
            store_ss(g + 2, vc);
        }
#else
        load_g(0, va);
        for (uint32_t g = 0; g < total; g += 2) {
            load_g(g + 1, vb);
            store_ss(g, va);
            if (g + 1 >= total) break;
            load_g(g + 2, va);
            store_ss(g + 1, vb);
        }
#endif
    } else if (warp == kV2Producers) {
        // ------------------------------- MMA issuer -----------------------------------------------
        if (lane == 0) {
            constexpr uint32_t idesc = make_idesc(NPAD);
            uint32_t g = 0, it = 0;
            for (; it < my_tiles; ++it) {
                const uint32_t aset = it & 1;
                if (it >= 2) mbar_wait(&acce_bar[aset], ((it >> 1) & 1) ^ 1);
                asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
                const uint32_t acc0 = tmem_base + aset * kAccCols;
                for (int ss = 0; ss < nss; ++ss, ++g) {
                    const uint32_t s = g & 1;
                    mbar_wait(&full_bar[s], (g >> 1) & 1);
                    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
                    const uint32_t a_hi = smem_u32(smem + (size_t)s * kStageBytes);
                    const uint32_t a_lo = a_hi + 2 * kAtomA;
                    const uint32_t b_hi = a_hi + 4 * kAtomA;
                    const uint32_t b_lo = b_hi + 2 * kAtomB;
                    const uint32_t d_main = acc0 + (uint32_t)((ss % kMain) * NPAD);
                    const uint32_t d_corr = acc0 + (uint32_t)(kMain * NPAD);
                    const int natoms = (p.K - ss * kSK) > 32 ? 2 : 1;
                    for (int at = 0; at < natoms; ++at) {
#pragma unroll
                        for (int k8 = 0; k8 < 4; ++k8) {
                            const uint32_t ao = (uint32_t)at * kAtomA + (uint32_t)k8 * 32;
                            const uint32_t bo = (uint32_t)at * kAtomB + (uint32_t)k8 * 32;
                            const uint32_t first = (uint32_t)((at | k8) == 0);
                            umma_tf32(d_main, make_desc(a_hi + ao), make_desc(b_hi + bo), idesc, !(first && ss < kMain));
                            umma_tf32(d_corr, make_desc(a_hi + ao), make_desc(b_lo + bo), idesc, !(first && ss == 0));
                            umma_tf32(d_corr, make_desc(a_lo + ao), make_desc(b_hi + bo), idesc, 1);
                        }
                    }
                    umma_commit(&empty_bar[s]);
                }
                umma_commit(&accf_bar[aset]);
            }
        }
    } else {
        // ------------------------------- epilogue warps -------------------------------------------
        const int q = warp & 3;                  // TMEM lane quarter this warp may read
        uint32_t it = 0;
        for (; it < my_tiles; ++it) {
            const uint32_t aset = it & 1;
            mbar_wait(&accf_bar[aset], (it >> 1) & 1);
            asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
            const int64_t r = tile_row(it) + q * 32 + lane;
            float acc[NPAD];
#pragma unroll
            for (int n = 0; n < NPAD; ++n) acc[n] = 0.0f;
#pragma unroll
            for (int c0 = 0; c0 < (int)kAccCols; c0 += 16) {
                if (c0 / NPAD < kMain && c0 / NPAD >= nss) continue;   // accumulator never written
                uint32_t u[16];
                const uint32_t taddr = tmem_base + ((uint32_t)(q * 32) << 16) + aset * kAccCols + (uint32_t)c0;
                asm volatile(
                    "tcgen05.ld.sync.aligned.32x32b.x16.b32 {%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15}, [%16];"
                    : "=r"(u[0]), "=r"(u[1]), "=r"(u[2]), "=r"(u[3]), "=r"(u[4]), "=r"(u[5]), "=r"(u[6]), "=r"(u[7]),
                      "=r"(u[8]), "=r"(u[9]), "=r"(u[10]), "=r"(u[11]), "=r"(u[12]), "=r"(u[13]), "=r"(u[14]), "=r"(u[15])
                    : "r"(taddr));
                asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
#pragma unroll
                for (int j = 0; j < 16; ++j) acc[(c0 + j) % NPAD] += __uint_as_float(u[j]);
            }
            asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
            mbar_arrive(&acce_bar[aset]);        // the accumulator set may be overwritten
            float a0 = p.att_b_dev ? __ldg(p.att_b_dev) : p.att_b0, a1 = p.att_b_dev ? __ldg(p.att_b_dev + 1) : p.att_b1;
            if (r < row_end) {
                const float rscale = p.row_scale ? __ldg(p.row_scale + r) : 1.0f;
#pragma unroll
                for (int n = 0; n < NPAD; ++n) {
                    if (n < p.N) {
                        float y = acc[n] + (p.bias ? __ldg(p.bias + n) : 0.0f);
                        if (p.att_w) {
                            a0 = fmaf(y, __ldg(p.att_w + n), a0);
                            a1 = fmaf(y, __ldg(p.att_w + p.N + n), a1);
                        }
                        y *= rscale;
                        if (p.relu) y = fmaxf(y, 0.0f);
                        acc[n] = y;
                    }
                }
                if (p.att_w) {
                    p.att_out[r] = a0;
                    p.att_out[p.M + r] = a1;
                    if (p.att_mo.count > 0) {
                        Vec<1> o1;
                        o1.v[0] = a1;
                        multi_store<1>(p.att_mo, r, o1, r);
                    }
                }
            }
            if ((p.N & 3) == 0 && (p.mo.count > 0 || (reinterpret_cast<uintptr_t>(p.Y) & 15) == 0)) {
                // transpose through shared memory: lane l holds row l -> a group of N/4 lanes stores one row
                float (*st)[NPAD + 4] = stage_out[q];
#pragma unroll
                for (int n = 0; n < NPAD; n += 4)
                    if (n < p.N) *reinterpret_cast<float4*>(&st[lane][n]) = make_float4(acc[n], acc[n + 1], acc[n + 2], acc[n + 3]);
                __syncwarp();
                const int vec_per_row = p.N >> 2;
                const int64_t tile_row0 = tile_row(it) + q * 32;
                for (int idx = lane; idx < 32 * vec_per_row; idx += 32) {
                    const int rr = idx / vec_per_row, cv = idx - rr * vec_per_row;
                    const int64_t row = tile_row0 + rr;
                    if (row < row_end) {
                        const float4 v = *reinterpret_cast<const float4*>(&st[rr][cv * 4]);
                        Vec<4> o;
                        o.v[0] = v.x; o.v[1] = v.y; o.v[2] = v.z; o.v[3] = v.w;
                        if (p.mo.count > 0) multi_store<4>(p.mo, row * p.N + cv * 4, o, row);
                        else o.store(p.Y + row * p.N + cv * 4);
                    }
                }
                __syncwarp();
            } else if (r < row_end) {
                float* yrow = p.Y + r * p.N;
#pragma unroll
                for (int n = 0; n < NPAD; ++n)
                    if (n < p.N) yrow[n] = acc[n];
            }
        }
    }
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    if (warp == kV2Producers) {
        asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
        asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(kTmemCols) : "memory");
    }
}

// SM count of the CURRENT device (cached per device ordinal; benign if two threads race to fill a slot)
inline int device_sm_count() {
    static int cache[64] = {0};
    int dev = 0;
    if (cudaGetDevice(&dev) != cudaSuccess || dev < 0 || dev >= 64) return 148;
    if (cache[dev] == 0) {
        int sms = 0;
        if (cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev) != cudaSuccess || sms <= 0) sms = 148;
        cache[dev] = sms;
    }
    return cache[dev];
}

// ---------------------------------------------------------------------------------------------
// Wide output, small K (the classifier Linear(hidden, classes) with 64 < classes <= 256 and hidden <= 64: 172 classes
// on the Papers shape).  The output is 5x the input here, so the kernel is a store stream: persistent, one CTA per SM,
//   * W (N x K) is split into hi / lo ONCE per CTA and stays in shared memory in UMMA's swizzled layout;
//   * warps 0-3 load / split the 128 x K tile of X (two register buffers over the flat (tile, atom) sequence);
//   * warp 4 issues the MMAs into ONE accumulator of NPAD columns (at most 24 accumulations: no drift to spread),
//     two accumulator sets in TMEM (2 x NPAD <= 512 columns);
//   * warps 5-12 drain set i while the tensor core fills set i+1: two warps per TMEM lane quarter, each taking every
//     other 32-column group (bias / row scale / ReLU) -- the epilogue is the long stage of this kernel (5x more bytes
//     leave than enter), so it gets the warps.
constexpr int kWsThreads = 13 * 32;

//   * STAGED (whenever the shared memory is there: K <= 32): an epilogue warp transposes its 32 x N block through shared
//     memory and writes it as ONE contiguous run (32 rows of a packed [M, N] matrix are 32*N*4 consecutive bytes):
//     full 128-byte lines per store instruction instead of 32 scattered 16-byte pieces (688 vs 5504 L2 requests per tile).
template <int NPAD, int NATOM, bool STAGED>
__global__ void __launch_bounds__(kWsThreads, 1) linear_tf32x3_wide_smallk_kernel(const __grid_constant__ LinearParams p) {
    constexpr uint32_t kAtomA = kBM * 128;                 // [128 rows x 32 fp32]
    constexpr uint32_t kAtomB = NPAD * 128;
    constexpr uint32_t kWBytes = 2 * NATOM * kAtomB;       // hi atoms, then lo atoms
    constexpr uint32_t kStageBytes = 2 * kAtomA;           // one atom of A: hi, lo
    constexpr uint32_t kTmemCols = 2 * NPAD <= 256 ? 256 : 512;
    static_assert(2 * NPAD <= 512 && NPAD % 16 == 0, "two accumulator sets in TMEM");
    extern __shared__ __align__(1024) uint8_t smem_raw[];
    uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
    uint8_t* w_hi = smem;
    uint8_t* w_lo = smem + NATOM * kAtomB;
    uint8_t* a_stage = smem + kWBytes;
    constexpr int kOutPitch = NPAD + 4;                    // floats; 16-byte aligned rows, conflict-free 128-bit row writes
    float* out_stage = reinterpret_cast<float*>(a_stage + 2 * kStageBytes);   // STAGED: [4 warps][32 rows][kOutPitch]
    __shared__ __align__(8) uint64_t full_bar[2], empty_bar[2], accf_bar[2], acce_bar[2];
    __shared__ uint32_t tmem_base_smem;
    __shared__ __align__(16) float bias_s[NPAD];

    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    const int64_t ntiles = (p.M + kBM - 1) / kBM;
    const uint32_t my_tiles = blockIdx.x < ntiles ? (uint32_t)((ntiles - blockIdx.x + gridDim.x - 1) / gridDim.x) : 0u;
    for (int n = tid; n < NPAD; n += kWsThreads) bias_s[n] = (p.bias && n < p.N) ? __ldg(p.bias + n) : 0.0f;

    if (tid == 0) {
        for (int s = 0; s < 2; ++s) {
            mbar_init(&full_bar[s], 128);
            mbar_init(&empty_bar[s], 1);
            mbar_init(&accf_bar[s], 1);
            mbar_init(&acce_bar[s], 256);
        }
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    if (warp == 4) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(&tmem_base_smem)),
                     "r"(kTmemCols)
                     : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
    }
    // W -> hi / lo, swizzled K-major, resident for the whole kernel (rows >= N and columns >= K are zero)
    for (int idx = tid; idx < NPAD * NATOM * 32; idx += kWsThreads) {
        const int n = idx / (NATOM * 32), k = idx - n * (NATOM * 32);
        const float w = (n < p.N && k < p.K) ? __ldg(p.W + (int64_t)n * p.K + k) : 0.0f;
        float hi, lo;
        split_tf32(w, hi, lo);
        const int atom = k >> 5, kk = k & 31;
        const uint32_t o = (uint32_t)atom * kAtomB + (uint32_t)(n * 128) + (uint32_t)((((kk >> 2) ^ (n & 7)) << 4) | ((kk & 3) << 2));
        *reinterpret_cast<float*>(w_hi + o) = hi;
        *reinterpret_cast<float*>(w_lo + o) = lo;
    }
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    const uint32_t tmem_base = tmem_base_smem;

    if (warp < 4) {
        // ------------------------------- producers: 32 rows each, one float per lane per row and atom ------------
        const uint32_t pitch = (uint32_t)p.K * 4u;
        const uint32_t total = my_tiles * NATOM;
        uint32_t sw[8];
#pragma unroll
        for (int j = 0; j < 8; ++j) sw[j] = (uint32_t)((((lane >> 2) ^ j) << 4) | ((lane & 3) << 2));
        // kBuf tile-atoms are in flight per lane (kBuf x 16 KB per SM): a tile is only 128 x K x 4 bytes here, so the
        // register pipeline has to be deep for the loads to cover the HBM latency
        constexpr int kBuf = 2;     // (4 measured no faster: the epilogue, not the loads, is the long stage)
        float v[kBuf][32];
        auto load_g = [&](uint32_t g, float (&vv)[32]) {
            if (g >= total) return;
            const uint32_t it = g / NATOM;
            const int atom = (int)(g - it * NATOM);
            const int64_t wrow0 = ((int64_t)blockIdx.x + (int64_t)it * gridDim.x) * kBM + warp * 32;
            const bool kok = atom * 32 + lane < p.K;
            const char* xc = reinterpret_cast<const char*>(p.X + wrow0 * p.K + atom * 32 + lane);
            if (kok && wrow0 + 32 <= p.M) {
#pragma unroll
                for (int i = 0; i < 32; ++i) vv[i] = __ldcs(reinterpret_cast<const float*>(xc + (uint64_t)i * pitch));
            } else {
#pragma unroll
                for (int i = 0; i < 32; ++i)
                    vv[i] = (kok && wrow0 + i < p.M) ? __ldcs(reinterpret_cast<const float*>(xc + (uint64_t)i * pitch)) : 0.0f;
            }
        };
        auto store_g = [&](uint32_t g, const float (&vv)[32]) {
            const uint32_t s = g & 1;
            if (g >= 2) mbar_wait(&empty_bar[s], ((g >> 1) & 1) ^ 1);
            uint8_t* a_hi = a_stage + (size_t)s * kStageBytes + warp * 4096;
            uint8_t* a_lo = a_hi + kAtomA;
#pragma unroll
            for (int i = 0; i < 32; ++i) {
                float hi, lo;
                split_tf32(vv[i], hi, lo);
                *reinterpret_cast<float*>(a_hi + i * 128 + sw[i & 7]) = hi;
                *reinterpret_cast<float*>(a_lo + i * 128 + sw[i & 7]) = lo;
            }
            asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
            mbar_arrive(&full_bar[s]);
        };
#pragma unroll
        for (int b = 0; b < kBuf - 1; ++b) load_g((uint32_t)b, v[b]);
        for (uint32_t g0 = 0; g0 < total; g0 += kBuf) {
#pragma unroll
            for (int b = 0; b < kBuf; ++b) {
                const uint32_t g = g0 + b;
                if (g < total) {
                    load_g(g + kBuf - 1, v[(b + kBuf - 1) % kBuf]);
                    store_g(g, v[b]);
                }
            }
        }
    } else if (warp == 4) {
        // ------------------------------- MMA issuer ---------------------------------------------------------------
        if (lane == 0) {
            constexpr uint32_t idesc = make_idesc(NPAD);
            uint32_t g = 0;
            for (uint32_t it = 0; it < my_tiles; ++it) {
                const uint32_t aset = it & 1;
                if (it >= 2) mbar_wait(&acce_bar[aset], ((it >> 1) & 1) ^ 1);
                asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
                const uint32_t acc = tmem_base + aset * NPAD;
                for (int atom = 0; atom < NATOM; ++atom, ++g) {
                    const uint32_t s = g & 1;
                    mbar_wait(&full_bar[s], (g >> 1) & 1);
                    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
                    const uint32_t a_hi = smem_u32(a_stage + (size_t)s * kStageBytes), a_lo = a_hi + kAtomA;
                    const uint32_t b_hi = smem_u32(w_hi) + (uint32_t)atom * kAtomB, b_lo = smem_u32(w_lo) + (uint32_t)atom * kAtomB;
                    if (atom * 32 < p.K) {
#pragma unroll
                        for (int k8 = 0; k8 < 4; ++k8) {
                            const uint32_t ko = (uint32_t)k8 * 32;
                            umma_tf32(acc, make_desc(a_hi + ko), make_desc(b_hi + ko), idesc, (atom | k8) != 0);
                            umma_tf32(acc, make_desc(a_hi + ko), make_desc(b_lo + ko), idesc, 1);
                            umma_tf32(acc, make_desc(a_lo + ko), make_desc(b_hi + ko), idesc, 1);
                        }
                    }
                    umma_commit(&empty_bar[s]);
                }
                umma_commit(&accf_bar[aset]);
            }
        }
    } else {
        // ------------------------------- epilogue warps ------------------------------------------------------------
        const int q = warp & 3;                  // TMEM lane quarter this warp may read (warp id mod 4)
        const int half = (warp - 5) >> 2;        // the two warps of a quarter take alternate 32-column groups
        const bool y16 = (p.N & 3) == 0 && (reinterpret_cast<uintptr_t>(p.Y) & 15) == 0;
        auto pair_sync = [&]() { asm volatile("bar.sync %0, 64;" ::"r"(1 + q) : "memory"); };   // the 2 warps of this quarter
        for (uint32_t it = 0; it < my_tiles; ++it) {
            const uint32_t aset = it & 1;
            mbar_wait(&accf_bar[aset], (it >> 1) & 1);
            asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
            const int64_t r = ((int64_t)blockIdx.x + (int64_t)it * gridDim.x) * kBM + q * 32 + lane;
            const float rscale = (p.row_scale && r < p.M) ? __ldg(p.row_scale + r) : 1.0f;
            float* yrow = p.Y + r * p.N;
#pragma unroll 1
            for (int c0 = half * 32; c0 < NPAD; c0 += 64) {
                if (c0 >= p.N) break;            // warp-uniform
                uint32_t u[32];
                const uint32_t taddr = tmem_base + ((uint32_t)(q * 32) << 16) + aset * NPAD + (uint32_t)c0;
                asm volatile(
                    "tcgen05.ld.sync.aligned.32x32b.x16.b32 {%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15}, [%16];"
                    : "=r"(u[0]), "=r"(u[1]), "=r"(u[2]), "=r"(u[3]), "=r"(u[4]), "=r"(u[5]), "=r"(u[6]), "=r"(u[7]),
                      "=r"(u[8]), "=r"(u[9]), "=r"(u[10]), "=r"(u[11]), "=r"(u[12]), "=r"(u[13]), "=r"(u[14]), "=r"(u[15])
                    : "r"(taddr));
                if (c0 + 16 < NPAD)
                    asm volatile(
                        "tcgen05.ld.sync.aligned.32x32b.x16.b32 {%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15}, [%16];"
                        : "=r"(u[16]), "=r"(u[17]), "=r"(u[18]), "=r"(u[19]), "=r"(u[20]), "=r"(u[21]), "=r"(u[22]), "=r"(u[23]),
                          "=r"(u[24]), "=r"(u[25]), "=r"(u[26]), "=r"(u[27]), "=r"(u[28]), "=r"(u[29]), "=r"(u[30]), "=r"(u[31])
                        : "r"(taddr + 16));
                asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
                float* srow = out_stage + ((size_t)q * 32 + lane) * kOutPitch + c0;
#pragma unroll
                for (int j = 0; j < 32; j += 4) {
                    if (c0 + j < NPAD && c0 + j < p.N) {
                        const float4 b4 = *reinterpret_cast<const float4*>(&bias_s[c0 + j]);     // broadcast read
                        float4 o;
                        o.x = (__uint_as_float(u[j]) + b4.x) * rscale;
                        o.y = (__uint_as_float(u[j + 1]) + b4.y) * rscale;
                        o.z = (__uint_as_float(u[j + 2]) + b4.z) * rscale;
                        o.w = (__uint_as_float(u[j + 3]) + b4.w) * rscale;
                        if (p.relu) {
                            o.x = fmaxf(o.x, 0.0f); o.y = fmaxf(o.y, 0.0f); o.z = fmaxf(o.z, 0.0f); o.w = fmaxf(o.w, 0.0f);
                        }
                        if constexpr (STAGED) {
                            *reinterpret_cast<float4*>(srow + j) = o;
                        } else if (r < p.M) {
                            if (y16) {
                                __stcs(reinterpret_cast<float4*>(yrow + c0 + j), o);
                            } else {
                                const float ov[4] = {o.x, o.y, o.z, o.w};
#pragma unroll
                                for (int v = 0; v < 4; ++v)
                                    if (c0 + j + v < p.N) yrow[c0 + j + v] = ov[v];
                            }
                        }
                    }
                }
            }
            // the accumulator set is in registers / shared memory now: hand it back before the stores go out
            asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
            mbar_arrive(&acce_bar[aset]);
            if constexpr (STAGED) {
                pair_sync();                     // both column halves of the quarter's 32 rows are staged
                // 32 packed rows of Y are one contiguous run; each of the two warps writes 16 of them
                const int64_t row0 = ((int64_t)blockIdx.x + (int64_t)it * gridDim.x) * kBM + q * 32 + half * 16;
                const int rows_here = (int)max((int64_t)0, min((int64_t)16, p.M - row0));
                const float* sbase = out_stage + ((size_t)q * 32 + half * 16) * kOutPitch;
                if (y16) {
                    const int vpr = p.N >> 2;                                 // float4 per row
                    float4* ybase = reinterpret_cast<float4*>(p.Y + row0 * p.N);
                    int rr = 0, cv = lane;
                    while (cv >= vpr) { cv -= vpr; ++rr; }
                    for (int idx = lane; idx < rows_here * vpr; idx += 32) {
                        __stcs(ybase + idx, *reinterpret_cast<const float4*>(sbase + rr * kOutPitch + cv * 4));
                        cv += 32;
                        while (cv >= vpr) { cv -= vpr; ++rr; }
                    }
                } else {
                    float* ybase = p.Y + row0 * p.N;
                    int rr = 0, c = lane;
                    while (c >= p.N) { c -= p.N; ++rr; }
                    for (int idx = lane; idx < rows_here * p.N; idx += 32) {
                        ybase[idx] = sbase[rr * kOutPitch + c];
                        c += 32;
                        while (c >= p.N) { c -= p.N; ++rr; }
                    }
                }
                pair_sync();                     // the staging rows may be overwritten by the next tile
            }
        }
    }
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    if (warp == 4) {
        asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
        asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(kTmemCols) : "memory");
    }
}

template <int NPAD, int NATOM>
int launch_linear_wide_smallk(const LinearParams& p, cudaStream_t st) {
    constexpr size_t base = (size_t)2 * NATOM * NPAD * 128 + (size_t)2 * 2 * kBM * 128 + 1024;
    constexpr size_t stage = (size_t)4 * 32 * (NPAD + 4) * 4;
    constexpr bool STAGED = base + stage <= 225 * 1024;
    constexpr size_t smem = base + (STAGED ? stage : 0);
    cudaError_t e0 = cudaFuncSetAttribute(linear_tf32x3_wide_smallk_kernel<NPAD, NATOM, STAGED>,
                                          cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    if (e0 != cudaSuccess) return (int)e0;
    const int64_t ntiles = (p.M + kBM - 1) / kBM;
    const unsigned grid = (unsigned)std::min<int64_t>(ntiles, device_sm_count());
    linear_tf32x3_wide_smallk_kernel<NPAD, NATOM, STAGED><<<grid, kWsThreads, smem, st>>>(p);
    cudaError_t e = cudaGetLastError();
    return e == cudaSuccess ? GALA_OK : (int)e;
}

template <int NPAD>
int launch_linear_v2(const LinearParams& p, cudaStream_t st) {
    constexpr size_t smem = 2 * (size_t)(4 * kBM * 128 + 4 * NPAD * 128) + 1024;
    // the shared-memory opt-in is a per-device function attribute and the grid follows the current device's
    // SM count: both are set on every launch (host-side, cheap) so that a process driving several GPUs, or
    // several host threads, needs no shared "configured" state
    cudaError_t e0 = cudaFuncSetAttribute(linear_tf32x3_v2_kernel<NPAD>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    if (e0 != cudaSuccess) return (int)e0;
    const int sms = device_sm_count();
    const int64_t ntiles = (p.M + kBM - 1) / kBM;
    const unsigned grid = (unsigned)std::min<int64_t>(ntiles, sms > 0 ? sms : 148);
    linear_tf32x3_v2_kernel<NPAD><<<grid, kV2Threads, smem, st>>>(p);
    cudaError_t e = cudaGetLastError();
    return e == cudaSuccess ? GALA_OK : (int)e;
}

template <int NPAD>
int launch_linear(const LinearParams& p, cudaStream_t st) {
    constexpr size_t smem = (size_t)kStages * (2 * kBM * 128 + 2 * NPAD * 128) + 1024;
    cudaError_t e0 = cudaFuncSetAttribute(linear_tf32x3_kernel<NPAD>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    if (e0 != cudaSuccess) return (int)e0;
    const unsigned grid = (unsigned)((p.M + kBM - 1) / kBM);
    linear_tf32x3_kernel<NPAD><<<grid, kThreads, smem, st>>>(p);
    cudaError_t e = cudaGetLastError();
    return e == cudaSuccess ? GALA_OK : (int)e;
}

}  // namespace

// Fewest full 128-row tiles for which the persistent kernel (one CTA per SM) is used; below, the 2-CTA/SM kernel runs
// every tile in one wave.  (Measured at the row counts a rank holds on 8 GPUs, profiles/r02_linear_small_m.txt: 29 121
// rows 39.4 us persistent vs 35.1 us one-wave; equal from 58 242 rows on -- the default stays at 4.)
#ifndef GALA_LINEAR_V2_MIN_TILES
#define GALA_LINEAR_V2_MIN_TILES 4
#endif

extern "C" int gala_linear_f32(const float* X, int64_t M, int32_t K, const float* W, const float* bias, int32_t N, float* Y,
                               const float* row_scale, int32_t relu, const float* att_w, const float* att_b,
                               int32_t att_b_on_device, float* att_out, const gala_multi_out_t* multi_out,
                               const gala_multi_out_t* att_multi_out, gala_stream_t stream) {
    if (M < 0 || K <= 0 || N <= 0) return GALA_ERR_BAD_SHAPE;
    if (N > 256) return GALA_ERR_UNSUPPORTED;   // one CTA holds every output column (UMMA_N <= 256)
    if (N > 64 && (att_w || (multi_out && multi_out->count > 0))) return GALA_ERR_UNSUPPORTED;   // row epilogues: N <= 64
    if (M == 0) return GALA_OK;
    const bool multi = multi_out && multi_out->count > 0;
    if (!X || !W || (!Y && !multi) || (att_w && (!att_b || !att_out))) return GALA_ERR_NULL_POINTER;
    if (multi && (multi_out->count > kMaxPeers || (N & 3) != 0)) return GALA_ERR_UNSUPPORTED;
    LinearParams p;
    std::memset(&p, 0, sizeof(p));
    p.X = X;
    p.W = W;
    p.bias = bias;
    p.Y = Y;
    p.row_scale = row_scale;
    p.att_w = att_w;
    p.att_out = att_out;
    if (multi) {
        p.mo.count = multi_out->count;
        p.mo.mc_base = multi_out->multicast_base;
        p.mo.need = multi_out->need_mask;
        for (int q = 0; q < multi_out->count; ++q) p.mo.base[q] = multi_out->base[q];
    }
    if (att_multi_out && att_multi_out->count > 0) {
        if (!att_w || att_multi_out->count > kMaxPeers) return GALA_ERR_UNSUPPORTED;
        p.att_mo.count = att_multi_out->count;
        p.att_mo.mc_base = att_multi_out->multicast_base;
        p.att_mo.need = att_multi_out->need_mask;
        for (int q = 0; q < att_multi_out->count; ++q) p.att_mo.base[q] = att_multi_out->base[q];
    }
    p.M = M;
    p.K = K;
    p.N = N;
    p.relu = relu;
    if (att_w) {   // att_b: two floats, on the host (read here) or on the device (read by the kernel)
        if (att_b_on_device) {
            p.att_b_dev = att_b;
        } else {
            p.att_b0 = att_b[0];
            p.att_b1 = att_b[1];
        }
    }
    cudaStream_t st = reinterpret_cast<cudaStream_t>(stream);
    const int64_t ntiles = (M + kBM - 1) / kBM;
    const bool v2 = (K % 2 == 0) && (reinterpret_cast<uintptr_t>(X) % 8 == 0) && (reinterpret_cast<uintptr_t>(W) % 8 == 0) &&
                    M >= (int64_t)GALA_LINEAR_V2_MIN_TILES * kBM;
    if (v2) {
        p.round_robin = (K >= 256 && ntiles >= 4 * (int64_t)device_sm_count()) ? 1 : 0;
        if (N <= 16) return launch_linear_v2<16>(p, st);
        if (N <= 32) return launch_linear_v2<32>(p, st);
        if (N <= 48) return launch_linear_v2<48>(p, st);
        // N in (48, 64]: the persistent variant would need 229 KB of shared memory; use the 2-CTA/SM kernel
    }
    if (N <= 16) return launch_linear<16>(p, st);
    if (N <= 32) return launch_linear<32>(p, st);
    if (N <= 48) return launch_linear<48>(p, st);
    if (N <= 64) return launch_linear<64>(p, st);
    if (K <= 64 && ntiles >= 2 * (int64_t)device_sm_count()) {
        // wide output, small K (classifier): W resident, persistent, accumulators double-buffered
        if (K <= 32) {
            if (N <= 128) return launch_linear_wide_smallk<128, 1>(p, st);
            if (N <= 176) return launch_linear_wide_smallk<176, 1>(p, st);
            return launch_linear_wide_smallk<256, 1>(p, st);
        }
        if (N <= 128) return launch_linear_wide_smallk<128, 2>(p, st);
        if (N <= 176) return launch_linear_wide_smallk<176, 2>(p, st);
        // N > 176 with K > 32: W alone would take 128 KB next to the A stages -- the general wide kernel below
    }
    if (N <= 96) return launch_linear<96>(p, st);
    if (N <= 128) return launch_linear<128>(p, st);
    if (N <= 176) return launch_linear<176>(p, st);
    return launch_linear<256>(p, st);
}
