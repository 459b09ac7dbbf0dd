// formats.cu -- GPU construction of the graph formats the generated program builds on the
// host today.  Integer work: every output is bit-exact
// against the reference (oracle/_ref) -- checked by tests/test_formats_gpu.py.
//
//   gala_csr_from_coo    CSRCMatrix::build, CSR branch    src/formats/csrc_matrix.h:148-282
//                        (count_atomic / partial_sum / count_sort_place / sort_range,
//                         src/utils/mtx_sort.h:52-64,165-174,114-137,683-722)
//   gala_csr_transpose   buildTranspose                    tests/common.h:107-123
//   gala_col_tile        static_ord_col_breakpoints +      src/ops/tiling.h:1594-1608
//                        ord_col_tiling_torch              src/ops/tiling.h:222-283
//   gala_sample_ab       inplace_sample_graph_ab           src/ops/tiling.h:454-508
//   gala_mask_subgraph   getMaskSubgraphs (one layer)      tests/common.h:20-105
//
// The reference builds a CSR with an atomic scatter and a per-row std::sort on the host
// (1.7 s for the Reddit shape on 8 cores).  Here: one stable LSD radix sort
// of the packed (row << 32 | col) keys, 8 bits per pass over only the bits that are in use,
// values travelling as payload -- HBM-streaming work (16-24 bytes per edge per pass).
#include <algorithm>
#include <cstring>

#include "common.cuh"

using namespace gala;

namespace {

inline cudaStream_t S(gala_stream_t s) { return reinterpret_cast<cudaStream_t>(s); }
inline int last_error() {
    cudaError_t e = cudaGetLastError();
    return e == cudaSuccess ? GALA_OK : (int)e;
}
inline size_t align_up(size_t x, size_t a = 256) { return (x + a - 1) / a * a; }

// ----------------------------------------------------------------------------- scan
constexpr int kScanThreads = 256;
constexpr int kScanItems = 8;
constexpr int kScanTile = kScanThreads * kScanItems;

// exclusive scan of one tile per block; block totals to `sums` (nullable when 1 block)
__global__ void __launch_bounds__(kScanThreads) scan_tile_kernel(const int* __restrict__ in, int* __restrict__ out,
                                                                 int64_t n, int* __restrict__ sums) {
    __shared__ int s_warp[kScanThreads / 32];
    const int64_t base = (int64_t)blockIdx.x * kScanTile + (int64_t)threadIdx.x * kScanItems;
    int v[kScanItems], run = 0;
#pragma unroll
    for (int i = 0; i < kScanItems; ++i) {
        v[i] = base + i < n ? in[base + i] : 0;
        run += v[i];
    }
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    int inc = run;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
        int t = __shfl_up_sync(kFull, inc, o);
        if (lane >= o) inc += t;
    }
    if (lane == 31) s_warp[warp] = inc;
    __syncthreads();
    int wbase = 0, total = 0;
#pragma unroll
    for (int w = 0; w < kScanThreads / 32; ++w) {
        if (w < warp) wbase += s_warp[w];
        total += s_warp[w];
    }
    int ex = wbase + inc - run;
#pragma unroll
    for (int i = 0; i < kScanItems; ++i) {
        if (base + i < n) out[base + i] = ex;
        ex += v[i];
    }
    if (sums && threadIdx.x == 0) sums[blockIdx.x] = total;
}

__global__ void __launch_bounds__(kScanThreads) scan_add_kernel(int* __restrict__ out, int64_t n,
                                                                const int* __restrict__ sums) {
    const int add = sums[blockIdx.x];
    const int64_t base = (int64_t)blockIdx.x * kScanTile + (int64_t)threadIdx.x * kScanItems;
#pragma unroll
    for (int i = 0; i < kScanItems; ++i)
        if (base + i < n) out[base + i] += add;
}

size_t scan_ws_ints(int64_t n) {
    size_t total = 0;
    while (n > kScanTile) {
        n = (n + kScanTile - 1) / kScanTile;
        total += align_up((size_t)n, 64);
    }
    return total + 64;
}

// out[i] = sum_{j<i} in[j]  (in == out allowed); ws holds the block-sum levels
void exclusive_scan(const int* in, int* out, int64_t n, int* ws, cudaStream_t st) {
    if (n <= 0) return;
    const int64_t nb = (n + kScanTile - 1) / kScanTile;
    if (nb == 1) {
        scan_tile_kernel<<<1, kScanThreads, 0, st>>>(in, out, n, nullptr);
        return;
    }
    scan_tile_kernel<<<(unsigned)nb, kScanThreads, 0, st>>>(in, out, n, ws);
    exclusive_scan(ws, ws, nb, ws + align_up((size_t)nb, 64), st);
    scan_add_kernel<<<(unsigned)nb, kScanThreads, 0, st>>>(out, n, ws);
}

// ----------------------------------------------------------------------------- radix sort
constexpr int kSortThreads = 256;
constexpr int kSortItems = 16;
constexpr int kSortTile = kSortThreads * kSortItems;  // 4096 keys per block
constexpr int kSortWarps = kSortThreads / 32;

__global__ void pack_keys_kernel(const int* __restrict__ rows, const int* __restrict__ cols, uint64_t* __restrict__ keys,
                                 int64_t n) {
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x)
        keys[i] = ((uint64_t)(uint32_t)rows[i] << 32) | (uint32_t)cols[i];
}

__global__ void __launch_bounds__(kSortThreads) radix_hist_kernel(const uint64_t* __restrict__ keys, int64_t n, int shift,
                                                                  int* __restrict__ hist, int nb) {
    __shared__ int s_hist[256];
    s_hist[threadIdx.x] = 0;
    __syncthreads();
    const int64_t base = (int64_t)blockIdx.x * kSortTile;
#pragma unroll 4
    for (int i = threadIdx.x; i < kSortTile; i += kSortThreads)
        if (base + i < n) atomicAdd(&s_hist[(keys[base + i] >> shift) & 0xff], 1);
    __syncthreads();
    hist[(int64_t)threadIdx.x * nb + blockIdx.x] = s_hist[threadIdx.x];   // digit-major
}

// Stable scatter of one 4096-key tile.  Warp w owns the contiguous sub-tile
// [w*512, (w+1)*512); key i of lane l is element i*32 + l of it, so (i, lane) order is
// memory order.  Ranks come from __match_any_sync + running per-warp digit counters.
template <bool HAS_VAL>
__global__ void __launch_bounds__(kSortThreads)
radix_scatter_kernel(const uint64_t* __restrict__ keys_in, uint64_t* __restrict__ keys_out,
                     const float* __restrict__ vals_in, float* __restrict__ vals_out, int64_t n, int shift,
                     const int* __restrict__ offs, int nb) {
    __shared__ int s_cnt[kSortWarps][256];
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    for (int i = threadIdx.x; i < kSortWarps * 256; i += kSortThreads) (&s_cnt[0][0])[i] = 0;
    __syncthreads();
    const int64_t base = (int64_t)blockIdx.x * kSortTile + warp * (kSortItems * 32);
    uint64_t key[kSortItems];
    float val[kSortItems];
    unsigned short rank[kSortItems];
    const unsigned lt = (1u << lane) - 1;
#pragma unroll
    for (int i = 0; i < kSortItems; ++i) {
        const int64_t idx = base + i * 32 + lane;
        const bool ok = idx < n;
        key[i] = ok ? keys_in[idx] : 0;
        if (HAS_VAL) val[i] = ok ? vals_in[idx] : 0.0f;
        const int d = ok ? (int)((key[i] >> shift) & 0xff) : (0x100 | lane);   // invalid lanes never match
        const unsigned peers = __match_any_sync(kFull, d);
        int before = 0;
        if (ok) before = s_cnt[warp][d];
        __syncwarp();
        if (ok && (peers & lt) == 0) s_cnt[warp][d] = before + __popc(peers);
        __syncwarp();
        rank[i] = (unsigned short)(before + __popc(peers & lt));
    }
    __syncthreads();
    {   // per digit: exclusive prefix over the warps, offset by the block's global base
        const int d = threadIdx.x;
        int run = offs[(int64_t)d * nb + blockIdx.x];
#pragma unroll
        for (int w = 0; w < kSortWarps; ++w) {
            int c = s_cnt[w][d];
            s_cnt[w][d] = run;
            run += c;
        }
    }
    __syncthreads();
#pragma unroll
    for (int i = 0; i < kSortItems; ++i) {
        const int64_t idx = base + i * 32 + lane;
        if (idx < n) {
            const int d = (int)((key[i] >> shift) & 0xff);
            const int64_t pos = (int64_t)s_cnt[warp][d] + rank[i];
            keys_out[pos] = key[i];
            if (HAS_VAL) vals_out[pos] = val[i];
        }
    }
}

int bits_for(uint32_t n) {   // bits needed for values in [0, n)
    int b = 0;
    while (b < 32 && (n == 0 ? 0u : (n - 1)) >> b) ++b;
    return b;
}

struct SortBuffers {
    uint64_t *ka, *kb;
    float *va, *vb;
    int *hist, *scan_ws;
};

size_t sort_ws_bytes(int64_t n) {
    const int64_t nb = std::max<int64_t>((n + kSortTile - 1) / kSortTile, 1);
    return 2 * align_up((size_t)n * 8) + 2 * align_up((size_t)n * 4) + align_up((size_t)nb * 256 * 4) +
           align_up(scan_ws_ints(nb * 256) * 4) + 1024;
}

SortBuffers carve(void* ws, int64_t n) {
    const int64_t nb = std::max<int64_t>((n + kSortTile - 1) / kSortTile, 1);
    char* p = static_cast<char*>(ws);
    SortBuffers b;
    b.ka = reinterpret_cast<uint64_t*>(p); p += align_up((size_t)n * 8);
    b.kb = reinterpret_cast<uint64_t*>(p); p += align_up((size_t)n * 8);
    b.va = reinterpret_cast<float*>(p); p += align_up((size_t)n * 4);
    b.vb = reinterpret_cast<float*>(p); p += align_up((size_t)n * 4);
    b.hist = reinterpret_cast<int*>(p); p += align_up((size_t)nb * 256 * 4);
    b.scan_ws = reinterpret_cast<int*>(p);
    return b;
}

// sorts b.ka (+ b.va) by the low `bits_lo` bits and bits [32, 32+bits_hi); result pointer returned
void radix_sort_pairs(SortBuffers& b, int64_t n, int bits_lo, int bits_hi, bool has_val, cudaStream_t st,
                      uint64_t** keys_sorted, float** vals_sorted) {
    const int nb = (int)((n + kSortTile - 1) / kSortTile);
    uint64_t *kin = b.ka, *kout = b.kb;
    float *vin = b.va, *vout = b.vb;
    auto pass = [&](int shift) {
        radix_hist_kernel<<<nb, kSortThreads, 0, st>>>(kin, n, shift, b.hist, nb);
        exclusive_scan(b.hist, b.hist, (int64_t)nb * 256, b.scan_ws, st);
        if (has_val)
            radix_scatter_kernel<true><<<nb, kSortThreads, 0, st>>>(kin, kout, vin, vout, n, shift, b.hist, nb);
        else
            radix_scatter_kernel<false><<<nb, kSortThreads, 0, st>>>(kin, kout, nullptr, nullptr, n, shift, b.hist, nb);
        std::swap(kin, kout);
        std::swap(vin, vout);
    };
    for (int s = 0; s < bits_lo; s += 8) pass(s);
    for (int s = 0; s < bits_hi; s += 8) pass(32 + s);
    *keys_sorted = kin;
    *vals_sorted = vin;
}

// offsets[r] = first position whose row is >= r; ids[e] = low word of the key
__global__ void unpack_sorted_kernel(const uint64_t* __restrict__ keys, int64_t n, int nrows, int* __restrict__ offsets,
                                     int* __restrict__ ids) {
    for (int64_t e = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; e < n; e += (int64_t)gridDim.x * blockDim.x) {
        // ids outside [0, nrows) are a caller error (the reference corrupts memory on them, count_atomic
        // mtx_sort.h:52-64); the clamp keeps the row-pointer writes inside offsets[0..nrows]
        const uint64_t k = keys[e];
        const int r = min((int)(uint32_t)(k >> 32), nrows - 1);
        const int p = e > 0 ? min((int)(uint32_t)(keys[e - 1] >> 32), nrows - 1) : -1;
        ids[e] = (int)(uint32_t)k;
        for (int q = p + 1; q <= r; ++q) offsets[q] = (int)e;
        if (e == n - 1)
            for (int q = r + 1; q <= nrows; ++q) offsets[q] = (int)n;
    }
}

__global__ void fill_int_kernel(int* p, int64_t n, int v) {
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x) p[i] = v;
}

// get_sids (src/formats/csrc_matrix.h:399-411) fused with the key packing of the transpose:
// key = (col << 32) | row
__global__ void pack_transposed_kernel(const int* __restrict__ offsets, const int* __restrict__ ids, int nrows,
                                       uint64_t* __restrict__ keys) {
    const int lane = threadIdx.x & 31;
    const int row = (blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    if (row >= nrows) return;
    for (int e = offsets[row] + lane; e < offsets[row + 1]; e += 32)
        keys[e] = ((uint64_t)(uint32_t)ids[e] << 32) | (uint32_t)row;
}

unsigned grid_for(int64_t n, int threads = 256) { return (unsigned)std::min<int64_t>((n + threads - 1) / threads, 148 * 32); }

// ----------------------------------------------------------------------------- column tiling
// counts[s*N + i] = #columns of row i inside [s*T, (s+1)*T)   (rows are column-sorted)
__global__ void tile_count_kernel(const int* __restrict__ offsets, const int* __restrict__ ids, int nrows, int S, int T,
                                  int* __restrict__ counts, int* __restrict__ row_seg_start) {
    const int64_t t = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (t >= (int64_t)nrows * S) return;
    const int s = (int)(t / nrows), i = (int)(t % nrows);
    const int b = offsets[i], e = offsets[i + 1];
    auto lower = [&](int64_t key) {   // first position in [b,e) with ids >= key
        int lo = b, hi = e;
        while (lo < hi) {
            int mid = (lo + hi) >> 1;
            if ((int64_t)ids[mid] < key) lo = mid + 1; else hi = mid;
        }
        return lo;
    };
    const int p0 = lower((int64_t)s * T), p1 = lower((int64_t)(s + 1) * T);
    counts[t] = p1 - p0;
    row_seg_start[t] = p0;
}

__global__ void tile_copy_kernel(const int* __restrict__ ids, const float* __restrict__ vals, int nrows, int S,
                                 const int* __restrict__ pos, const int* __restrict__ row_seg_start, int64_t nvals,
                                 int* __restrict__ out_offsets, int* __restrict__ out_cols, float* __restrict__ out_vals) {
    const int lane = threadIdx.x & 31;
    const int64_t t = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5;   // one warp per (segment, row)
    if (t >= (int64_t)nrows * S) return;
    const int s = (int)(t / nrows), i = (int)(t % nrows);
    const int seg0 = pos[(int64_t)s * nrows];
    const int dst = pos[t];
    const int end = t + 1 < (int64_t)nrows * S ? pos[t + 1] : (int)nvals;
    const int src = row_seg_start[t];
    if (lane == 0) {
        out_offsets[(int64_t)s * (nrows + 1) + i] = dst - seg0;
        if (i == nrows - 1) out_offsets[(int64_t)s * (nrows + 1) + nrows] = end - seg0;
    }
    for (int k = lane; k < end - dst; k += 32) {
        out_cols[dst + k] = ids[src + k];
        out_vals[dst + k] = vals[src + k];
    }
}

__global__ void tile_bounds_kernel(const int* __restrict__ pos, int nrows, int S, int64_t nvals, int* __restrict__ bounds) {
    const int s = blockIdx.x * blockDim.x + threadIdx.x;
    if (s >= S) return;
    bounds[2 * s] = pos[(int64_t)s * nrows];
    bounds[2 * s + 1] = s + 1 < S ? pos[(int64_t)(s + 1) * nrows] : (int)nvals;
}

// ----------------------------------------------------------------------------- a/b sampling
constexpr int kMaxSample = 128;
__global__ void sample_ab_kernel(const int* __restrict__ offsets, const int* __restrict__ ids, const float* __restrict__ vals,
                                 int nrows, int s, int ra, int rb, int* __restrict__ new_off, int* __restrict__ new_ids,
                                 float* __restrict__ new_vals, int* __restrict__ bad) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i == 0) new_off[0] = 0;
    if (i >= nrows) return;
    const int first = offsets[i], total = offsets[i + 1] - first;
    new_off[i + 1] = (i + 1) * s;
    if (total <= 0) {   // `% 0` in the reference (tiling.h:482): undefined there, flagged here
        *bad = 1;
        return;
    }
    int e_used[kMaxSample];
    for (int ji = 0; ji < s; ++ji) {   // insertion sort of (ra*ji+rb) % total, int arithmetic as the reference
        int j = (ra * ji + rb) % total, k = ji - 1;
        while (k >= 0 && e_used[k] > j) {
            e_used[k + 1] = e_used[k];
            --k;
        }
        e_used[k + 1] = j;
    }
    for (int j = 0; j < s; ++j) {
        new_ids[(int64_t)i * s + j] = ids[first + e_used[j]];
        new_vals[(int64_t)i * s + j] = vals[first + e_used[j]];
    }
}

// ----------------------------------------------------------------------------- mask sub-graphs
__global__ void mask_count_kernel(const int* __restrict__ offsets, const unsigned char* __restrict__ mask, int nrows,
                                  int* __restrict__ counts) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i <= nrows) counts[i] = (i < nrows && mask[i] > 0) ? offsets[i + 1] - offsets[i] : 0;
}

__global__ void mask_copy_kernel(const int* __restrict__ offsets, const int* __restrict__ ids, const float* __restrict__ vals,
                                 const unsigned char* __restrict__ mask, int nrows, const int* __restrict__ new_off,
                                 int* __restrict__ new_ids, float* __restrict__ new_vals) {
    const int lane = threadIdx.x & 31;
    const int row = (blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    if (row >= nrows || mask[row] == 0) return;
    const int src = offsets[row], dst = new_off[row], n = offsets[row + 1] - src;
    for (int k = lane; k < n; k += 32) {
        new_ids[dst + k] = ids[src + k];
        new_vals[dst + k] = vals[src + k];
    }
}

// next[v] = max(0, max_{u in N(v)} mask[u]) -- gSpMM with maxAgg into a zeroed buffer
__global__ void mask_next_kernel(const int* __restrict__ offsets, const int* __restrict__ ids,
                                 const unsigned char* __restrict__ mask, int nrows, unsigned char* __restrict__ next) {
    const int lane = threadIdx.x & 31;
    const int row = (blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    if (row >= nrows) return;
    int m = 0;
    for (int e = offsets[row] + lane; e < offsets[row + 1]; e += 32) m = max(m, (int)mask[ids[e]]);
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) m = max(m, __shfl_xor_sync(kFull, m, o));
    if (lane == 0) next[row] = (unsigned char)m;
}

// ----------------------------------------------------------------------------- reordering
// rowReorderToAdj (src/ops/reordering.h:940-1013): key = (perm[row] << 32) | perm[col]
__global__ void pack_permuted_kernel(const int* __restrict__ offsets, const int* __restrict__ ids, const int* __restrict__ perm,
                                     int nrows, uint64_t* __restrict__ keys) {
    const int lane = threadIdx.x & 31;
    const int row = (blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    if (row >= nrows) return;
    // perm is required to be a permutation of [0, nrows); stray entries are clamped so that nothing is
    // written out of bounds (the reference skips such rows and leaves their slots uninitialised)
    auto at = [&](int i) { return (uint32_t)min(max(perm[min(max(i, 0), nrows - 1)], 0), nrows - 1); };
    const uint64_t hi = (uint64_t)at(row) << 32;
    for (int e = offsets[row] + lane; e < offsets[row + 1]; e += 32) keys[e] = hi | at(ids[e]);
}

// The reference sorts each row's (column, value) PAIRS (reordering.h:1000), so duplicate edges end up ordered by
// value; the radix sort above is stable in the source order.  One thread per run of equal (row, column) keys
// re-orders that run's values (runs are rare and short).
__global__ void sort_duplicate_runs_kernel(const uint64_t* __restrict__ keys, float* __restrict__ vals, int64_t n) {
    for (int64_t e = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; e + 1 < n; e += (int64_t)gridDim.x * blockDim.x) {
        const uint64_t k = keys[e];
        if (keys[e + 1] != k || (e > 0 && keys[e - 1] == k)) continue;   // not the head of a run
        int64_t end = e + 2;
        while (end < n && keys[end] == k) ++end;
        for (int64_t i = e + 1; i < end; ++i) {   // insertion sort, ascending (std::pair operator<)
            const float v = vals[i];
            int64_t j = i;
            while (j > e && v < vals[j - 1]) {
                vals[j] = vals[j - 1];
                --j;
            }
            vals[j] = v;
        }
    }
}

// rowPermuteDenseTo (reordering.h:244-283): Y[perm[i], :] = X[i, :];  rowPermuteDenseFrom (:207-236): Y[i, :] = X[perm[i], :]
template <typename V>
__global__ void permute_rows_kernel(const V* __restrict__ X, const int* __restrict__ perm, V* __restrict__ Y, int nrows,
                                    int kv, int from) {
    const int64_t total = (int64_t)nrows * kv;
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (int64_t)gridDim.x * blockDim.x) {
        const int r = (int)(i / kv), c = (int)(i - (int64_t)r * kv);
        const int p = min(max(perm[r], 0), nrows - 1);   // memory-safe for stray entries
        const int64_t src = from ? (int64_t)p * kv + c : i;
        const int64_t dst = from ? i : (int64_t)p * kv + c;
        Y[dst] = X[src];
    }
}

// degree-descending order: key = (~degree << 32) | row  -> ascending sort puts the widest rows first, ties by row id
__global__ void degree_keys_kernel(const int* __restrict__ offsets, int nrows, uint64_t* __restrict__ keys) {
    for (int r = blockIdx.x * blockDim.x + threadIdx.x; r < nrows; r += gridDim.x * blockDim.x)
        keys[r] = ((uint64_t)(0x7fffffffu - (uint32_t)(offsets[r + 1] - offsets[r])) << 32) | (uint32_t)r;
}
__global__ void invert_order_kernel(const uint64_t* __restrict__ keys, int nrows, int* __restrict__ perm,
                                    int* __restrict__ order) {
    for (int k = blockIdx.x * blockDim.x + threadIdx.x; k < nrows; k += gridDim.x * blockDim.x) {
        const int r = (int)(uint32_t)keys[k];
        perm[r] = k;
        if (order) order[k] = r;
    }
}

int csr_from_keys(SortBuffers& b, int32_t nrows, int32_t ncols, int64_t nvals, bool has_val, int32_t* offsets, int32_t* ids,
                  float* out_vals, cudaStream_t st) {
    uint64_t* ks;
    float* vs;
    radix_sort_pairs(b, nvals, bits_for((uint32_t)ncols), bits_for((uint32_t)nrows), has_val, st, &ks, &vs);
    unpack_sorted_kernel<<<grid_for(nvals), 256, 0, st>>>(ks, nvals, nrows, offsets, ids);
    if (has_val && out_vals) {
        cudaError_t e = cudaMemcpyAsync(out_vals, vs, (size_t)nvals * 4, cudaMemcpyDeviceToDevice, st);
        if (e != cudaSuccess) return (int)e;
    }
    return last_error();
}

}  // namespace

extern "C" {

size_t gala_csr_from_coo_workspace_bytes(int32_t nrows, int32_t ncols, int64_t nvals) {
    (void)nrows;
    (void)ncols;
    return sort_ws_bytes(std::max<int64_t>(nvals, 1));
}

int gala_csr_from_coo(int32_t nrows, int32_t ncols, int64_t nvals, const int32_t* row_ids, const int32_t* col_ids,
                      const float* vals, int32_t* offsets, int32_t* ids, float* out_vals, void* workspace,
                      size_t workspace_bytes, gala_stream_t stream) {
    if (nrows < 0 || ncols < 0 || nvals < 0) return GALA_ERR_BAD_SHAPE;
    if (nvals > 0x7fffffffLL) return GALA_ERR_UNSUPPORTED;
    if (!offsets || (nvals > 0 && (!row_ids || !col_ids || !ids || !workspace))) return GALA_ERR_NULL_POINTER;
    if (nvals > 0 && workspace_bytes < gala_csr_from_coo_workspace_bytes(nrows, ncols, nvals)) return GALA_ERR_WORKSPACE;
    cudaStream_t st = S(stream);
    if (nvals == 0) {
        fill_int_kernel<<<grid_for(nrows + 1), 256, 0, st>>>(offsets, (int64_t)nrows + 1, 0);
        return last_error();
    }
    SortBuffers b = carve(workspace, nvals);
    pack_keys_kernel<<<grid_for(nvals), 256, 0, st>>>(row_ids, col_ids, b.ka, nvals);
    const bool has_val = vals != nullptr && out_vals != nullptr;
    if (has_val) {
        cudaError_t e = cudaMemcpyAsync(b.va, vals, (size_t)nvals * 4, cudaMemcpyDeviceToDevice, st);
        if (e != cudaSuccess) return (int)e;
    }
    return csr_from_keys(b, nrows, ncols, nvals, has_val, offsets, ids, out_vals, st);
}

int gala_csr_transpose(int32_t nrows, int32_t ncols, int64_t nvals, const int32_t* offsets, const int32_t* ids,
                       const float* vals, int32_t* t_offsets, int32_t* t_ids, float* t_vals, void* workspace,
                       size_t workspace_bytes, gala_stream_t stream) {
    if (nrows < 0 || ncols < 0 || nvals < 0) return GALA_ERR_BAD_SHAPE;
    if (nvals > 0x7fffffffLL) return GALA_ERR_UNSUPPORTED;
    if (!t_offsets || (nrows > 0 && !offsets) || (nvals > 0 && (!ids || !t_ids || !workspace))) return GALA_ERR_NULL_POINTER;
    if (nvals > 0 && workspace_bytes < gala_csr_from_coo_workspace_bytes(ncols, nrows, nvals)) return GALA_ERR_WORKSPACE;
    cudaStream_t st = S(stream);
    if (nvals == 0) {
        fill_int_kernel<<<grid_for(ncols + 1), 256, 0, st>>>(t_offsets, (int64_t)ncols + 1, 0);
        return last_error();
    }
    SortBuffers b = carve(workspace, nvals);
    pack_transposed_kernel<<<(unsigned)(((int64_t)nrows * 32 + 255) / 256), 256, 0, st>>>(offsets, ids, nrows, b.ka);
    const bool has_val = vals != nullptr && t_vals != nullptr;
    if (has_val) {
        cudaError_t e = cudaMemcpyAsync(b.va, vals, (size_t)nvals * 4, cudaMemcpyDeviceToDevice, st);
        if (e != cudaSuccess) return (int)e;
    }
    // rows of the transpose = columns of the source, and vice versa
    return csr_from_keys(b, ncols, nrows, nvals, has_val, t_offsets, t_ids, t_vals, st);
}

int gala_csr_reorder(int32_t nrows, int64_t nvals, const int32_t* offsets, const int32_t* ids, const float* vals,
                     const int32_t* perm, int32_t* new_offsets, int32_t* new_ids, float* new_vals, void* workspace,
                     size_t workspace_bytes, gala_stream_t stream) {
    if (nrows < 0 || nvals < 0) return GALA_ERR_BAD_SHAPE;
    if (nvals > 0x7fffffffLL) return GALA_ERR_UNSUPPORTED;
    if (!new_offsets || (nrows > 0 && (!offsets || !perm)) || (nvals > 0 && (!ids || !new_ids || !workspace))) return GALA_ERR_NULL_POINTER;
    if (nvals > 0 && workspace_bytes < gala_csr_from_coo_workspace_bytes(nrows, nrows, nvals)) return GALA_ERR_WORKSPACE;
    cudaStream_t st = S(stream);
    if (nvals == 0) {
        fill_int_kernel<<<grid_for(nrows + 1), 256, 0, st>>>(new_offsets, (int64_t)nrows + 1, 0);
        return last_error();
    }
    SortBuffers b = carve(workspace, nvals);
    pack_permuted_kernel<<<(unsigned)(((int64_t)nrows * 32 + 255) / 256), 256, 0, st>>>(offsets, ids, perm, nrows, b.ka);
    const bool has_val = vals != nullptr && new_vals != nullptr;
    if (has_val) {
        cudaError_t e = cudaMemcpyAsync(b.va, vals, (size_t)nvals * 4, cudaMemcpyDeviceToDevice, st);
        if (e != cudaSuccess) return (int)e;
    }
    uint64_t* ks;
    float* vs;
    radix_sort_pairs(b, nvals, bits_for((uint32_t)nrows), bits_for((uint32_t)nrows), has_val, st, &ks, &vs);
    if (has_val) sort_duplicate_runs_kernel<<<grid_for(nvals), 256, 0, st>>>(ks, vs, nvals);
    unpack_sorted_kernel<<<grid_for(nvals), 256, 0, st>>>(ks, nvals, nrows, new_offsets, new_ids);
    if (has_val) {
        cudaError_t e = cudaMemcpyAsync(new_vals, vs, (size_t)nvals * 4, cudaMemcpyDeviceToDevice, st);
        if (e != cudaSuccess) return (int)e;
    }
    return last_error();
}

int gala_permute_rows_f32(const float* X, const int32_t* perm, float* Y, int32_t nrows, int32_t K, int32_t from,
                          gala_stream_t stream) {
    if (nrows < 0 || K < 0) return GALA_ERR_BAD_SHAPE;
    if (nrows == 0 || K == 0) return GALA_OK;
    if (!X || !perm || !Y) return GALA_ERR_NULL_POINTER;
    if (X == Y) return GALA_ERR_UNSUPPORTED;   // out of place only
    cudaStream_t st = S(stream);
    const bool v4 = K % 4 == 0 && (reinterpret_cast<uintptr_t>(X) | reinterpret_cast<uintptr_t>(Y)) % 16 == 0;
    if (v4)
        permute_rows_kernel<float4><<<grid_for((int64_t)nrows * (K / 4)), 256, 0, st>>>(
            reinterpret_cast<const float4*>(X), perm, reinterpret_cast<float4*>(Y), nrows, K / 4, from != 0);
    else
        permute_rows_kernel<float><<<grid_for((int64_t)nrows * K), 256, 0, st>>>(X, perm, Y, nrows, K, from != 0);
    return last_error();
}

size_t gala_degree_order_workspace_bytes(int32_t nrows) { return sort_ws_bytes(std::max<int64_t>(nrows, 1)); }

int gala_degree_order(int32_t nrows, const int32_t* offsets, int32_t* perm, int32_t* order, void* workspace,
                      size_t workspace_bytes, gala_stream_t stream) {
    if (nrows < 0) return GALA_ERR_BAD_SHAPE;
    if (nrows == 0) return GALA_OK;
    if (!offsets || !perm || !workspace) return GALA_ERR_NULL_POINTER;
    if (workspace_bytes < gala_degree_order_workspace_bytes(nrows)) return GALA_ERR_WORKSPACE;
    cudaStream_t st = S(stream);
    SortBuffers b = carve(workspace, nrows);
    degree_keys_kernel<<<grid_for(nrows), 256, 0, st>>>(offsets, nrows, b.ka);
    uint64_t* ks;
    float* vs;
    // the row id in the low word is already ascending: only the degree word needs sorting (stable)
    radix_sort_pairs(b, nrows, 0, 32, false, st, &ks, &vs);
    invert_order_kernel<<<grid_for(nrows), 256, 0, st>>>(ks, nrows, perm, order);
    return last_error();
}

int32_t gala_col_tile_segments(int32_t ncols, int32_t cols_per_partition) {
    if (ncols <= 0 || cols_per_partition <= 0) return 0;
    return (int32_t)(((int64_t)ncols + cols_per_partition - 1) / cols_per_partition);
}

size_t gala_col_tile_workspace_bytes(int32_t nrows, int32_t ncols, int32_t cols_per_partition) {
    const int64_t S = gala_col_tile_segments(ncols, cols_per_partition);
    const int64_t m = std::max<int64_t>((int64_t)nrows * S, 1);
    return 2 * align_up((size_t)m * 4) + align_up(scan_ws_ints(m) * 4) + align_up((size_t)2 * std::max<int64_t>(S, 1) * 4) + 1024;
}

int gala_col_tile(int32_t nrows, int32_t ncols, int64_t nvals, const int32_t* offsets, const int32_t* ids,
                  const float* vals, int32_t cols_per_partition, int32_t* out_offsets, int32_t* out_cols, float* out_vals,
                  int32_t* bounds_host, void* workspace, size_t workspace_bytes, gala_stream_t stream) {
    if (nrows < 0 || ncols <= 0 || nvals < 0 || cols_per_partition <= 0) return GALA_ERR_BAD_SHAPE;
    if (nvals > 0x7fffffffLL) return GALA_ERR_UNSUPPORTED;
    const int nseg = gala_col_tile_segments(ncols, cols_per_partition);
    if ((int64_t)nrows * nseg > 0x7fffffffLL) return GALA_ERR_UNSUPPORTED;
    if (!out_offsets || !bounds_host || !workspace || (nrows > 0 && !offsets) || (nvals > 0 && (!ids || !vals || !out_cols || !out_vals)))
        return GALA_ERR_NULL_POINTER;
    if (workspace_bytes < gala_col_tile_workspace_bytes(nrows, ncols, cols_per_partition)) return GALA_ERR_WORKSPACE;
    cudaStream_t st = S(stream);
    const int64_t m = (int64_t)nrows * nseg;
    char* p = static_cast<char*>(workspace);
    int* counts = reinterpret_cast<int*>(p); p += align_up((size_t)std::max<int64_t>(m, 1) * 4);
    int* starts = reinterpret_cast<int*>(p); p += align_up((size_t)std::max<int64_t>(m, 1) * 4);
    int* scan_ws = reinterpret_cast<int*>(p); p += align_up(scan_ws_ints(std::max<int64_t>(m, 1)) * 4);
    int* bounds_dev = reinterpret_cast<int*>(p);
    if (nrows == 0) {
        for (int s = 0; s < nseg; ++s) bounds_host[2 * s] = bounds_host[2 * s + 1] = 0;
        fill_int_kernel<<<1, 256, 0, st>>>(out_offsets, nseg, 0);
        return last_error();
    }
    tile_count_kernel<<<(unsigned)((m + 255) / 256), 256, 0, st>>>(offsets, ids, nrows, nseg, cols_per_partition, counts, starts);
    exclusive_scan(counts, counts, m, scan_ws, st);   // counts -> global write positions
    tile_copy_kernel<<<(unsigned)((m * 32 + 255) / 256), 256, 0, st>>>(ids, vals, nrows, nseg, counts, starts, nvals, out_offsets,
                                                                        out_cols, out_vals);
    tile_bounds_kernel<<<(nseg + 255) / 256, 256, 0, st>>>(counts, nrows, nseg, nvals, bounds_dev);
    if (int rc = last_error()) return rc;
    cudaError_t e = cudaMemcpyAsync(bounds_host, bounds_dev, (size_t)2 * nseg * 4, cudaMemcpyDeviceToHost, st);
    if (e != cudaSuccess) return (int)e;
    e = cudaStreamSynchronize(st);   // bounds live on the host in this layout (cuda.h:472-475)
    return e == cudaSuccess ? GALA_OK : (int)e;
}

int gala_sample_ab(int32_t nrows, const int32_t* offsets, const int32_t* ids, const float* vals, int32_t sample_size,
                   int32_t ra, int32_t rb, int32_t* new_offsets, int32_t* new_ids, float* new_vals, int32_t* status_dev,
                   gala_stream_t stream) {
    if (nrows < 0 || sample_size < 0) return GALA_ERR_BAD_SHAPE;
    if (sample_size > kMaxSample || (int64_t)nrows * sample_size > 0x7fffffffLL) return GALA_ERR_UNSUPPORTED;
    if (!new_offsets || !status_dev || (nrows > 0 && (!offsets || !ids || !vals || (sample_size > 0 && (!new_ids || !new_vals)))))
        return GALA_ERR_NULL_POINTER;
    cudaStream_t st = S(stream);
    cudaError_t e = cudaMemsetAsync(status_dev, 0, sizeof(int), st);
    if (e != cudaSuccess) return (int)e;
    sample_ab_kernel<<<(std::max(nrows, 1) + 127) / 128, 128, 0, st>>>(offsets, ids, vals, nrows, sample_size, ra, rb, new_offsets,
                                                                        new_ids, new_vals, status_dev);
    return last_error();
}

size_t gala_mask_subgraph_workspace_bytes(int32_t nrows) {
    return align_up(((size_t)nrows + 2) * 4) + align_up(scan_ws_ints((int64_t)nrows + 1) * 4) + 1024;
}

int gala_mask_subgraph(int32_t nrows, const int32_t* offsets, const int32_t* ids, const float* vals, const uint8_t* mask,
                       int32_t* new_offsets, int32_t* new_ids, float* new_vals, int64_t* new_nvals_host, uint8_t* next_mask,
                       void* workspace, size_t workspace_bytes, gala_stream_t stream) {
    if (nrows < 0) return GALA_ERR_BAD_SHAPE;
    if (!new_offsets || !new_nvals_host || !workspace || (nrows > 0 && (!offsets || !mask))) return GALA_ERR_NULL_POINTER;
    if (workspace_bytes < gala_mask_subgraph_workspace_bytes(nrows)) return GALA_ERR_WORKSPACE;
    cudaStream_t st = S(stream);
    char* p = static_cast<char*>(workspace);
    int* scan_ws = reinterpret_cast<int*>(p + align_up(((size_t)nrows + 2) * 4));
    mask_count_kernel<<<(nrows + 1 + 255) / 256, 256, 0, st>>>(offsets, mask, nrows, new_offsets);
    exclusive_scan(new_offsets, new_offsets, (int64_t)nrows + 1, scan_ws, st);
    if (nrows > 0) {
        mask_copy_kernel<<<(unsigned)(((int64_t)nrows * 32 + 255) / 256), 256, 0, st>>>(offsets, ids, vals, mask, nrows, new_offsets,
                                                                                     new_ids, new_vals);
        if (next_mask)
            mask_next_kernel<<<(unsigned)(((int64_t)nrows * 32 + 255) / 256), 256, 0, st>>>(offsets, ids, mask, nrows, next_mask);
    }
    if (int rc = last_error()) return rc;
    int total = 0;
    cudaError_t e = cudaMemcpyAsync(&total, new_offsets + nrows, sizeof(int), cudaMemcpyDeviceToHost, st);
    if (e != cudaSuccess) return (int)e;
    e = cudaStreamSynchronize(st);
    if (e != cudaSuccess) return (int)e;
    *new_nvals_host = total;
    return GALA_OK;
}

}  // extern "C"
