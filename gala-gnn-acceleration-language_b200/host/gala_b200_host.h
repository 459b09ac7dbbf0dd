// gala_b200_host.h -- the host / data-preparation side of a generated gala.cu, B200-native.
//
// The reference's generated program #includes the header-only templates of src/formats, src/ops/tiling.h,
// src/utils and tests/common.h (cuda.h:959-975) and prepares every graph on the CPU: CSRCMatrix::build
// (atomic count + serial prefix + per-row std::sort, src/formats/csrc_matrix.h:148-376), the single-threaded
// ord_col_tiling_torch (src/ops/tiling.h:222-283), inplace_sample_graph_ab (:454-508), getMaskSubgraphs
// (tests/common.h:20-105), then cudaMemcpy's the results (cuda.h:1092-1300).  The retargeted generator emits
// `#include "gala_b200_host.h"` instead of those headers: the SAME class and function names (the emitted main
// is unchanged text), but the sparse matrices live on the GPU from the moment the .npy edge lists are read and
// every format is built there through the C-ABI (gala_csr_from_coo, gala_col_tile, gala_sample_ab,
// gala_mask_subgraph, gala_csr_transpose) -- integer outputs bit-identical to the reference's
// (tests/test_formats_gpu.py), pushed into the same global_* slots.
//
//   reference (tests/common.h, src/formats, src/ops/tiling.h)           here
//   CSRCMatrix<I,N,V>           csrc_matrix.h:43-479                    same name: device CSR (torch CUDA tensors)
//   DenseMatrix<I,N,V>          dense_matrix.h:9-216                    same name: pinned host rows (features, labels, masks)
//   readSM_npy32 / readDM_npy   tests/common.h:331-389 (+ libnpy)       same names: .npy parsed here, edge lists -> device
//   repopulate                  tests/common.h:125-139                  same
//   static_ord_col_breakpoints  tiling.h:1594-1608                      same
//   ord_col_tiling_torch        tiling.h:222-283                        same signature; the four tensors come back as
//                                                                       CUDA tensors (bounds stays on the CPU)
//   inplace_sample_graph_ab     tiling.h:454-508                        same
//   getMaskSubgraphs            tests/common.h:20-105                   same (mask buffer zero-initialised)
//   get_time / calc_mean        threading_utils.h:5, tests/common.h:621 same
//
// The emitted transfer code (cuda.h:1092-1300) cudaMemcpy's `offset_ptr_*`, `col_ptr_*`, `val_ptr_*` with
// cudaMemcpyHostToDevice; those pointers are device pointers here, so that constant is mapped to
// cudaMemcpyDefault (the direction is inferred from the pointers, unified addressing) -- the copies of the host
// feature / label / mask arrays keep working unchanged.
#pragma once
#include <c10/cuda/CUDAStream.h>
#include <cuda_runtime_api.h>
#include <fcntl.h>
#include <sys/mman.h>
#include <sys/stat.h>
#include <torch/torch.h>
#include <unistd.h>

#include <chrono>
#include <cstring>
#include <stdexcept>
#include <string>
#include <vector>

#include "gala_b200.h"

#define cudaMemcpyHostToDevice cudaMemcpyDefault

namespace gala_b200 {
namespace host {

inline void check(int rc, const char* what) {
    if (rc != 0) throw std::runtime_error(std::string(what) + ": " + gala_b200_error_string(rc));
}
inline void cuda_check(cudaError_t e, const char* what) {
    if (e != cudaSuccess) throw std::runtime_error(std::string(what) + ": " + cudaGetErrorString(e));
}
inline torch::TensorOptions dev_i32() { return torch::TensorOptions().dtype(torch::kInt).device(torch::kCUDA, 0); }
inline torch::TensorOptions dev_f32() { return torch::TensorOptions().dtype(torch::kFloat).device(torch::kCUDA, 0); }
inline torch::TensorOptions dev_u8() { return torch::TensorOptions().dtype(torch::kUInt8).device(torch::kCUDA, 0); }
inline gala_stream_t stream() { return (gala_stream_t)c10::cuda::getCurrentCUDAStream().stream(); }

// ---- .npy (format 1.0 / 2.0 / 3.0), memory-mapped; dtype-strict like the vendored loader (npy.hpp:527-553) ----
struct NpyFile {
    void* map = nullptr;
    size_t map_bytes = 0;
    const char* data = nullptr;
    std::vector<unsigned long> shape;
    std::string descr;
    size_t count = 1;

    explicit NpyFile(const std::string& path) {
        int fd = ::open(path.c_str(), O_RDONLY);
        if (fd < 0) throw std::runtime_error("cannot open " + path);
        struct stat st;
        fstat(fd, &st);
        map_bytes = (size_t)st.st_size;
        map = mmap(nullptr, map_bytes, PROT_READ, MAP_PRIVATE, fd, 0);
        ::close(fd);
        if (map == MAP_FAILED) throw std::runtime_error("cannot map " + path);
        const unsigned char* p = static_cast<const unsigned char*>(map);
        if (map_bytes < 12 || std::memcmp(p, "\x93NUMPY", 6) != 0) throw std::runtime_error(path + ": not a .npy file");
        size_t hlen, hoff;
        if (p[6] == 1) {
            hlen = p[8] | (p[9] << 8);
            hoff = 10;
        } else {
            hlen = p[8] | (p[9] << 8) | (p[10] << 16) | ((size_t)p[11] << 24);
            hoff = 12;
        }
        std::string h(reinterpret_cast<const char*>(p) + hoff, hlen);
        auto field = [&](const std::string& key) {
            size_t k = h.find("'" + key + "'");
            if (k == std::string::npos) throw std::runtime_error(path + ": header lacks " + key);
            return h.find(':', k) + 1;
        };
        size_t d = h.find('\'', field("descr"));
        descr = h.substr(d + 1, h.find('\'', d + 1) - d - 1);
        if (h.compare(h.find_first_not_of(' ', field("fortran_order")), 4, "True") == 0)
            throw std::runtime_error(path + ": fortran_order arrays are not supported");
        size_t s0 = h.find('(', field("shape")), s1 = h.find(')', s0);
        std::string sh = h.substr(s0 + 1, s1 - s0 - 1);
        for (size_t i = 0; i < sh.size();) {
            if (isdigit((unsigned char)sh[i])) {
                size_t j = i;
                while (j < sh.size() && isdigit((unsigned char)sh[j])) ++j;
                shape.push_back(std::stoul(sh.substr(i, j - i)));
                i = j;
            } else {
                ++i;
            }
        }
        for (auto v : shape) count *= v;
        data = reinterpret_cast<const char*>(p) + hoff + hlen;
    }
    ~NpyFile() {
        if (map && map != MAP_FAILED) munmap(map, map_bytes);
    }
    void expect(const char* want, const std::string& path) const {
        if (descr != want && descr != std::string("|") + (want + 1))
            throw std::runtime_error(path + ": dtype " + descr + ", expected " + want);
    }
};

template <class T> struct npy_descr;
template <> struct npy_descr<float> { static const char* get() { return "<f4"; } };
template <> struct npy_descr<int> { static const char* get() { return "<i4"; } };
template <> struct npy_descr<long> { static const char* get() { return "<i8"; } };
template <> struct npy_descr<unsigned int> { static const char* get() { return "<u4"; } };

}  // namespace host
}  // namespace gala_b200

// ---- DenseMatrix: host rows (features / labels / masks), pinned so the emitted cudaMemcpy runs at PCIe speed ----
template <class I, class N, class V>
class DenseMatrix {
public:
    typedef I itype;
    typedef N ntype;
    typedef V vtype;
    enum DENSE_MTX_TYPE { RM, CM };

    DenseMatrix() = default;
    DenseMatrix(const DenseMatrix&) = delete;
    DenseMatrix& operator=(const DenseMatrix&) = delete;
    ~DenseMatrix() { release(); }

    void build(I nrows, I ncols, DENSE_MTX_TYPE type, int = 0) {
        release();
        nrows_ = nrows;
        ncols_ = ncols;
        type_ = type;
        const size_t bytes = std::max<size_t>((size_t)nrows * (size_t)ncols * sizeof(V), 1);
        if (cudaHostAlloc((void**)&vals_, bytes, cudaHostAllocDefault) == cudaSuccess) {
            pinned_ = true;
        } else {
            cudaGetLastError();
            vals_ = static_cast<V*>(std::malloc(bytes));
            pinned_ = false;
        }
        std::memset(vals_, 0, bytes);      // the reference leaves new buffers uninitialised (dense_matrix.h:128-141)
    }
    I nrows() const { return nrows_; }
    I ncols() const { return ncols_; }
    N nvals() const { return (N)nrows_ * (N)ncols_; }
    V* vals_ptr() { return vals_; }
    DENSE_MTX_TYPE type() const { return type_; }

private:
    void release() {
        if (vals_) pinned_ ? (void)cudaFreeHost(vals_) : std::free(vals_);
        vals_ = nullptr;
    }
    I nrows_ = 0, ncols_ = 0;
    V* vals_ = nullptr;
    bool pinned_ = false;
    DENSE_MTX_TYPE type_ = RM;
};

// ---- CSRCMatrix: CSR on the device ----
template <class I, class N, class V>
class CSRCMatrix {
public:
    typedef I itype;
    typedef N ntype;
    typedef V vtype;
    static_assert(sizeof(I) == 4 && sizeof(N) == 4 && sizeof(V) == 4, "int32 indices, float32 values (common.h:1682-1693)");

    I nrows() const { return nrows_; }
    I ncols() const { return ncols_; }
    N nvals() const { return (N)ids.numel(); }
    N* offset_ptr() { return reinterpret_cast<N*>(offsets.data_ptr<int>()); }     // DEVICE pointers
    I* ids_ptr() { return reinterpret_cast<I*>(ids.data_ptr<int>()); }
    V* vals_ptr() { return reinterpret_cast<V*>(vals.data_ptr<float>()); }
    void set_all(V v) { vals.fill_(v); }
    void import_device(I nrows, I ncols, torch::Tensor off, torch::Tensor id, torch::Tensor va) {
        nrows_ = nrows;
        ncols_ = ncols;
        offsets = std::move(off);
        ids = std::move(id);
        vals = std::move(va);
    }

    torch::Tensor offsets, ids, vals;   // CUDA int32 [nrows+1], int32 [nvals], float32 [nvals]

private:
    I nrows_ = 0, ncols_ = 0;
};

// ---- readers ----
// Adj_src.npy = uint32 [nrows, ncols, src...], Adj_dst.npy = uint32 [dst...] (scripts/Data/gala_export_npy.py).
// The edge lists go from the page cache to the device and the CSR is built there (gala_csr_from_coo).
template <class SM>
void readSM_npy32(std::string path, SM* adj) {
    using namespace gala_b200::host;
    NpyFile src(path + "Adj_src.npy"), dst(path + "Adj_dst.npy");
    src.expect("<u4", path + "Adj_src.npy");
    dst.expect("<u4", path + "Adj_dst.npy");
    const uint32_t* s = reinterpret_cast<const uint32_t*>(src.data);
    const int nrows = (int)s[0], ncols = (int)s[1];
    const int64_t nvals = (int64_t)dst.count;
    if ((int64_t)src.count != nvals + 2) throw std::runtime_error(path + ": Adj_src / Adj_dst length mismatch");
    torch::Tensor rows = torch::empty({nvals}, dev_i32()), cols = torch::empty({nvals}, dev_i32());
    cuda_check(cudaMemcpy(rows.data_ptr<int>(), s + 2, nvals * 4, cudaMemcpyDefault), "upload Adj_src");
    cuda_check(cudaMemcpy(cols.data_ptr<int>(), dst.data, nvals * 4, cudaMemcpyDefault), "upload Adj_dst");
    torch::Tensor off = torch::empty({(int64_t)nrows + 1}, dev_i32()), ids = torch::empty({nvals}, dev_i32());
    const size_t wsb = gala_csr_from_coo_workspace_bytes(nrows, ncols, nvals);
    torch::Tensor ws = torch::empty({(int64_t)std::max<size_t>(wsb, 16)}, dev_u8());
    check(gala_csr_from_coo(nrows, ncols, nvals, rows.data_ptr<int>(), cols.data_ptr<int>(), nullptr, off.data_ptr<int>(),
                            ids.data_ptr<int>(), nullptr, ws.data_ptr(), wsb, stream()),
          "gala_csr_from_coo");
    adj->import_device(nrows, ncols, off, ids, torch::ones({nvals}, dev_f32()));   // set_all(1), tests/common.h:363
}

template <class DM>
void readDM_npy(std::string filename, DM* mtx, typename DM::DENSE_MTX_TYPE type) {
    using namespace gala_b200::host;
    typedef typename DM::vtype V;
    NpyFile f(filename);
    f.expect(npy_descr<V>::get(), filename);
    if (f.shape.size() != 2) throw std::runtime_error(filename + ": expected a 2-D array");
    mtx->build((typename DM::itype)f.shape[0], (typename DM::itype)f.shape[1], type, 0);
    std::memcpy(mtx->vals_ptr(), f.data, f.count * sizeof(V));
}

template <class DM1, class DM2>
void repopulate(DM1* src, DM2* dst) {
    dst->build(src->nrows(), src->ncols(), DM2::RM, 0);
    const int64_t n = (int64_t)src->nrows() * src->ncols();
    for (int64_t i = 0; i < n; ++i) dst->vals_ptr()[i] = (typename DM2::vtype)src->vals_ptr()[i];
}

// ---- column tiling ----
template <class SM>
std::vector<typename SM::itype> static_ord_col_breakpoints(SM* mtx, typename SM::itype cols_per_partition) {
    typedef typename SM::itype iT;
    std::vector<iT> res;
    res.push_back(0);
    for (iT i = 0; i < mtx->ncols(); i += cols_per_partition) res.push_back(std::min(mtx->ncols(), i + cols_per_partition));
    return res;
}

template <class SM>
void ord_col_tiling_torch(std::vector<typename SM::itype>& col_breakpoints, torch::Tensor& output_offsets,
                          torch::Tensor& output_cols, torch::Tensor& output_vals, torch::Tensor& output_bounds, SM* src) {
    using namespace gala_b200::host;
    const int segments = (int)col_breakpoints.size() - 1;
    // uniform breakpoints as static_ord_col_breakpoints produces them: the partition width is the first one
    const int T = segments > 1 ? (int)(col_breakpoints[1] - col_breakpoints[0]) : std::max<int>(src->ncols(), 1);
    if (gala_col_tile_segments(src->ncols(), T) != segments) throw std::runtime_error("ord_col_tiling_torch: non-uniform column breakpoints");
    const int64_t nvals = src->nvals();
    output_offsets = torch::empty({(int64_t)(src->nrows() + 1) * segments}, dev_i32());
    output_cols = torch::empty({nvals}, dev_i32());
    output_vals = torch::empty({nvals}, dev_f32());
    output_bounds = torch::zeros({2 * (int64_t)segments}, torch::TensorOptions().dtype(torch::kInt));   // CPU, as the emitted wrappers read it
    const size_t wsb = gala_col_tile_workspace_bytes(src->nrows(), src->ncols(), T);
    torch::Tensor ws = torch::empty({(int64_t)std::max<size_t>(wsb, 16)}, dev_u8());
    check(gala_col_tile(src->nrows(), src->ncols(), nvals, src->offsets.template data_ptr<int>(), src->ids.template data_ptr<int>(),
                        src->vals.template data_ptr<float>(), T, output_offsets.data_ptr<int>(), output_cols.data_ptr<int>(),
                        output_vals.data_ptr<float>(), output_bounds.data_ptr<int>(), ws.data_ptr(), wsb, stream()),
          "gala_col_tile");
}

// ---- sampling ----
template <class SM>
void inplace_sample_graph_ab(SM* src, int sample_size, int ra, int rb) {
    using namespace gala_b200::host;
    const int n = src->nrows();
    torch::Tensor no = torch::empty({(int64_t)n + 1}, dev_i32()), ni = torch::empty({(int64_t)n * sample_size}, dev_i32());
    torch::Tensor nv = torch::empty({(int64_t)n * sample_size}, dev_f32()), status = torch::zeros({1}, dev_i32());
    check(gala_sample_ab(n, src->offsets.template data_ptr<int>(), src->ids.template data_ptr<int>(),
                         src->vals.template data_ptr<float>(), sample_size, ra, rb, no.data_ptr<int>(), ni.data_ptr<int>(),
                         nv.data_ptr<float>(), status.data_ptr<int>(), stream()),
          "gala_sample_ab");
    if (status.item<int>() != 0) throw std::runtime_error("inplace_sample_graph_ab: a row has no edge (`% 0` in the reference, tiling.h:480)");
    src->import_device(n, src->ncols(), no, ni, nv);
}

// ---- training sub-graphs ----
template <class SM, class DM>
void getMaskSubgraphs(SM* adj, DM* mask, int layers, std::vector<SM*>& forward_vec, std::vector<SM*>& backward_vec) {
    using namespace gala_b200::host;
    const int n = adj->nrows(), nc = adj->ncols();
    const int64_t E = adj->nvals();
    torch::Tensor cur = torch::empty({(int64_t)n}, dev_u8());
    static_assert(sizeof(typename DM::vtype) == 1, "bool masks (repopulate<DBL, DB>)");
    cuda_check(cudaMemcpy(cur.data_ptr(), mask->vals_ptr(), n, cudaMemcpyDefault), "upload mask");
    const size_t wsb = gala_mask_subgraph_workspace_bytes(n);
    torch::Tensor ws = torch::empty({(int64_t)std::max<size_t>(wsb, 16)}, dev_u8());
    const size_t twsb = gala_csr_from_coo_workspace_bytes(nc, n, E);
    torch::Tensor tws = torch::empty({(int64_t)std::max<size_t>(twsb, 16)}, dev_u8());
    for (int l = 0; l < layers; ++l) {
        torch::Tensor no = torch::empty({(int64_t)n + 1}, dev_i32()), ni = torch::empty({std::max<int64_t>(E, 1)}, dev_i32());
        torch::Tensor nv = torch::empty({std::max<int64_t>(E, 1)}, dev_f32()), nxt = torch::empty({(int64_t)n}, dev_u8());
        int64_t total = 0;
        check(gala_mask_subgraph(n, adj->offsets.template data_ptr<int>(), adj->ids.template data_ptr<int>(),
                                 adj->vals.template data_ptr<float>(), cur.template data_ptr<uint8_t>(), no.data_ptr<int>(),
                                 ni.data_ptr<int>(), nv.data_ptr<float>(), &total, nxt.template data_ptr<uint8_t>(), ws.data_ptr(), wsb,
                                 stream()),
              "gala_mask_subgraph");
        ni = ni.narrow(0, 0, total).clone();
        nv = nv.narrow(0, 0, total).clone();
        SM* f = new SM();
        f->import_device(n, nc, no, ni, nv);
        forward_vec.push_back(f);
        torch::Tensor to = torch::empty({(int64_t)nc + 1}, dev_i32()), ti = torch::empty({std::max<int64_t>(total, 1)}, dev_i32());
        torch::Tensor tv = torch::empty({std::max<int64_t>(total, 1)}, dev_f32());
        check(gala_csr_transpose(n, nc, total, no.data_ptr<int>(), ni.data_ptr<int>(), nv.data_ptr<float>(), to.data_ptr<int>(),
                                 ti.data_ptr<int>(), tv.data_ptr<float>(), tws.data_ptr(), twsb, stream()),
              "gala_csr_transpose");
        SM* b = new SM();
        b->import_device(nc, n, to, ti.narrow(0, 0, total), tv.narrow(0, 0, total));
        backward_vec.push_back(b);
        cur = nxt;
    }
}

inline double get_time() {
    return std::chrono::duration<double>(std::chrono::steady_clock::now().time_since_epoch()).count();
}

inline double calc_mean(std::vector<double>& vec) {
    if (vec.empty()) return 0.0;
    double mean = 0;
    for (double v : vec) mean += v;
    return mean / (double)vec.size();
}
