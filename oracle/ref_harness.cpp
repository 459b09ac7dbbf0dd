// ref_harness.cpp -- extern "C" wrappers around the UNMODIFIED reference headers,
// compiled from the sources where they lie under /root/reference (never copied).
//
// TEST INFRASTRUCTURE ONLY (see oracle/gala_oracle.c header).  Built by
// oracle/Makefile into oracle/_ref/libgala_ref.so; used to pin the C restatement
// in oracle/gala_oracle.c, to generate tests/golden/*.npz, and as the
// `--impl reference` / cpu_baseline arm of bench.py.
//
// Wrapped reference entry points (file:line relative to /root/reference):
//   CSRCMatrix::build            src/formats/csrc_matrix.h:148-376
//   buildTranspose               tests/common.h:107-123
//   gSpMM<wsumAgg>, gSpMM<maxAgg> src/ops/aggregators.h:55-127
//   static_ord_col_breakpoints   src/ops/tiling.h:1594-1608
//   ord_col_tiling_torch         src/ops/tiling.h:222-283
//   inplace_sample_graph_ab      src/ops/tiling.h:454-508
//   getMaskSubgraphs             tests/common.h:20-105
//
// Build flags mirror the generated CMake (src/codegen/cuda.h:46-51):
//   -DGALA_TORCH -DGN_1 -DPT_0 -DST_0 -DA_ALLOC, g++ -O3 -march=native -fopenmp.
#include <torch/torch.h>

#include <cstdint>
#include <cstdlib>
#include <cstring>
#include <vector>

#include "src/formats/csrc_matrix.h"
#include "src/formats/dense_matrix.h"
#include "src/ops/aggregators.h"
#include "src/ops/tiling.h"
#include "src/ops/reordering.h"
#include "tests/common.h"

typedef int ind1_t;
typedef int ind2_t;
typedef float val_t;
typedef DenseMatrix<ind1_t, ind2_t, val_t> DM;
typedef DenseMatrix<ind1_t, ind2_t, bool> DB;
typedef CSRCMatrix<ind1_t, ind2_t, val_t> SM;

namespace {

template <class T>
T *dup(const T *src, size_t n) {
    T *p = (T *)aligned_alloc(64, ((n ? n : 1) * sizeof(T) + 63) / 64 * 64);
    if (src) std::memcpy(p, src, n * sizeof(T));
    return p;
}

// A CSR view over caller memory; never frees (CSRCMatrix::~CSRCMatrix would
// deallocate through std::allocator, so the object is leaked on purpose).
SM *view_csr(int nrows, int ncols, const int *offset, const int *ids, const float *vals) {
    SM *m = new SM();
    m->import_csr(nrows, ncols, offset[nrows], const_cast<int *>(ids), const_cast<float *>(vals),
                  const_cast<int *>(offset));
    return m;
}

}  // namespace

extern "C" {

int ref_num_threads() { return omp_get_max_threads(); }

// COO -> CSR through the reference's build().  Inputs are copied first because
// build() uses the caller's col_ids/vals arrays as sort scratch
// (csrc_matrix.h:249-260).
int ref_csr_build(int nrows, int ncols, int64_t nvals, const int *row_ids, const int *col_ids,
                  const float *vals, int *offset, int *ids, float *out_vals) {
    int *r = dup(row_ids, (size_t)nvals);
    int *c = dup(col_ids, (size_t)nvals);
    float *v = dup(vals, (size_t)nvals);
    SM *m = new SM();
    auto info = m->build(nrows, ncols, (int)nvals, r, c, v, CSRC_TYPE::CSR);
    if (info != INFO::SUCCESS) return 1;
    std::memcpy(offset, m->offset_ptr(), ((size_t)nrows + 1) * sizeof(int));
    std::memcpy(ids, m->ids_ptr(), (size_t)nvals * sizeof(int));
    std::memcpy(out_vals, m->vals_ptr(), (size_t)nvals * sizeof(float));
    delete m;
    free(r);
    free(c);
    free(v);
    return 0;
}

int ref_csr_transpose(int nrows, int ncols, const int *offset, const int *ids, const float *vals,
                      int *t_offset, int *t_ids, float *t_vals) {
    int64_t nvals = offset[nrows];
    // buildTranspose hands the source's own ids/vals to build(), which scribbles
    // on them: work on copies.
    int *o = dup(offset, (size_t)nrows + 1);
    int *c = dup(ids, (size_t)nvals);
    float *v = dup(vals, (size_t)nvals);
    SM *src = view_csr(nrows, ncols, o, c, v);
    SM *res = new SM();
    buildTranspose(src, res);
    std::memcpy(t_offset, res->offset_ptr(), ((size_t)ncols + 1) * sizeof(int));
    std::memcpy(t_ids, res->ids_ptr(), (size_t)nvals * sizeof(int));
    std::memcpy(t_vals, res->vals_ptr(), (size_t)nvals * sizeof(float));
    delete res;
    free(o);
    free(c);
    free(v);
    return 0;
}

// out += A * B with wsumAgg; out is caller-zeroed.
void ref_gspmm_wsum(int nrows, int ncols, const int *offset, const int *ids, const float *vals,
                    const float *B, int K, float *out) {
    SM *A = view_csr(nrows, ncols, offset, ids, vals);
    DM Bm, Om;
    Bm.import_mtx(ncols, K, (int)((int64_t)ncols * K), const_cast<float *>(B));
    Om.import_mtx(nrows, K, (int)((int64_t)nrows * K), out);
    auto agg = wsumAgg<val_t, val_t, ind2_t>;
    gSpMM(A, &Bm, &Om, agg);
    Bm.import_mtx((float *)nullptr);
    Om.import_mtx((float *)nullptr);
}

int ref_col_breakpoints(int nrows, int ncols, const int *offset, const int *ids,
                        const float *vals, int cols_per_partition, int *out, int cap) {
    SM *A = view_csr(nrows, ncols, offset, ids, vals);
    std::vector<int> bp = static_ord_col_breakpoints<SM>(A, cols_per_partition);
    if ((int)bp.size() > cap) return -1;
    std::memcpy(out, bp.data(), bp.size() * sizeof(int));
    return (int)bp.size();
}

void ref_col_tile(int nrows, int ncols, const int *offset, const int *ids, const float *vals,
                  int nbreak, const int *breakpoints, int *out_offsets, int *out_cols,
                  float *out_vals, int *out_bounds) {
    SM *A = view_csr(nrows, ncols, offset, ids, vals);
    std::vector<int> bp(breakpoints, breakpoints + nbreak);
    int S = nbreak - 1;
    int64_t nvals = offset[nrows];
    auto oi = torch::TensorOptions().dtype(torch::kInt).requires_grad(false);
    auto of = torch::TensorOptions().dtype(torch::kFloat).requires_grad(false);
    torch::Tensor t_off = torch::zeros({(int64_t)(nrows + 1) * S}, oi);
    torch::Tensor t_col = torch::zeros({nvals}, oi);
    torch::Tensor t_val = torch::zeros({nvals}, of);
    torch::Tensor t_bnd = torch::zeros({2 * S}, oi);
    ord_col_tiling_torch(bp, t_off, t_col, t_val, t_bnd, A);
    std::memcpy(out_offsets, t_off.data_ptr<int>(), (size_t)(nrows + 1) * S * sizeof(int));
    std::memcpy(out_cols, t_col.data_ptr<int>(), (size_t)nvals * sizeof(int));
    std::memcpy(out_vals, t_val.data_ptr<float>(), (size_t)nvals * sizeof(float));
    std::memcpy(out_bounds, t_bnd.data_ptr<int>(), (size_t)2 * S * sizeof(int));
}

void ref_sample_ab(int nrows, int ncols, const int *offset, const int *ids, const float *vals,
                   int sample_size, int ra, int rb, int *new_offset, int *new_ids,
                   float *new_vals) {
    SM *A = view_csr(nrows, ncols, offset, ids, vals);
    inplace_sample_graph_ab(A, sample_size, ra, rb);
    int64_t nv = (int64_t)nrows * sample_size;
    std::memcpy(new_offset, A->offset_ptr(), ((size_t)nrows + 1) * sizeof(int));
    std::memcpy(new_ids, A->ids_ptr(), (size_t)nv * sizeof(int));
    std::memcpy(new_vals, A->vals_ptr(), (size_t)nv * sizeof(float));
}

// getMaskSubgraphs for `layers` layers.  The reference reads an uninitialised
// next-mask buffer (tests/common.h:99-102); glibc hands back zero pages for a
// fresh large allocation but not in general, so the harness can only be
// trusted when `layers == 1` (no propagated mask is consumed) or when the
// caller checks against the zero-initialised restatement.  Outputs: for layer l
// forward CSR (offset at fwd_offsets + l*(nrows+1), ids/vals concatenated with
// per-layer nnz in fwd_nvals[l]) and the same for the transposes.
int ref_mask_subgraphs(int nrows, int ncols, const int *offset, const int *ids, const float *vals,
                       const uint8_t *mask, int layers, int *fwd_offsets, int *fwd_ids,
                       float *fwd_vals, int *fwd_nvals, int *bwd_offsets, int *bwd_ids,
                       float *bwd_vals) {
    SM *A = view_csr(nrows, ncols, offset, ids, vals);
    DB m;
    m.build(nrows, 1, DB::DENSE_MTX_TYPE::RM, 0);
    for (int i = 0; i < nrows; i++) m.vals_ptr()[i] = mask[i] != 0;
    std::vector<SM *> fwd, bwd;
    getMaskSubgraphs(A, &m, layers, fwd, bwd);
    int64_t pos = 0;
    for (int l = 0; l < layers; l++) {
        int nv = fwd[l]->nvals();
        fwd_nvals[l] = nv;
        std::memcpy(fwd_offsets + (size_t)l * (nrows + 1), fwd[l]->offset_ptr(),
                    ((size_t)nrows + 1) * sizeof(int));
        std::memcpy(fwd_ids + pos, fwd[l]->ids_ptr(), (size_t)nv * sizeof(int));
        std::memcpy(fwd_vals + pos, fwd[l]->vals_ptr(), (size_t)nv * sizeof(float));
        std::memcpy(bwd_offsets + (size_t)l * (ncols + 1), bwd[l]->offset_ptr(),
                    ((size_t)ncols + 1) * sizeof(int));
        std::memcpy(bwd_ids + pos, bwd[l]->ids_ptr(), (size_t)nv * sizeof(int));
        std::memcpy(bwd_vals + pos, bwd[l]->vals_ptr(), (size_t)nv * sizeof(float));
        pos += nv;
    }
    return 0;
}

// rowReorderToAdj (src/ops/reordering.h:940-1013).  It free()s the matrix' previous arrays, so it
// is handed malloc'd copies.
void ref_row_reorder_to_adj(int nrows, const int *offset, const int *ids, const float *vals,
                            const int *perm, int *new_offset, int *new_ids, float *new_vals) {
    int64_t nvals = offset[nrows];
    int *o = dup(offset, (size_t)nrows + 1);
    int *c = dup(ids, (size_t)nvals);
    float *v = dup(vals, (size_t)nvals);
    SM *A = view_csr(nrows, nrows, o, c, v);
    rowReorderToAdj(A, const_cast<int *>(perm));
    std::memcpy(new_offset, A->offset_ptr(), ((size_t)nrows + 1) * sizeof(int));
    std::memcpy(new_ids, A->ids_ptr(), (size_t)nvals * sizeof(int));
    std::memcpy(new_vals, A->vals_ptr(), (size_t)nvals * sizeof(float));
}

// rowPermuteDenseTo (:244-283, PR_0 flavour) / rowPermuteDenseFrom (:207-236); both work in place.
void ref_row_permute_dense(int nrows, int K, float *X, const int *perm, int from) {
    DM Xm;
    Xm.import_mtx(nrows, K, (int)((int64_t)nrows * K), X);
    DM *p = &Xm;
    if (from)
        rowPermuteDenseFrom(p, const_cast<int *>(perm));
    else
        rowPermuteDenseTo(p, const_cast<int *>(perm));
    Xm.import_mtx((float *)nullptr);
}

}  // extern "C"
