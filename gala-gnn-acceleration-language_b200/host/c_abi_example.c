/* c_abi_example.c -- the drop-in boundary used from plain C: no torch, no C++, only
 * include/gala_b200.h + the CUDA runtime.  COO -> CSR on the device (gala_csr_from_coo),
 * column tiling (gala_col_tile), plan (gala_plan_build), aggregation (gala_spmm_f32) and the fused
 * GAT layer (gala_gat_forward_f32); results are compared with a host loop.  Built by
 * host/Makefile, run by tests/test_shim_gpu.py.  This is what a binding from any other host
 * language (cgo / JNI / ctypes / N-API) does, spelled out. */
#include <cuda_runtime_api.h>
#include <math.h>
#include <stdio.h>
#include <stdlib.h>

#include "gala_b200.h"

#define CK(x)                                                                         \
    do {                                                                              \
        int rc_ = (int)(x);                                                           \
        if (rc_ != 0) {                                                               \
            fprintf(stderr, "%s -> %d (%s)\n", #x, rc_, gala_b200_error_string(rc_)); \
            return 1;                                                                 \
        }                                                                             \
    } while (0)

static void *dmalloc(size_t n) {
    void *p = NULL;
    if (cudaMalloc(&p, n ? n : 16) != cudaSuccess) exit(2);
    return p;
}

int main(void) {
    const int N = 1500, K = 32, T = 400; /* 4 column segments */
    long long E = 0; /* row i has 20 + i % 40 edges: with hub_threshold 48 some rows run on whole CTAs */
    for (int i = 0; i < N; i++) E += 20 + i % 40;
    int *h_r = malloc(E * sizeof(int)), *h_c = malloc(E * sizeof(int));
    float *h_v = malloc(E * sizeof(float)), *h_x = malloc((size_t)N * K * sizeof(float));
    unsigned s = 12345u;
    long long e = 0;
    for (int i = 0; i < N; i++)
        for (int j = 0; j < 20 + i % 40; j++, e++) {
            s = s * 1664525u + 1013904223u;
            h_r[e] = i;
            h_c[e] = (int)((s >> 8) % (unsigned)N);
            h_v[e] = (float)((s >> 4) & 1023) / 1024.0f;
        }
    for (long long i = 0; i < (long long)N * K; i++) {
        s = s * 1664525u + 1013904223u;
        h_x[i] = (float)(s >> 9) / 8388608.0f - 0.5f;
    }
    /* host reference: Y = A X (any edge order: fp32 sums compared with a tolerance) */
    double *want = calloc((size_t)N * K, sizeof(double));
    for (e = 0; e < E; e++)
        for (int k = 0; k < K; k++) want[(size_t)h_r[e] * K + k] += (double)h_v[e] * h_x[(size_t)h_c[e] * K + k];

    int *d_r = dmalloc(E * 4), *d_c = dmalloc(E * 4), *d_off = dmalloc((N + 1) * 4), *d_ids = dmalloc(E * 4);
    float *d_v = dmalloc(E * 4), *d_vs = dmalloc(E * 4), *d_x = dmalloc((size_t)N * K * 4), *d_y = dmalloc((size_t)N * K * 4);
    cudaMemcpy(d_r, h_r, E * 4, cudaMemcpyHostToDevice);
    cudaMemcpy(d_c, h_c, E * 4, cudaMemcpyHostToDevice);
    cudaMemcpy(d_v, h_v, E * 4, cudaMemcpyHostToDevice);
    cudaMemcpy(d_x, h_x, (size_t)N * K * 4, cudaMemcpyHostToDevice);

    /* CSRCMatrix::build on the device */
    size_t ws_b = gala_csr_from_coo_workspace_bytes(N, N, E);
    void *ws = dmalloc(ws_b);
    CK(gala_csr_from_coo(N, N, E, d_r, d_c, d_v, d_off, d_ids, d_vs, ws, ws_b, NULL));

    /* static_ord_col_breakpoints + ord_col_tiling_torch on the device; bounds come back on the host */
    int S = gala_col_tile_segments(N, T);
    int *d_toff = dmalloc((size_t)S * (N + 1) * 4), *d_tcol = dmalloc(E * 4);
    float *d_tval = dmalloc(E * 4);
    int *bounds = malloc(2 * S * sizeof(int));
    size_t tw_b = gala_col_tile_workspace_bytes(N, N, T);
    void *tws = dmalloc(tw_b);
    CK(gala_col_tile(N, N, E, d_off, d_ids, d_vs, T, d_toff, d_tcol, d_tval, bounds, tws, tw_b, NULL));

    gala_graph_t g;
    g.offsets = d_toff;
    g.cols = d_tcol;
    g.bounds = bounds;
    g.nrows = N;
    g.ncols = N;
    g.segments = S;
    g.nvals = E;
    g.bounds_dev = NULL;   /* only needed beyond 64 column segments */
    gala_plan_t plan;
    size_t pw_b = gala_plan_workspace_bytes(&g);
    void *pws = dmalloc(pw_b);
    CK(gala_plan_build(&g, 48, pws, pw_b, &plan, NULL));

    CK(gala_spmm_f32(&g, d_tval, d_x, K, d_y, NULL, &plan, NULL));
    float *h_y = malloc((size_t)N * K * sizeof(float));
    CK(cudaMemcpy(h_y, d_y, (size_t)N * K * 4, cudaMemcpyDeviceToHost));
    double num = 0, den = 0;
    for (long long i = 0; i < (long long)N * K; i++) {
        num += (h_y[i] - want[i]) * (h_y[i] - want[i]);
        den += want[i] * want[i];
    }
    double err = sqrt(num / den);
    printf("gala_spmm_f32 over %d segments: rel err %.3e\n", S, err);
    if (!(err < 1e-5)) return 1;

    /* fused GAT layer: attention rows are a convex combination -> with X = 1 every output is 1 */
    float *h_a = malloc(N * sizeof(float));
    for (int i = 0; i < N; i++) h_a[i] = (float)(i % 17) / 17.0f - 0.5f;
    float *d_a = dmalloc(N * 4);
    cudaMemcpy(d_a, h_a, N * 4, cudaMemcpyHostToDevice);
    for (long long i = 0; i < (long long)N * K; i++) h_x[i] = 1.0f;
    cudaMemcpy(d_x, h_x, (size_t)N * K * 4, cudaMemcpyHostToDevice);
    CK(gala_gat_forward_f32(&g, d_a, d_a, d_x, K, 0.2f, d_y, NULL, 0, &plan, NULL));
    CK(cudaMemcpy(h_y, d_y, (size_t)N * K * 4, cudaMemcpyDeviceToHost));
    double worst = 0;
    for (long long i = 0; i < (long long)N * K; i++) worst = fmax(worst, fabs(h_y[i] - 1.0));
    printf("gala_gat_forward_f32: max |y - 1| = %.3e\n", worst);
    if (!(worst < 1e-5)) return 1;

    /* errors are codes, never exits */
    if (gala_spmm_f32(NULL, NULL, NULL, K, NULL, NULL, NULL, NULL) != GALA_ERR_NULL_POINTER) return 1;
    printf("C ABI EXAMPLE OK (abi %d)\n", gala_b200_abi_version());
    return 0;
}
