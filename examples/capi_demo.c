/* capi_demo.c -- libgenpc_b200 used from plain C: no Python, no torch, only include/genpc_b200.h and the CUDA runtime.
 *
 *   gcc -O2 -I include -I /usr/local/cuda/include examples/capi_demo.c -o examples/capi_demo \
 *       -L genpc_b200 -lgenpc_b200 -L /usr/local/cuda/lib64 -lcudart -Wl,-rpath,$PWD/genpc_b200
 *   examples/capi_demo in1.f32 N in2.f32 M out_prefix      (clouds as raw little-endian float32 [N][3] / [M][3], B = 1)
 *
 * Writes <out_prefix>.dist1/.dist2 (float32) and .idx1/.idx2 (int32): the outputs of the reference's
 * chamfer_3D.forward (chamfer_cuda.cpp:17-19), plus the FPS indices of cloud 1 (<out_prefix>.fps, 64 picks).
 */
#include <cuda_runtime.h>
#include <stdio.h>
#include <stdlib.h>

#include "genpc_b200.h"

static void *slurp(const char *path, size_t bytes) {
    FILE *f = fopen(path, "rb");
    void *p = malloc(bytes);
    if (!f || fread(p, 1, bytes, f) != bytes) { fprintf(stderr, "cannot read %s\n", path); exit(2); }
    fclose(f);
    return p;
}
static void dump(const char *prefix, const char *ext, const void *dev, size_t bytes) {
    char path[1024];
    void *h = malloc(bytes);
    cudaMemcpy(h, dev, bytes, cudaMemcpyDeviceToHost);
    snprintf(path, sizeof path, "%s.%s", prefix, ext);
    FILE *f = fopen(path, "wb");
    fwrite(h, 1, bytes, f);
    fclose(f);
    free(h);
}
#define CK(x) do { int rc_ = (x); if (rc_ != 0) { fprintf(stderr, "%s failed: %d\n", #x, rc_); return 1; } } while (0)

int main(int argc, char **argv) {
    if (argc != 6) { fprintf(stderr, "usage: %s in1.f32 N in2.f32 M out_prefix\n", argv[0]); return 2; }
    const int N = atoi(argv[2]), M = atoi(argv[4]), B = 1, K = N < 64 ? N : 64;
    float *h1 = slurp(argv[1], (size_t)N * 12), *h2 = slurp(argv[3], (size_t)M * 12);
    float *x1, *x2, *d1, *d2;
    int *i1, *i2, *fps;
    void *ws;
    cudaStream_t stream;
    CK(cudaStreamCreate(&stream));
    CK(cudaMalloc((void **)&x1, (size_t)N * 12)); CK(cudaMalloc((void **)&x2, (size_t)M * 12));
    CK(cudaMalloc((void **)&d1, (size_t)N * 4)); CK(cudaMalloc((void **)&d2, (size_t)M * 4));
    CK(cudaMalloc((void **)&i1, (size_t)N * 4)); CK(cudaMalloc((void **)&i2, (size_t)M * 4));
    CK(cudaMalloc((void **)&fps, (size_t)K * 4));
    size_t wsb = genpc_chamfer_workspace_bytes(B, N, M), fwb = genpc_fps_workspace_bytes(B, N, K);
    CK(cudaMalloc(&ws, (wsb > fwb ? wsb : fwb) + 16));
    CK(cudaMemcpyAsync(x1, h1, (size_t)N * 12, cudaMemcpyHostToDevice, stream));
    CK(cudaMemcpyAsync(x2, h2, (size_t)M * 12, cudaMemcpyHostToDevice, stream));
    CK(genpc_chamfer_forward(x1, x2, d1, d2, i1, i2, B, N, M, ws, wsb, stream));
    CK(genpc_fps(x1, B, N, K, 0, fps, NULL, ws, fwb, stream));
    CK(cudaStreamSynchronize(stream));
    dump(argv[5], "dist1", d1, (size_t)N * 4); dump(argv[5], "dist2", d2, (size_t)M * 4);
    dump(argv[5], "idx1", i1, (size_t)N * 4); dump(argv[5], "idx2", i2, (size_t)M * 4);
    dump(argv[5], "fps", fps, (size_t)K * 4);
    printf("%s: chamfer %d x %d and %d FPS picks done\n", genpc_version(), N, M, K);
    return 0;
}
