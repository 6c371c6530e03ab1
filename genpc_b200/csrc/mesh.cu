// mesh.cu -- area-weighted surface sampling of a triangle mesh on the GPU.
//
// Replaces the sampling step of the reference's glb2point (utils/dataUtils.py:217-250: trimesh `mesh.sample(num_points,
// return_index=True)` + barycentric colour interpolation :231-243), which feeds reg() with 163 840 samples of the
// generated shape (reg_xyz.py:125) and object_pose_optimization with 120 000 (diff_obj_pose.py:504).  trimesh is not
// vendored and its sampler is unseeded; the semantics are DEFINED here so that the oracle (oracle/mesh.py) and the kernels
// agree bit for bit:
//   area_f   = 0.5 * |e1 x e2|, e1 = v1 - v0, e2 = v2 - v0, every operation individually rounded in fp32 (no fma);
//   weight_f = floor(area_f / max_area * 2^32) as uint64 (division and product in double: correctly rounded, exact);
//              inclusive prefix sums in uint64 (exact, order independent) -- done by the caller;
//   sample i : four 32-bit draws from splitmix64(seed, i); face = first f with cum[f] > mulhi64(draw01, total);
//              (u, v) = two 24-bit uniforms, reflected into the triangle when u + v > 1;
//              p = v0 + (e1*u + e2*v), colour = (1-u-v) c0 + u c1 + v c2 (same rounding rules).
#include "common.cuh"

namespace genpc {

__device__ __forceinline__ unsigned long long splitmix64(unsigned long long x) {
    x += 0x9E3779B97F4A7C15ull;
    x = (x ^ (x >> 30)) * 0xBF58476D1CE4E5B9ull;
    x = (x ^ (x >> 27)) * 0x94D049BB133111EBull;
    return x ^ (x >> 31);
}

__global__ void mesh_face_area_kernel(const float *__restrict__ verts, const int *__restrict__ faces, int F, int V,
                                      float *__restrict__ area) {
    const int f = blockIdx.x * blockDim.x + threadIdx.x;
    if (f >= F) return;
    const int i0 = faces[f * 3], i1 = faces[f * 3 + 1], i2 = faces[f * 3 + 2];
    if ((unsigned)i0 >= (unsigned)V || (unsigned)i1 >= (unsigned)V || (unsigned)i2 >= (unsigned)V) {
        area[f] = 0.f;  // a face with an out-of-range vertex is never sampled
        return;
    }
    const float *a = verts + (size_t)i0 * 3, *b = verts + (size_t)i1 * 3, *c = verts + (size_t)i2 * 3;
    const float e1x = __fsub_rn(b[0], a[0]), e1y = __fsub_rn(b[1], a[1]), e1z = __fsub_rn(b[2], a[2]);
    const float e2x = __fsub_rn(c[0], a[0]), e2y = __fsub_rn(c[1], a[1]), e2z = __fsub_rn(c[2], a[2]);
    const float cx = __fsub_rn(__fmul_rn(e1y, e2z), __fmul_rn(e1z, e2y));
    const float cy = __fsub_rn(__fmul_rn(e1z, e2x), __fmul_rn(e1x, e2z));
    const float cz = __fsub_rn(__fmul_rn(e1x, e2y), __fmul_rn(e1y, e2x));
    const float n2 = __fadd_rn(__fadd_rn(__fmul_rn(cx, cx), __fmul_rn(cy, cy)), __fmul_rn(cz, cz));
    const float ar = __fmul_rn(0.5f, __fsqrt_rn(n2));
    area[f] = (ar == ar && ar < __int_as_float(0x7f800000)) ? ar : 0.f;  // NaN / inf areas carry no weight
}

__global__ void mesh_sample_kernel(const float *__restrict__ verts, const int *__restrict__ faces,
                                   const float *__restrict__ vcol, const unsigned long long *__restrict__ cum, int F,
                                   int n, unsigned long long seed, float *__restrict__ out_xyz, float *__restrict__ out_rgb,
                                   int *__restrict__ out_face) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    const unsigned long long r0 = splitmix64(seed ^ (0xD1B54A32D192ED03ull * (unsigned long long)(2 * i + 1)));
    const unsigned long long r1 = splitmix64(r0);
    const unsigned long long total = cum[F - 1];
    const unsigned long long target = __umul64hi(r0, total);  // uniform in [0, total)
    int lo = 0, hi = F - 1;                                    // first face with cum[f] > target
    while (lo < hi) {
        const int mid = (lo + hi) >> 1;
        if (__ldg(cum + mid) > target) hi = mid;
        else lo = mid + 1;
    }
    const int f = lo;
    float u = __fmul_rn((float)(unsigned)(r1 >> 40), 5.9604644775390625e-08f);           // 24 bits * 2^-24
    float v = __fmul_rn((float)(unsigned)((r1 >> 16) & 0xffffffu), 5.9604644775390625e-08f);
    if (__fadd_rn(u, v) > 1.f) u = __fsub_rn(1.f, u), v = __fsub_rn(1.f, v);
    const int i0 = faces[f * 3], i1 = faces[f * 3 + 1], i2 = faces[f * 3 + 2];
    const float *a = verts + (size_t)i0 * 3, *b = verts + (size_t)i1 * 3, *c = verts + (size_t)i2 * 3;
#pragma unroll
    for (int k = 0; k < 3; ++k) {
        const float e1 = __fsub_rn(b[k], a[k]), e2 = __fsub_rn(c[k], a[k]);
        out_xyz[(size_t)i * 3 + k] = __fadd_rn(a[k], __fadd_rn(__fmul_rn(e1, u), __fmul_rn(e2, v)));
    }
    if (out_rgb != nullptr) {
        const float w0 = __fsub_rn(__fsub_rn(1.f, u), v);
#pragma unroll
        for (int k = 0; k < 3; ++k) {
            float col = 0.5f;  // the reference's default for meshes without colour (utils/dataUtils.py:234-236)
            if (vcol != nullptr)
                col = __fadd_rn(__fadd_rn(__fmul_rn(w0, vcol[(size_t)i0 * 3 + k]), __fmul_rn(u, vcol[(size_t)i1 * 3 + k])),
                                __fmul_rn(v, vcol[(size_t)i2 * 3 + k]));
            out_rgb[(size_t)i * 3 + k] = fminf(fmaxf(col, 0.f), 1.f);                    // np.clip(color, 0, 1) :243
        }
    }
    if (out_face != nullptr) out_face[i] = f;
}

}  // namespace genpc

using namespace genpc;

extern "C" int genpc_mesh_face_areas(const float *verts, const int *faces, int n_verts, int n_faces, float *areas,
                                     genpc_stream_t stream_) {
    if (n_verts < 0 || n_faces < 0) return GENPC_ERR_SHAPE;
    if (n_faces == 0) return GENPC_OK;
    mesh_face_area_kernel<<<(n_faces + 255) / 256, 256, 0, (cudaStream_t)stream_>>>(verts, faces, n_faces, n_verts, areas);
    GENPC_CHECK_LAUNCH();
    return GENPC_OK;
}

extern "C" int genpc_mesh_sample(const float *verts, const int *faces, const float *vertex_rgb,
                                 const unsigned long long *cum_weights, int n_faces, int n_samples, unsigned long long seed,
                                 float *out_xyz, float *out_rgb, int *out_face, genpc_stream_t stream_) {
    if (n_faces <= 0 || n_samples < 0) return GENPC_ERR_SHAPE;
    if (n_samples == 0) return GENPC_OK;
    mesh_sample_kernel<<<(n_samples + 255) / 256, 256, 0, (cudaStream_t)stream_>>>(verts, faces, vertex_rgb, cum_weights, n_faces,
                                                                                   n_samples, seed, out_xyz, out_rgb, out_face);
    GENPC_CHECK_LAUNCH();
    return GENPC_OK;
}
