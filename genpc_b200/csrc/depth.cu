// depth.cu -- DepthPrompting geometry for sm_100a: point -> uv projection, z-buffer render, depth -> point.
//
// Replaces:
//   getUvs            DepthPrompting.py:239-271  python loop over cameras + [V,N,3] temporaries + 2 reductions
//   pixel mapping     DepthPrompting.py:179-184
//   paintPixels /     DepthPrompting.py:292-391  index_put scatter, last writer wins, no depth test
//   getRawDepth
// by three HBM/atomic-bound passes.  The z-buffer is a packed 64-bit atomicMin
// ((ordered_key(ndc_z) << 32) | point index): nearest point wins, ties go to the lowest index, so the image
// is deterministic (the reference's duplicate-index index_put is not).  Unprojection (no reference
// counterpart) inverts the mapping at pixel centres with a deterministic raster-order compaction.
// Arithmetic is spelled with explicit rounding and matches oracle/genpc_oracle.c bit for bit.
#include "common.cuh"

namespace genpc {

constexpr int DP_THREADS = 256;

__device__ __forceinline__ unsigned f2key(float f) {
    unsigned u = __float_as_uint(f);
    return (u & 0x80000000u) ? ~u : (u | 0x80000000u);
}
__device__ __forceinline__ float key2f(unsigned k) {
    return __uint_as_float((k & 0x80000000u) ? (k & 0x7fffffffu) : ~k);
}

__device__ __forceinline__ void project_point(const float *__restrict__ cam, float px, float py, float pz,
                                              float &ox, float &oy, float &oz) {
    const float cx = __fmaf_rn(cam[2], pz, __fmaf_rn(cam[1], py, __fmaf_rn(cam[0], px, cam[9])));
    const float cy = __fmaf_rn(cam[5], pz, __fmaf_rn(cam[4], py, __fmaf_rn(cam[3], px, cam[10])));
    const float cz = __fmaf_rn(cam[8], pz, __fmaf_rn(cam[7], py, __fmaf_rn(cam[6], px, cam[11])));
    const float depth = -cz;
    ox = __fdiv_rn(__fmul_rn(cam[12], cx), depth);
    oy = __fdiv_rn(__fmul_rn(cam[13], cy), depth);
    oz = __fsub_rn(cam[14], __fdiv_rn(cam[15], depth));
}

// pass 1: ndc[V][N][3] + per-view min/max of ndc.xy (ordered-key atomics: exact and order independent)
__global__ void __launch_bounds__(DP_THREADS) project_kernel(const float *__restrict__ cams,
                                                             const float *__restrict__ xyz, int N,
                                                             float *__restrict__ ndc, unsigned *__restrict__ keys) {
    __shared__ float scam[16];
    __shared__ unsigned red[4][DP_THREADS / 32];
    const int v = blockIdx.y;
    if (threadIdx.x < 16) scam[threadIdx.x] = cams[v * 16 + threadIdx.x];
    __syncthreads();
    unsigned mnx = 0xffffffffu, mny = 0xffffffffu, mxx = 0u, mxy = 0u;
    for (int i = blockIdx.x * DP_THREADS + threadIdx.x; i < N; i += gridDim.x * DP_THREADS) {
        float ox, oy, oz;
        project_point(scam, __ldg(xyz + i * 3), __ldg(xyz + i * 3 + 1), __ldg(xyz + i * 3 + 2), ox, oy, oz);
        float *o = ndc + ((size_t)v * N + i) * 3;
        o[0] = ox, o[1] = oy, o[2] = oz;
        if (ox == ox) {
            const unsigned k = f2key(ox);
            mnx = min(mnx, k), mxx = max(mxx, k);
        }
        if (oy == oy) {
            const unsigned k = f2key(oy);
            mny = min(mny, k), mxy = max(mxy, k);
        }
    }
    mnx = __reduce_min_sync(0xffffffffu, mnx);
    mny = __reduce_min_sync(0xffffffffu, mny);
    mxx = __reduce_max_sync(0xffffffffu, mxx);
    mxy = __reduce_max_sync(0xffffffffu, mxy);
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    if (lane == 0) red[0][warp] = mnx, red[1][warp] = mny, red[2][warp] = mxx, red[3][warp] = mxy;
    __syncthreads();
    if (threadIdx.x < 4) {
        unsigned r = red[threadIdx.x][0];
        for (int w = 1; w < DP_THREADS / 32; ++w)
            r = (threadIdx.x < 2) ? min(r, red[threadIdx.x][w]) : max(r, red[threadIdx.x][w]);
        if (threadIdx.x < 2)
            atomicMin(keys + v * 4 + threadIdx.x, r);
        else
            atomicMax(keys + v * 4 + threadIdx.x, r);
    }
}

__device__ __forceinline__ void view_bounds(const unsigned *__restrict__ keys, int v, float padding, float &cx,
                                            float &cy, float &sc, float &k) {
    const float mnx = key2f(keys[v * 4 + 0]), mny = key2f(keys[v * 4 + 1]);
    const float mxx = key2f(keys[v * 4 + 2]), mxy = key2f(keys[v * 4 + 3]);
    cx = __fdiv_rn(__fadd_rn(mnx, mxx), 2.f);
    cy = __fdiv_rn(__fadd_rn(mny, mxy), 2.f);
    const float rx = __fsub_rn(mxx, mnx), ry = __fsub_rn(mxy, mny);
    sc = rx > ry ? rx : ry;
    k = __fsub_rn(1.0f, __fmul_rn(2.0f, padding));
}

// pass 2: uv[V][N][2] (+ bounds[V][4] = cx, cy, scale, k)
__global__ void __launch_bounds__(DP_THREADS) uv_kernel(const float *__restrict__ ndc, const unsigned *__restrict__ keys,
                                                        int N, int rescale, float padding, float *__restrict__ uv,
                                                        float *__restrict__ bounds) {
    const int v = blockIdx.y;
    float cx, cy, sc, k;
    view_bounds(keys, v, padding, cx, cy, sc, k);
    if (blockIdx.x == 0 && threadIdx.x == 0 && bounds != nullptr) {
        bounds[v * 4 + 0] = cx, bounds[v * 4 + 1] = cy, bounds[v * 4 + 2] = sc, bounds[v * 4 + 3] = k;
    }
    for (int i = blockIdx.x * DP_THREADS + threadIdx.x; i < N; i += gridDim.x * DP_THREADS) {
        const float *o = ndc + ((size_t)v * N + i) * 3;
        float2 w;
        if (rescale) {
            w.x = __fadd_rn(__fmul_rn(__fdiv_rn(__fsub_rn(o[0], cx), sc), k), 0.5f);
            w.y = __fadd_rn(__fmul_rn(__fdiv_rn(__fsub_rn(o[1], cy), sc), k), 0.5f);
        } else {
            w.x = __fmul_rn(__fadd_rn(o[0], 1.0f), 0.5f);
            w.y = __fmul_rn(__fadd_rn(o[1], 1.0f), 0.5f);
        }
        reinterpret_cast<float2 *>(uv)[(size_t)v * N + i] = w;
    }
}

// z-buffer splat: one thread per (view, point); also reduces ndc_z min/max over the painted points
__global__ void __launch_bounds__(DP_THREADS) zsplat_kernel(const float *__restrict__ uv, const float *__restrict__ ndc,
                                                            const unsigned char *__restrict__ valid, int N, int res,
                                                            int point_size, unsigned long long *__restrict__ zbuf,
                                                            unsigned *__restrict__ zkeys) {
    const int v = blockIdx.y;
    unsigned zmn = 0xffffffffu, zmx = 0u;
    for (int i = blockIdx.x * DP_THREADS + threadIdx.x; i < N; i += gridDim.x * DP_THREADS) {
        if (valid != nullptr && !valid[(size_t)v * N + i]) continue;
        const float2 w = reinterpret_cast<const float2 *>(uv)[(size_t)v * N + i];
        const float z = ndc[((size_t)v * N + i) * 3 + 2];
        const float fu = __fmul_rn(w.x, (float)res), fv = __fmul_rn(w.y, (float)res);
        if (!(fu == fu) || !(fv == fv) || !(z == z)) continue;
        int col = __float2int_rz(fu), row = __float2int_rz(fv);  // saturating truncation toward zero
        col = min(max(col, 0), res - 1);
        row = min(max(row, 0), res - 1);
        const unsigned zk = f2key(z);
        zmn = min(zmn, zk), zmx = max(zmx, zk);
        const unsigned long long word = ((unsigned long long)zk << 32) | (unsigned)i;
        for (int dr = -point_size + 1; dr < point_size; ++dr) {
            const int r = row + dr;
            if (r < 0 || r >= res) continue;
            for (int dc = -point_size + 1; dc < point_size; ++dc) {
                const int c = col + dc;
                if (c < 0 || c >= res) continue;
                atomicMin(zbuf + ((size_t)v * res + (res - 1 - r)) * res + c, word);  // vertical flip (:339)
            }
        }
    }
    zmn = __reduce_min_sync(0xffffffffu, zmn);
    zmx = __reduce_max_sync(0xffffffffu, zmx);
    if ((threadIdx.x & 31) == 0) {
        if (zmn != 0xffffffffu) atomicMin(zkeys + v * 2, zmn);
        if (zmx != 0u) atomicMax(zkeys + v * 2 + 1, zmx);
    }
}

// resolve: idx image, depth image 0.1+0.8*(1-(z-zmin)/(zmax-zmin)) (getRawDepth :362-366), optional colours
__global__ void __launch_bounds__(DP_THREADS) zresolve_kernel(const unsigned long long *__restrict__ zbuf,
                                                              const float *__restrict__ ndc,
                                                              const unsigned *__restrict__ zkeys,
                                                              const float *__restrict__ colors, int N, int res,
                                                              int *__restrict__ idx_img, float *__restrict__ depth_img,
                                                              float *__restrict__ color_img,
                                                              float *__restrict__ zminmax) {
    const int v = blockIdx.y;
    const float zmin = key2f(zkeys[v * 2]), zmax = key2f(zkeys[v * 2 + 1]);
    const float range = __fsub_rn(zmax, zmin);
    if (blockIdx.x == 0 && threadIdx.x == 0 && zminmax != nullptr) zminmax[v * 2] = zmin, zminmax[v * 2 + 1] = zmax;
    const size_t npix = (size_t)res * res;
    for (size_t p = (size_t)blockIdx.x * DP_THREADS + threadIdx.x; p < npix; p += (size_t)gridDim.x * DP_THREADS) {
        const unsigned long long w = zbuf[v * npix + p];
        int i = -1;
        float dep = 0.f;
        if (w != ~0ull) {
            i = (int)(unsigned)(w & 0xffffffffu);
            const float z = ndc[((size_t)v * N + i) * 3 + 2];
            dep = __fadd_rn(0.1f, __fmul_rn(0.8f, __fsub_rn(1.0f, __fdiv_rn(__fsub_rn(z, zmin), range))));
        }
        if (idx_img != nullptr) idx_img[v * npix + p] = i;
        if (depth_img != nullptr) depth_img[v * npix + p] = dep;
        if (color_img != nullptr) {
#pragma unroll
            for (int c = 0; c < 3; ++c) color_img[(v * 3 + c) * npix + p] = (i >= 0) ? __ldg(colors + (size_t)i * 3 + c) : 0.f;
        }
    }
}

// unprojection with raster-order compaction: one 1024-thread CTA per view
__global__ void __launch_bounds__(1024) unproject_kernel(const float *__restrict__ cams, const float *__restrict__ bounds,
                                                         int rescale, const unsigned long long *__restrict__ zbuf,
                                                         const float *__restrict__ ndc, int N, int res,
                                                         float *__restrict__ out, int *__restrict__ own,
                                                         int *__restrict__ counts) {
    __shared__ int wsum[32];
    __shared__ float scam[16];
    const int v = blockIdx.x, tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    if (tid < 16) scam[tid] = cams[v * 16 + tid];
    const size_t npix = (size_t)res * res;
    const size_t per = (npix + 1023) / 1024;
    const size_t p0 = min(npix, per * tid), p1 = min(npix, p0 + per);
    const unsigned long long *zb = zbuf + v * npix;
    int cnt = 0;
    for (size_t p = p0; p < p1; ++p) cnt += (zb[p] != ~0ull);
    int incl = cnt;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
        const int t = __shfl_up_sync(0xffffffffu, incl, o);
        if (lane >= o) incl += t;
    }
    if (lane == 31) wsum[warp] = incl;
    __syncthreads();
    if (warp == 0) {
        int w = wsum[lane], wi = w;
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) {
            const int t = __shfl_up_sync(0xffffffffu, wi, o);
            if (lane >= o) wi += t;
        }
        wsum[lane] = wi - w;  // exclusive
        if (lane == 31) counts[v] = wi;
    }
    __syncthreads();
    int pos = wsum[warp] + incl - cnt;
    const float cx = bounds[v * 4 + 0], cy = bounds[v * 4 + 1], sc = bounds[v * 4 + 2], k = bounds[v * 4 + 3];
    for (size_t p = p0; p < p1; ++p) {
        const unsigned long long w = zb[p];
        if (w == ~0ull) continue;
        const int i = (int)(unsigned)(w & 0xffffffffu);
        const int rs = (int)(p / res), c = (int)(p - (size_t)rs * res);
        const float z = ndc[((size_t)v * N + i) * 3 + 2];
        const float u = __fdiv_rn(__fadd_rn((float)c, 0.5f), (float)res);
        const float vv = __fdiv_rn(__fadd_rn((float)(res - 1 - rs), 0.5f), (float)res);
        float nx, ny;
        if (rescale) {
            nx = __fadd_rn(__fmul_rn(__fdiv_rn(__fsub_rn(u, 0.5f), k), sc), cx);
            ny = __fadd_rn(__fmul_rn(__fdiv_rn(__fsub_rn(vv, 0.5f), k), sc), cy);
        } else {
            nx = __fsub_rn(__fmul_rn(u, 2.0f), 1.0f);
            ny = __fsub_rn(__fmul_rn(vv, 2.0f), 1.0f);
        }
        const float depth = __fdiv_rn(scam[15], __fsub_rn(scam[14], z));
        const float dx = __fsub_rn(__fdiv_rn(__fmul_rn(nx, depth), scam[12]), scam[9]);
        const float dy = __fsub_rn(__fdiv_rn(__fmul_rn(ny, depth), scam[13]), scam[10]);
        const float dz = __fsub_rn(-depth, scam[11]);
        float *o = out + (v * npix + pos) * 3;
        o[0] = __fmaf_rn(scam[6], dz, __fmaf_rn(scam[3], dy, __fmul_rn(scam[0], dx)));
        o[1] = __fmaf_rn(scam[7], dz, __fmaf_rn(scam[4], dy, __fmul_rn(scam[1], dx)));
        o[2] = __fmaf_rn(scam[8], dz, __fmaf_rn(scam[5], dy, __fmul_rn(scam[2], dx)));
        own[v * npix + pos] = i;
        ++pos;
    }
}

static inline unsigned grid_for(int n) {
    int g = (n + DP_THREADS - 1) / DP_THREADS;
    const int cap = GENPC_NUM_SMS * 8;
    return (unsigned)(g < 1 ? 1 : (g > cap ? cap : g));
}

}  // namespace genpc

using namespace genpc;

extern "C" size_t genpc_depth_workspace_bytes(int V) { return V < 0 ? 0 : (size_t)V * 6 * sizeof(unsigned); }

extern "C" int genpc_project_uv(const float *cams, const float *xyz, int V, int N, int rescale, float padding,
                                float *ndc, float *uv, float *bounds, void *workspace, size_t workspace_bytes,
                                genpc_stream_t stream_) {
    cudaStream_t stream = (cudaStream_t)stream_;
    if (V < 0 || N < 0) return GENPC_ERR_SHAPE;
    if ((reinterpret_cast<size_t>(uv) & 7) != 0) return GENPC_ERR_SHAPE;  // uv is accessed as float2 (8-byte aligned pairs)
    if (V == 0 || N == 0) return GENPC_OK;
    if (workspace == nullptr || workspace_bytes < genpc_depth_workspace_bytes(V)) return GENPC_ERR_WORKSPACE;
    unsigned *keys = (unsigned *)workspace;  // [V][4]: min x, min y, max x, max y
    // min keys start at 0xffffffff, max keys at 0: write the pattern with one 2D memset per half
    cudaError_t e = cudaMemset2DAsync(keys, 16, 0xff, 8, V, stream);
    if (e != cudaSuccess) return (int)e;
    e = cudaMemset2DAsync(keys + 2, 16, 0x00, 8, V, stream);
    if (e != cudaSuccess) return (int)e;
    dim3 grid(grid_for(N), V);
    project_kernel<<<grid, DP_THREADS, 0, stream>>>(cams, xyz, N, ndc, keys);
    GENPC_CHECK_LAUNCH();
    uv_kernel<<<grid, DP_THREADS, 0, stream>>>(ndc, keys, N, rescale, padding, uv, bounds);
    GENPC_CHECK_LAUNCH();
    return GENPC_OK;
}

extern "C" int genpc_zbuffer_render(const float *uv, const float *ndc, const unsigned char *valid,
                                    const float *colors, int V, int N, int res, int point_size,
                                    unsigned long long *zbuf, int *idx_img, float *depth_img, float *color_img,
                                    float *zminmax, void *workspace, size_t workspace_bytes, genpc_stream_t stream_) {
    cudaStream_t stream = (cudaStream_t)stream_;
    if (V < 0 || N < 0 || res <= 0 || point_size < 1) return GENPC_ERR_SHAPE;
    if ((reinterpret_cast<size_t>(uv) & 7) != 0) return GENPC_ERR_SHAPE;  // uv is accessed as float2 (8-byte aligned pairs)
    if (V == 0) return GENPC_OK;
    if (workspace == nullptr || workspace_bytes < genpc_depth_workspace_bytes(V)) return GENPC_ERR_WORKSPACE;
    unsigned *zkeys = (unsigned *)workspace + (size_t)V * 4;  // [V][2]: min z key, max z key
    cudaError_t e = cudaMemsetAsync(zbuf, 0xff, (size_t)V * res * res * 8, stream);
    if (e != cudaSuccess) return (int)e;
    e = cudaMemset2DAsync(zkeys, 8, 0xff, 4, V, stream);
    if (e != cudaSuccess) return (int)e;
    e = cudaMemset2DAsync(zkeys + 1, 8, 0x00, 4, V, stream);
    if (e != cudaSuccess) return (int)e;
    if (N > 0) {
        dim3 grid(grid_for(N), V);
        zsplat_kernel<<<grid, DP_THREADS, 0, stream>>>(uv, ndc, valid, N, res, point_size, zbuf, zkeys);
        GENPC_CHECK_LAUNCH();
    }
    if (idx_img != nullptr || depth_img != nullptr || color_img != nullptr || zminmax != nullptr) {
        if (color_img != nullptr && colors == nullptr) return GENPC_ERR_SHAPE;
        dim3 grid(grid_for(res * res), V);
        zresolve_kernel<<<grid, DP_THREADS, 0, stream>>>(zbuf, ndc, zkeys, colors, N, res, idx_img, depth_img, color_img,
                                                         zminmax);
        GENPC_CHECK_LAUNCH();
    }
    return GENPC_OK;
}

extern "C" int genpc_unproject(const float *cams, const float *bounds, int rescale, const unsigned long long *zbuf,
                               const float *ndc, int V, int N, int res, float *out, int *own, int *counts,
                               genpc_stream_t stream_) {
    cudaStream_t stream = (cudaStream_t)stream_;
    if (V < 0 || N < 0 || res <= 0) return GENPC_ERR_SHAPE;
    if (V == 0) return GENPC_OK;
    unproject_kernel<<<V, 1024, 0, stream>>>(cams, bounds, rescale, zbuf, ndc, N, res, out, own, counts);
    GENPC_CHECK_LAUNCH();
    return GENPC_OK;
}
