// nn_core.cuh -- the nearest-neighbour scan work item shared by the Chamfer forward (chamfer.cu), the fused
// registration step (register.cu) and the target-sharded Chamfer (sharded path in chamfer.cu).
//
// A work item scans NN_THREADS*QT queries against NN_SPAN targets:
//   * targets staged once into shared memory as SoA x[]/y[]/z[] (NaN padded; NaN never wins a min),
//     optionally pushed through a similarity transform while staging (registration: the moving cloud);
//   * inner loop on the packed FP32 pipe (FADD2/FMUL2/FFMA2, two targets per instruction) with one FMNMX3 per
//     two pairs; distance rounding = the reference's fma(dz,dz,fma(dx,dx,dy*dy)) (chamfer3D.cu:35);
//   * only the id of the 16-target chunk that lowered the running minimum is tracked (strict `<`: earliest
//     chunk keeps ties); the exact lowest index is recovered by re-scanning that chunk;
//   * result merged with a packed 64-bit atomicMin ((dist_bits<<32)|idx): dist ascending, then idx ascending,
//     i.e. the reference's lowest-index tie rule (chamfer3D.cu:46,:126).
#pragma once
#include "common.cuh"

namespace genpc {

// tunables (overridable for tools/nn_variants.sh experiments; the defaults are what ships)
#ifndef GENPC_NN_SPAN
#define GENPC_NN_SPAN 1024
#endif
#ifndef GENPC_NN_CHUNK
#define GENPC_NN_CHUNK 8
#endif
#ifndef GENPC_NN_MINBLOCKS
#define GENPC_NN_MINBLOCKS 2
#endif
#ifndef GENPC_NN_QT_MAX
#define GENPC_NN_QT_MAX 4
#endif
constexpr int NN_THREADS = 256;
constexpr int NN_SPAN = GENPC_NN_SPAN;    // targets per work item (12 B of shared memory each)
constexpr int NN_CHUNK = GENPC_NN_CHUNK;  // index-recovery granularity
constexpr int NN_MINBLOCKS = GENPC_NN_MINBLOCKS;

// p' = R (s (p - c)) + c + t with explicit rounding (ObjectPoseOptim.forward, diff_obj_pose.py:419-423):
//   l = p - c; u = l * s; r_x = fma(R02,u_z, fma(R01,u_y, R00*u_x)); p'_x = (r_x + c_x) + t_x
struct Similarity {
    float R[9];
    float s;
    float c[3];
    float t[3];
};

__device__ __forceinline__ void apply_similarity(const Similarity &T, float &x, float &y, float &z) {
    const float ux = __fmul_rn(__fsub_rn(x, T.c[0]), T.s);
    const float uy = __fmul_rn(__fsub_rn(y, T.c[1]), T.s);
    const float uz = __fmul_rn(__fsub_rn(z, T.c[2]), T.s);
    const float rx = __fmaf_rn(T.R[2], uz, __fmaf_rn(T.R[1], uy, __fmul_rn(T.R[0], ux)));
    const float ry = __fmaf_rn(T.R[5], uz, __fmaf_rn(T.R[4], uy, __fmul_rn(T.R[3], ux)));
    const float rz = __fmaf_rn(T.R[8], uz, __fmaf_rn(T.R[7], uy, __fmul_rn(T.R[6], ux)));
    x = __fadd_rn(__fadd_rn(rx, T.c[0]), T.t[0]);
    y = __fadd_rn(__fadd_rn(ry, T.c[1]), T.t[1]);
    z = __fadd_rn(__fadd_rn(rz, T.c[2]), T.t[2]);
}

// rotation_6d_to_matrix (pytorch3d semantics restated): rows b1, b2, b1 x b2.
// F.normalize: v / max(|v|, 1e-12).  Every step explicitly rounded; mirrored by oracle_pose_matrix.
__device__ __forceinline__ void rot6d_to_matrix(const float *d6, float *R) {
    const float a1x = d6[0], a1y = d6[1], a1z = d6[2], a2x = d6[3], a2y = d6[4], a2z = d6[5];
    const float n1 = fmaxf(__fsqrt_rn(__fmaf_rn(a1z, a1z, __fmaf_rn(a1y, a1y, __fmul_rn(a1x, a1x)))), 1e-12f);
    const float b1x = __fdiv_rn(a1x, n1), b1y = __fdiv_rn(a1y, n1), b1z = __fdiv_rn(a1z, n1);
    const float dp = __fmaf_rn(b1z, a2z, __fmaf_rn(b1y, a2y, __fmul_rn(b1x, a2x)));
    const float wx = __fsub_rn(a2x, __fmul_rn(dp, b1x)), wy = __fsub_rn(a2y, __fmul_rn(dp, b1y)),
                wz = __fsub_rn(a2z, __fmul_rn(dp, b1z));
    const float n2 = fmaxf(__fsqrt_rn(__fmaf_rn(wz, wz, __fmaf_rn(wy, wy, __fmul_rn(wx, wx)))), 1e-12f);
    const float b2x = __fdiv_rn(wx, n2), b2y = __fdiv_rn(wy, n2), b2z = __fdiv_rn(wz, n2);
    R[0] = b1x, R[1] = b1y, R[2] = b1z;
    R[3] = b2x, R[4] = b2y, R[5] = b2z;
    R[6] = __fsub_rn(__fmul_rn(b1y, b2z), __fmul_rn(b1z, b2y));
    R[7] = __fsub_rn(__fmul_rn(b1z, b2x), __fmul_rn(b1x, b2z));
    R[8] = __fsub_rn(__fmul_rn(b1x, b2y), __fmul_rn(b1y, b2x));
}

// One work item.  q/t point at the first query / target of the CLOUD (batch offset applied by the caller);
// j0 = first query of the tile, t0 = first target of the span, idx_base = value added to the reported index
// (0 for plain Chamfer; the shard offset for target-sharded clouds).  qT / tT: optional similarity applied to
// the queries / targets on the fly (nullptr = identity).  out = packed words of this cloud's queries.
// SPAN (a multiple of NN_CHUNK): targets per item -- NN_SPAN everywhere except the persistent small-registration kernel, which
// uses shorter spans to spread a scan over more CTAs (the packed atomicMin merge makes the result independent of the split)
template <int QT, int SPAN = NN_SPAN>
__device__ __forceinline__ void nn_scan_item(float (*s)[SPAN], const float *__restrict__ q, int nq, int j0,
                                             const float *__restrict__ t, int mt, int t0, int idx_base,
                                             const Similarity *qT, const Similarity *tT,
                                             unsigned long long *__restrict__ out) {
    const int tid = threadIdx.x;
    const int cnt = min(SPAN, mt - t0);
    // ---- stage targets: AoS global -> SoA shared ----
    {
        const float qnan = __int_as_float(0x7fc00000);
        for (int k = tid; k < SPAN; k += NN_THREADS) {
            float x = qnan, y = qnan, z = qnan;
            if (k < cnt) {
                const float *tp = t + (size_t)(t0 + k) * 3;
                x = __ldg(tp), y = __ldg(tp + 1), z = __ldg(tp + 2);
                if (tT != nullptr) apply_similarity(*tT, x, y, z);
            }
            s[0][k] = x, s[1][k] = y, s[2][k] = z;
        }
    }
    // ---- queries into registers (negated, broadcast into both halves of the packed ops) ----
    float2 nqx[QT], nqy[QT], nqz[QT];
    float best[QT];
    int bchunk[QT];
    const int jbase = j0 + tid;
#pragma unroll
    for (int qi = 0; qi < QT; ++qi) {
        const int j = jbase + qi * NN_THREADS;
        float x = 0.f, y = 0.f, z = 0.f;
        if (j < nq) {
            const float *qp = q + (size_t)j * 3;
            x = __ldg(qp), y = __ldg(qp + 1), z = __ldg(qp + 2);
            if (qT != nullptr) apply_similarity(*qT, x, y, z);
        }
        nqx[qi] = make_float2(-x, -x);
        nqy[qi] = make_float2(-y, -y);
        nqz[qi] = make_float2(-z, -z);
        best[qi] = __int_as_float(0x7f800000);
        bchunk[qi] = 0;
    }
    __syncthreads();

    const int nchunks = (cnt + NN_CHUNK - 1) / NN_CHUNK;
    const float4 *sx4 = reinterpret_cast<const float4 *>(s[0]);
    const float4 *sy4 = reinterpret_cast<const float4 *>(s[1]);
    const float4 *sz4 = reinterpret_cast<const float4 *>(s[2]);
    for (int c = 0; c < nchunks; ++c) {
        float cm[QT];
#pragma unroll
        for (int qi = 0; qi < QT; ++qi) cm[qi] = __int_as_float(0x7f800000);
#pragma unroll
        for (int kk = 0; kk < NN_CHUNK / 4; ++kk) {
            const float4 X = sx4[c * (NN_CHUNK / 4) + kk];
            const float4 Y = sy4[c * (NN_CHUNK / 4) + kk];
            const float4 Z = sz4[c * (NN_CHUNK / 4) + kk];
#pragma unroll
            for (int qi = 0; qi < QT; ++qi) {
                const float2 a = sqdist_ref_x2(nqx[qi], nqy[qi], nqz[qi], make_float2(X.x, X.y),
                                               make_float2(Y.x, Y.y), make_float2(Z.x, Z.y));
                const float2 e = sqdist_ref_x2(nqx[qi], nqy[qi], nqz[qi], make_float2(X.z, X.w),
                                               make_float2(Y.z, Y.w), make_float2(Z.z, Z.w));
                cm[qi] = fmin3(cm[qi], a.x, a.y);
                cm[qi] = fmin3(cm[qi], e.x, e.y);
            }
        }
#pragma unroll
        for (int qi = 0; qi < QT; ++qi) {
            if (cm[qi] < best[qi]) {  // strict: the earliest chunk keeps ties
                best[qi] = cm[qi];
                bchunk[qi] = c;
            }
        }
    }

    // ---- recover the exact (lowest) index inside the winning chunk, merge across target spans ----
#pragma unroll
    for (int qi = 0; qi < QT; ++qi) {
        const int j = jbase + qi * NN_THREADS;
        if (j >= nq) continue;
        const float qx = -nqx[qi].x, qy = -nqy[qi].x, qz = -nqz[qi].x;
        const int cb = bchunk[qi] * NN_CHUNK;
        int kbest = 0;
#pragma unroll
        for (int k = NN_CHUNK - 1; k >= 0; --k) {
            const float dd = sqdist_ref(qx, qy, qz, s[0][cb + k], s[1][cb + k], s[2][cb + k]);
            if (dd == best[qi]) kbest = k;
        }
        atomicMin(out + j, pack_dist_idx(best[qi], idx_base + t0 + cb + kbest));
    }
}

static inline int nn_pick_qt(int nq) {
    // queries per thread: large tiles amortise the shared-memory reads, small clouds keep lanes busy
    if (GENPC_NN_QT_MAX >= 4 && nq >= 4 * NN_THREADS) return 4;
    if (GENPC_NN_QT_MAX >= 2 && nq >= 2 * NN_THREADS) return 2;
    return 1;
}

}  // namespace genpc
