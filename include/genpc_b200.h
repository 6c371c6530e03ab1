/*
 * genpc_b200.h -- C ABI of libgenpc_b200.so, the B200 (sm_100a) drop-in for GenPC's geometric hot path.
 *
 * Conventions (mirroring the reference's pybind surface):
 *   - every pointer is a DEVICE pointer unless the name ends in `_host`;
 *   - the caller owns every buffer, including outputs and workspaces; nothing is allocated or freed here
 *     (reference: dist_chamfer_3D.py:33-42, emd_module.py:43-54);
 *   - clouds are contiguous AoS float[B][N][3] (chamfer3D.cu:19,23-25), indices are int32;
 *   - all work is enqueued on `stream` (a cudaStream_t; pass 0 for the legacy default stream the
 *     reference uses, chamfer3D.cu:142) and the call returns without synchronising;
 *   - return value: 0 = OK, >0 = cudaError_t of the failed launch, <0 = argument violation
 *     (GENPC_ERR_*).  The reference returns 1/0/-1 and prints (chamfer3D.cu:145-151,
 *     emd_cuda.cu:236-249); the Python mirrors translate.
 * There is no CPU fallback: without a CUDA device every entry point returns an error.
 */
#ifndef GENPC_B200_H
#define GENPC_B200_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

typedef void *genpc_stream_t; /* cudaStream_t */

#define GENPC_OK 0
#define GENPC_ERR_SHAPE (-1)     /* negative sizes, n != m, n % 256, B > 512 ... */
#define GENPC_ERR_WORKSPACE (-2) /* workspace NULL or too small */
#define GENPC_ERR_RANGE (-3)     /* B*N does not fit the 32-bit index space of the reference */

/* Library / build identification: returns a static string such as "genpc_b200 0.1 sm_100a". */
const char *genpc_version(void);

/* ---- Chamfer3D -------------------------------------------------------------------------------
 * Replaces chamfer_3D.forward (chamfer_cuda.cpp:17-19 -> chamfer_cuda_forward, chamfer3D.cu:136-154,
 * kernel NmDistanceKernel :12-134).
 *   dist1[b,j] = min_k |xyz1[b,j]-xyz2[b,k]|^2, idx1 = argmin (lowest k on ties); dist2/idx2 symmetric.
 *   Distance rounding is the reference's: fma(dz,dz,fma(dx,dx,dy*dy)).  Bit-exact outputs.
 * workspace: genpc_chamfer_workspace_bytes(B,N,M) bytes of device scratch (packed (dist,idx) words). */
size_t genpc_chamfer_workspace_bytes(int B, int N, int M);
int genpc_chamfer_forward(const float *xyz1, const float *xyz2, float *dist1, float *dist2, int *idx1,
                          int *idx2, int B, int N, int M, void *workspace, size_t workspace_bytes,
                          genpc_stream_t stream);

/* Replaces chamfer_3D.backward (chamfer_cuda.cpp:22-27 -> chamfer_cuda_backward, chamfer3D.cu:176-195,
 * kernel NmDistanceGradKernel :155-174).  ACCUMULATES into gradxyz1/gradxyz2, which must arrive zeroed
 * exactly as in the reference (dist_chamfer_3D.py:56-57). */
int genpc_chamfer_backward(const float *xyz1, const float *xyz2, const float *graddist1,
                           const float *graddist2, const int *idx1, const int *idx2, float *gradxyz1,
                           float *gradxyz2, int B, int N, int M, genpc_stream_t stream);

#ifdef __cplusplus
}
#endif
#endif /* GENPC_B200_H */
