// emd.cu -- EMD approximate matching (auction algorithm) for sm_100a: ONE persistent cooperative kernel.
//
// Replaces emd_cuda_forward (emd_cuda.cu:228-282): the reference runs `iters` x 7 launches (clear,
// calc_unass_cnt, calc_unass_cnt_sum, calc_unass_idx, Bid, GetMax, Assign) + CalcDist = 351 launches at
// iters = 50, launch-latency bound, and every Bid block re-stages the whole target cloud for <= 256 bidders.
// Here a group of CTAs owns each batch entry for the whole run:
//   * Bid is spread over the group: an item = P unassigned points x all targets, T = 256/P threads per point
//     (P adapts so the items fill the group); targets + prices staged as float4 chunks in shared memory;
//   * the last CTA of the group to finish an iteration's items (atomic ticket) runs the O(n) tail on its own:
//     GetMax -> Assign -> reset -> ascending compaction of the still-unassigned points, then releases a
//     per-batch flag the group spins on.  No grid-wide barrier, no host round trip.
// Arithmetic and tie rules are the reference's, bit for bit (oracle_emd_forward has the derivation):
//   value = (float)((3.0 - (double)sqrtf(fma(dz,dz,fma(dx,dx,dy*dy)))) - (double)price)   (emd_cuda.cu:146)
//   best / second-best with multiplicity; among equal best values the winner is the target with the smallest
//   (reference_thread(k), k) -- the order in which the reference's thread slices visit targets (:136-139,:167);
//   GetMax window +-1e-6 in double (:188); the reference's last-writer race is resolved as "highest j".
#include <cooperative_groups.h>

#include "common.cuh"

namespace genpc {

constexpr int EMD_THREADS = 256;
constexpr int EMD_CHUNK = 2048;  // same chunking as the reference's Bid (it defines the tie order)

struct EmdArgs {
    const float *xyz1, *xyz2;
    float *dist;
    int *assignment;
    float *price;
    int *assignment_inv, *bid;
    float *bid_increments, *max_increments;
    int *unass_idx, *unass_cnt, *max_idx;
    int *flags, *tickets;  // workspace [B] each, zeroed by the host
    int B, n;
    float eps;
    int iters, group;      // group = CTAs per batch entry (1 when B >= grid)
};

struct BidState {
    float best, better;
    int bi;
};

// order in which the reference visits target k: (thread slot inside its 2048-chunk, k)
__device__ __forceinline__ long long ref_visit_key(int k, int n, int tpu_ref) {
    const int kl = k & (EMD_CHUNK - 1);
    const int end_k = min(EMD_CHUNK, n - (k - kl));
    const int delta = (end_k + tpu_ref - 1) / tpu_ref;
    return ((long long)(kl / delta) << 32) | (unsigned)k;
}

__device__ __forceinline__ void bid_merge(BidState &a, const BidState &b, int n, int tpu_ref) {
    bool take_b;
    if (b.best > a.best) take_b = true;
    else if (b.best < a.best) take_b = false;
    else if (b.bi < 0) take_b = false;
    else if (a.bi < 0) take_b = true;
    else take_b = ref_visit_key(b.bi, n, tpu_ref) < ref_visit_key(a.bi, n, tpu_ref);
    if (take_b) {
        a.better = fmaxf(b.better, a.best);
        a.best = b.best;
        a.bi = b.bi;
    } else {
        a.better = fmaxf(a.better, b.best);
    }
}

__device__ __forceinline__ int ld_acquire(const int *p) {
    int v;
    asm volatile("ld.acquire.gpu.global.s32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
    return v;
}
__device__ __forceinline__ void st_release(int *p, int v) {
    asm volatile("st.release.gpu.global.s32 [%0], %1;" ::"l"(p), "r"(v) : "memory");
}

// ascending compaction of {j : assignment[j] == -1} into uidx; returns the count (valid in every thread).
// Warp w owns the contiguous slice [w * n/8, (w+1) * n/8): pass 1 counts it with coalesced loads + ballots (no block
// barrier inside, loads independent), one barrier publishes the eight warp totals, pass 2 re-reads the slice (L1) and
// writes the indices.  (The first form walked the array 256 elements at a time with two block barriers per step:
// 64 barriers at n = 8192, most of an iteration's serial tail.)
__device__ int compact_unassigned(const int *__restrict__ asg, int *__restrict__ uidx, int *__restrict__ midx, int n,
                                  int *sscan) {
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    constexpr int WARPS = EMD_THREADS / 32;
    const int per_warp = n / WARPS;  // n % 256 == 0
    const int j0 = warp * per_warp;
    int cnt = 0;
#pragma unroll 4
    for (int r = 0; r < per_warp; r += 32) {
        const int j = j0 + r + lane;
        const int f = (__ldcg(asg + j) == -1);
        midx[j] = -1;  // re-arm GetMax for the next iteration
        cnt += __popc(__ballot_sync(0xffffffffu, f));
    }
    if (lane == 0) sscan[warp] = cnt;
    __syncthreads();
    int base = 0, tot = 0;
#pragma unroll
    for (int w = 0; w < WARPS; ++w) {
        const int c = sscan[w];
        base += (w < warp) ? c : 0;
        tot += c;
    }
    if (cnt > 0) {
#pragma unroll 4
        for (int r = 0; r < per_warp; r += 32) {
            const int j = j0 + r + lane;
            const int f = (__ldcg(asg + j) == -1);
            const unsigned bal = __ballot_sync(0xffffffffu, f);
            if (f) uidx[base + __popc(bal & ((1u << lane) - 1u))] = j;
            base += __popc(bal);
        }
    }
    __syncthreads();  // sscan may be reused
    return tot;
}

__global__ void __launch_bounds__(EMD_THREADS) emd_auction_kernel(const EmdArgs a) {
    __shared__ __align__(16) float sx[EMD_CHUNK], sy[EMD_CHUNK], sz[EMD_CHUNK];  // targets, SoA (pairs feed FADD2/FFMA2)
    __shared__ __align__(16) float sc[EMD_CHUNK];                                 // c = fl(3 - price): pre-filter operand
    __shared__ float sprice[EMD_CHUNK];
    __shared__ BidState smerge[EMD_THREADS];
    __shared__ int sscan[EMD_THREADS / 32];
    __shared__ int s_last;
    const int tid = threadIdx.x;
    const int n = a.n;
    const int block_cnt = n / 256;

    for (int b = (a.group > 1 ? (int)blockIdx.x / a.group : (int)blockIdx.x); b < a.B;
         b += (a.group > 1 ? a.B : (int)gridDim.x)) {
        const int rank = a.group > 1 ? (int)blockIdx.x % a.group : 0;
        const float *p1 = a.xyz1 + (size_t)b * n * 3;
        const float *p2 = a.xyz2 + (size_t)b * n * 3;
        int *asg = a.assignment + (size_t)b * n;
        int *asg_inv = a.assignment_inv + (size_t)b * n;
        float *pr = a.price + (size_t)b * n;
        int *bd = a.bid + (size_t)b * n;
        float *binc = a.bid_increments + (size_t)b * n;
        float *minc = a.max_increments + (size_t)b * n;
        int *uidx = a.unass_idx + (size_t)b * n;
        int *midx = a.max_idx + (size_t)b * n;
        int *flag = a.flags + b, *ticket = a.tickets + b;

        if (rank == 0) {  // initial compaction (calc_unass_cnt / calc_unass_idx of iteration 0)
            const int U0 = compact_unassigned(asg, uidx, midx, n, sscan);
            __threadfence();
            __syncthreads();
            if (tid == 0) {
                a.unass_cnt[b] = U0;
                __threadfence();
                st_release(flag, 1);
            }
        }

        bool complete = false;  // every point assigned: the remaining iterations change nothing (reference: empty launches)
        for (int it = 0; it < a.iters; ++it) {
            const bool last = (it == a.iters - 1);
            if (tid == 0) {
                while (ld_acquire(flag) < it + 1) __nanosleep(64);
            }
            __syncthreads();
            const int U = __ldcg(a.unass_cnt + b);
            if (U == 0) {  // same value in every CTA of the group (read behind the same flag generation)
                complete = true;
                break;
            }
            // ---- Bid (emd_cuda.cu:95-179) ----
            if (U > 0) {
                const int upb_ref = (U + block_cnt - 1) / block_cnt;
                const int tpu_ref = 256 / upb_ref;
                // points per item: spread the U bidders over the group; every item re-stages the whole target cloud, so
                // fewer, fuller items win (a rounds*P balance model measured 3x slower at B=32: staging dominated)
                int P = (U + a.group - 1) / a.group;
                P = max(1, min(P, EMD_THREADS));
                const int T = EMD_THREADS / P;
                const int items = (U + P - 1) / P;
                for (int g = rank; g < items; g += a.group) {
                    const int ps = tid / T, tpt = tid - ps * T;
                    const int u = g * P + ps;
                    const bool active = (ps < P) && (u < U);
                    int j = -1;
                    float x1 = 0.f, y1 = 0.f, z1 = 0.f;
                    if (active) {
                        j = __ldcg(uidx + u);
                        x1 = __ldg(p1 + j * 3), y1 = __ldg(p1 + j * 3 + 1), z1 = __ldg(p1 + j * 3 + 2);
                    }
                    BidState st;
                    st.best = -1e9f, st.better = -1e9f, st.bi = -1;
                    float bm = -2e9f;  // better - margin (pre-filter threshold)
                    for (int k2 = 0; k2 < n; k2 += EMD_CHUNK) {
                        const int end_k = min(EMD_CHUNK, n - k2);
                        __syncthreads();
                        for (int k = tid; k < end_k; k += EMD_THREADS) {
                            const float *tp = p2 + (size_t)(k2 + k) * 3;
                            const float pk = __ldcg(pr + k2 + k);
                            sx[k] = __ldg(tp), sy[k] = __ldg(tp + 1), sz[k] = __ldg(tp + 2);
                            sc[k] = __fsub_rn(3.0f, pk);
                            sprice[k] = pk;
                        }
                        __syncthreads();
                        if (active) {
                            // exact evaluation of one candidate (the reference's arithmetic and tie rule)
                            auto consider = [&](int kl, float s) {
                                const float d = (float)((3.0 - (double)__fsqrt_rn(s)) - (double)sprice[kl]);
                                if (d > st.best) {
                                    st.better = st.best;
                                    st.best = d;
                                    st.bi = k2 + kl;
                                } else {
                                    st.better = fmaxf(st.better, d);
                                    if (d == st.best &&
                                        ref_visit_key(k2 + kl, n, tpu_ref) < ref_visit_key(st.bi, n, tpu_ref))
                                        st.bi = k2 + kl;
                                }
                                bm = __fsub_rn(st.better, __fmul_rn(1e-4f, fmaxf(1.f, fabsf(st.better))));
                            };
                            // two targets per step on the packed FP32 pipe.  Conservative pre-filter (no sqrt, no FP64):
                            // a candidate can only matter if its value d >= better, i.e. sqrt(s) <= 3 - price - better up
                            // to a few ulps; bm = better - margin with margin = 1e-4*max(1,|better|) (hundreds of ulps)
                            // makes "tq > 0 && s <= tq^2" a superset of those candidates; whatever passes takes the exact
                            // path above, so the result is bit-identical to evaluating every candidate exactly.
                            const float2 nx = make_float2(-x1, -x1), ny = make_float2(-y1, -y1), nz = make_float2(-z1, -z1);
                            for (int kp = tpt; kp < (end_k >> 1); kp += T) {   // end_k is even (n % 256 == 0)
                                const float2 tx = reinterpret_cast<const float2 *>(sx)[kp];
                                const float2 ty = reinterpret_cast<const float2 *>(sy)[kp];
                                const float2 tz = reinterpret_cast<const float2 *>(sz)[kp];
                                const float2 tc = reinterpret_cast<const float2 *>(sc)[kp];
                                const float2 s2 = sqdist_ref_x2(nx, ny, nz, tx, ty, tz);
                                const float2 tq = __fadd2_rn(tc, make_float2(-bm, -bm));
                                const float2 tq2 = __fmul2_rn(tq, tq);
                                const bool c0 = tq.x > 0.f && s2.x <= tq2.x;
                                const bool c1 = tq.y > 0.f && s2.y <= tq2.y;
                                if (c0) consider(2 * kp, s2.x);
                                if (c1) {
                                    // bm may have risen while handling the first candidate; the filter stays conservative
                                    // because it was evaluated against the OLDER (lower) threshold
                                    consider(2 * kp + 1, s2.y);
                                }
                            }
                        }
                    }
                    // merge the T partial states of each point (tree over shared memory)
                    __syncthreads();
                    smerge[tid] = st;
                    __syncthreads();
                    for (int stride = 1; stride < T; stride <<= 1) {
                        if (active && (tpt % (2 * stride)) == 0 && tpt + stride < T) {
                            BidState mine = smerge[tid];
                            bid_merge(mine, smerge[tid + stride], n, tpu_ref);
                            smerge[tid] = mine;
                        }
                        __syncthreads();
                    }
                    if (active && tpt == 0) {
                        const BidState r = smerge[tid];
                        const float inc = __fadd_rn(__fsub_rn(r.best, r.better), a.eps);
                        bd[j] = r.bi;
                        binc[j] = inc;
                        atomicMax(reinterpret_cast<int *>(minc + r.bi), __float_as_int(inc));  // inc > 0
                    }
                }
            }
            // ---- ticket: the last CTA of the group runs the O(n) tail of this iteration ----
            __threadfence();
            __syncthreads();
            if (tid == 0) s_last = (a.group == 1) || (atomicAdd(ticket, 1) == (it + 1) * a.group - 1);
            __syncthreads();
            if (s_last) {
                __threadfence();
                // GetMax (:181-194): highest j inside the +-1e-6 window wins
                for (int u = tid; u < U; u += EMD_THREADS) {
                    const int j = __ldcg(uidx + u);
                    const int bid_id = __ldcg(bd + j);
                    const float bid_inc = __ldcg(binc + j);
                    const float max_inc = __ldcg(minc + bid_id);
                    if ((double)bid_inc - 1e-6 <= (double)max_inc && (double)max_inc <= (double)bid_inc + 1e-6)
                        atomicMax(midx + bid_id, j);
                }
                __syncthreads();
                // Assign (:196-215)
                for (int u = tid; u < U; u += EMD_THREADS) {
                    const int j = __ldcg(uidx + u);
                    const int bid_id = __ldcg(bd + j);
                    if (last) {
                        asg[j] = bid_id;
                        atomicMax(asg_inv + bid_id, j);
                        atomicAdd(pr + bid_id, __ldcg(binc + j));
                        minc[bid_id] = -1e9f;
                    } else if (__ldcg(midx + bid_id) == j) {
                        const int ass_inv = __ldcg(asg_inv + bid_id);
                        if (ass_inv != -1) asg[ass_inv] = -1;
                        asg_inv[bid_id] = j;
                        asg[j] = bid_id;
                        pr[bid_id] = __fadd_rn(__ldcg(pr + bid_id), __ldcg(binc + j));
                        minc[bid_id] = -1e9f;
                    }
                }
                __syncthreads();
                const int U2 = compact_unassigned(asg, uidx, midx, n, sscan);
                __threadfence();
                __syncthreads();
                if (tid == 0) {
                    a.unass_cnt[b] = U2;
                    __threadfence();
                    st_release(flag, it + 2);
                }
            }
        }
        // ---- CalcDist (:217-226), split over the group ----
        if (tid == 0 && !complete) {
            while (ld_acquire(flag) < a.iters + 1) __nanosleep(64);
        }
        __syncthreads();
        for (int j = rank * EMD_THREADS + tid; j < n; j += a.group * EMD_THREADS) {
            const int k = __ldcg(asg + j);
            const float dx = __fsub_rn(__ldg(p1 + j * 3), __ldg(p2 + k * 3));
            const float dy = __fsub_rn(__ldg(p1 + j * 3 + 1), __ldg(p2 + k * 3 + 1));
            const float dz = __fsub_rn(__ldg(p1 + j * 3 + 2), __ldg(p2 + k * 3 + 2));
            a.dist[(size_t)b * n + j] = __fmaf_rn(dz, dz, __fmaf_rn(dx, dx, __fmul_rn(dy, dy)));
        }
        __syncthreads();
    }
}

// NmDistanceGradKernel of emd_cuda.cu:284-300: gradient to xyz1 only, ACCUMULATED (atomicAdd in the reference,
// one term per element so a plain read-add-write is identical).
__global__ void emd_grad_kernel(const float *__restrict__ xyz1, const float *__restrict__ xyz2,
                                const float *__restrict__ graddist, const int *__restrict__ idx, float *gradxyz, int B,
                                int n) {
    const size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= (size_t)B * n) return;
    const size_t b = i / n;
    const size_t t = b * n + __ldg(idx + i);
    const float g = __fmul_rn(__ldg(graddist + i), 2.f);
#pragma unroll
    for (int c = 0; c < 3; ++c)
        gradxyz[i * 3 + c] = __fadd_rn(gradxyz[i * 3 + c], __fmul_rn(g, __fsub_rn(__ldg(xyz1 + i * 3 + c), __ldg(xyz2 + t * 3 + c))));
}

}  // namespace genpc

using namespace genpc;

extern "C" size_t genpc_emd_workspace_bytes(int B) { return B < 0 ? 0 : (size_t)B * 2 * sizeof(int); }

extern "C" int genpc_emd_forward(const float *xyz1, const float *xyz2, float *dist, int *assignment, float *price,
                                 int *assignment_inv, int *bid, float *bid_increments, float *max_increments,
                                 int *unass_idx, int *unass_cnt, int *max_idx, int B, int n, int m, float eps, int iters,
                                 void *workspace, size_t workspace_bytes, genpc_stream_t stream_) {
    cudaStream_t stream = (cudaStream_t)stream_;
    // the reference's checks (emd_cuda.cu:236-249)
    if (n != m || B > 512 || n % 256 != 0 || B < 0 || n < 0 || iters < 0) return GENPC_ERR_SHAPE;
    if (B == 0 || n == 0) return GENPC_OK;
    if (workspace == nullptr || workspace_bytes < genpc_emd_workspace_bytes(B)) return GENPC_ERR_WORKSPACE;
    cudaError_t e = cudaMemsetAsync(workspace, 0, genpc_emd_workspace_bytes(B), stream);
    if (e != cudaSuccess) return (int)e;
    int dev = 0, sms = 0, per_sm = 0;
    cudaGetDevice(&dev);
    cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
    e = cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, emd_auction_kernel, EMD_THREADS, 0);
    if (e != cudaSuccess) return (int)e;
    const int resident = sms * per_sm;
    if (resident <= 0) return (int)cudaErrorLaunchOutOfResources;
    EmdArgs a;
    a.xyz1 = xyz1, a.xyz2 = xyz2, a.dist = dist, a.assignment = assignment, a.price = price;
    a.assignment_inv = assignment_inv, a.bid = bid, a.bid_increments = bid_increments;
    a.max_increments = max_increments, a.unass_idx = unass_idx, a.unass_cnt = unass_cnt, a.max_idx = max_idx;
    a.flags = (int *)workspace, a.tickets = (int *)workspace + B;
    a.B = B, a.n = n, a.eps = eps, a.iters = iters;
    int grid;
    if (B >= resident) {
        a.group = 1;
        grid = resident;
    } else {
        a.group = resident / B;
        // more CTAs than 256-point items in the first (all-unassigned) iteration is wasted spinning
        const int useful = (n + 15) / 16;
        if (a.group > useful) a.group = useful;
        if (a.group < 1) a.group = 1;
        grid = a.group * B;
    }
    void *kargs[] = {(void *)&a};
    e = cudaLaunchCooperativeKernel((void *)emd_auction_kernel, dim3(grid), dim3(EMD_THREADS), kargs, 0, stream);
    if (e != cudaSuccess) return (int)e;
    return GENPC_OK;
}

extern "C" int genpc_emd_backward(const float *xyz1, const float *xyz2, float *gradxyz, const float *graddist,
                                  const int *idx, int B, int n, genpc_stream_t stream_) {
    cudaStream_t stream = (cudaStream_t)stream_;
    if (B < 0 || n < 0) return GENPC_ERR_SHAPE;
    const size_t tot = (size_t)B * n;
    if (tot == 0) return GENPC_OK;
    emd_grad_kernel<<<(unsigned)((tot + 255) / 256), 256, 0, stream>>>(xyz1, xyz2, graddist, idx, gradxyz, B, n);
    GENPC_CHECK_LAUNCH();
    return GENPC_OK;
}
