// fps.cu -- farthest point sampling for sm_100a: one persistent CTA per cloud.
//
// Replaces the reference's CPU call `fpsample.fps_sampling(xyz, K)` (main.py:21-22, reg_xyz.py:215,
// DepthPrompting.py:88-90; un-vendored Rust package -> semantics defined by oracle_fps):
// start index given, running distance +inf, d = fma(dz,dz,fma(dx,dx,dy*dy)), next = arg-max of the
// running distance with the LOWEST index on ties.
//
// The K picks are strictly sequential, so the kernel is a latency machine: the cloud lives in registers
// (PPT points + running distances per thread, 1024 threads), each pick costs one distance update per
// point, two REDUX warp reductions (max of the value bits, then min index among the lanes that hold it),
// ONE __syncthreads (double-buffered 32-entry exchange), and the same two REDUX again.
// Clouds larger than 1024*PPT_MAX points fall back to a global-memory loop with the same arithmetic.
#include <cooperative_groups.h>

#include <cstdlib>
#include <cstring>

#include "common.cuh"

namespace genpc {

constexpr int FPS_THREADS = 1024;
constexpr int FPS_WARPS = FPS_THREADS / 32;

__device__ __forceinline__ void fps_block_argmax(float best, int best_i, unsigned (*sval)[FPS_WARPS],
                                                 int (*sidx)[FPS_WARPS], int parity, int lane, int warp,
                                                 int &winner, unsigned *bmax_out = nullptr) {
    // running distances are >= 0 (or -1 for padding, mapped to 0 bits below) -> bit pattern orders like the float
    unsigned vb = best < 0.f ? 0u : __float_as_uint(best) + 1u;  // +1 keeps real 0.0 above the padding
    unsigned wmax = __reduce_max_sync(0xffffffffu, vb);
    int cand = (vb == wmax) ? best_i : 0x7fffffff;
    int wmin = __reduce_min_sync(0xffffffffu, cand);
    if (lane == 0) {
        sval[parity][warp] = wmax;
        sidx[parity][warp] = wmin;
    }
    __syncthreads();
    unsigned v2 = sval[parity][lane];
    int i2 = sidx[parity][lane];
    unsigned bmax = __reduce_max_sync(0xffffffffu, v2);
    int c2 = (v2 == bmax) ? i2 : 0x7fffffff;
    winner = __reduce_min_sync(0xffffffffu, c2);
    if (bmax_out != nullptr) *bmax_out = bmax;
}

template <int PPT, bool CREG>
__global__ void __launch_bounds__(FPS_THREADS, 1) fps_reg_kernel(const float *__restrict__ xyz, int N, int K,
                                                                   int start, int *__restrict__ idx_out,
                                                                   float *__restrict__ seq_out) {
    __shared__ unsigned sval[2][FPS_WARPS];
    __shared__ int sidx[2][FPS_WARPS];
    const int b = blockIdx.x, tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const float *p = xyz + (size_t)b * N * 3;
    // running distances always live in registers; coordinates too when the cloud is small (CREG),
    // otherwise they are re-read through L1 (196 KB for 16384 points fits the 228 KB L1 of one SM).
    constexpr int CP = CREG ? PPT : 1;
    float px[CP], py[CP], pz[CP], run[PPT];
#pragma unroll
    for (int k = 0; k < PPT; ++k) {
        const int i = tid + k * FPS_THREADS;
        if (CREG) px[k % CP] = py[k % CP] = pz[k % CP] = 0.f;
        if (i < N) {
            if (CREG) px[k % CP] = __ldg(p + i * 3), py[k % CP] = __ldg(p + i * 3 + 1), pz[k % CP] = __ldg(p + i * 3 + 2);
            run[k] = __int_as_float(0x7f800000);
        } else {
            run[k] = -1.f;  // padding: never selected (distances are >= 0)
        }
    }
    int cur = start;
    for (int s = 0; s < K; ++s) {
        if (tid == 0) idx_out[(size_t)b * K + s] = cur;
        const float lx = __ldg(p + cur * 3), ly = __ldg(p + cur * 3 + 1), lz = __ldg(p + cur * 3 + 2);
        float best = -2.f;
        int best_i = 0x7fffffff;
#pragma unroll
        for (int k = 0; k < PPT; ++k) {
            float x, y, z;
            if (CREG) {
                x = px[k % CP], y = py[k % CP], z = pz[k % CP];
            } else {
                const int i = min(tid + k * FPS_THREADS, N - 1);
                x = __ldg(p + i * 3), y = __ldg(p + i * 3 + 1), z = __ldg(p + i * 3 + 2);
            }
            const float dx = __fsub_rn(x, lx), dy = __fsub_rn(y, ly), dz = __fsub_rn(z, lz);
            const float d = __fmaf_rn(dz, dz, __fmaf_rn(dx, dx, __fmul_rn(dy, dy)));
            const float r = (run[k] < d) ? run[k] : d;  // as oracle_fps; padding (-1) stays -1
            run[k] = r;
            if (r > best) {  // strict: k ascending == index ascending -> lowest index kept
                best = r;
                best_i = tid + k * FPS_THREADS;
            }
        }
        int winner;
        fps_block_argmax(best, best_i, sval, sidx, s & 1, lane, warp, winner);
        if (seq_out != nullptr && tid == 0) {
            if (s == 0) seq_out[(size_t)b * K] = __int_as_float(0x7f800000);
            if (s + 1 < K) {
                // the winner's running distance: max value bits - 1
                unsigned m = 0;
#pragma unroll
                for (int w = 0; w < FPS_WARPS; ++w) m = max(m, sval[s & 1][w]);
                seq_out[(size_t)b * K + s + 1] = __uint_as_float(m - 1u);
            }
        }
        cur = winner;
    }
}

// Shared-memory form of the single-CTA kernel for 4096 < N <= 16384: the cloud is staged ONCE as SoA x[] / y[] / z[] in
// the CTA's dynamic shared memory (12 B per point, 196 KB at N = 16384) and a thread owns PPT/4 groups of FOUR consecutive
// points (indices k*4096 + 4*tid + j), so a pick reads its coordinates with three conflict-free LDS.128 per four points
// instead of twelve LDG through L1 with their address arithmetic (fps_reg_kernel<.., false>): about a third fewer
// instructions per pick in the kernel that serves batched FPS (one CTA per cloud).  Same arithmetic, same tie rule
// (ascending (k, j) == ascending index inside a thread; the block arg-max takes the lowest index among equal values).
template <int PPT>
__global__ void __launch_bounds__(FPS_THREADS, 1) fps_smem_kernel(const float *__restrict__ xyz, int N, int K, int start,
                                                                    int *__restrict__ idx_out, float *__restrict__ seq_out) {
    static_assert(PPT % 4 == 0, "a thread owns groups of four consecutive points");
    extern __shared__ __align__(16) float dyn_xyz[];
    constexpr int NP = PPT * FPS_THREADS;  // padded cloud size
    float *sx = dyn_xyz, *sy = dyn_xyz + NP, *sz = dyn_xyz + 2 * NP;
    __shared__ unsigned sval[2][FPS_WARPS];
    __shared__ int sidx[2][FPS_WARPS];
    const int b = blockIdx.x, tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const float *p = xyz + (size_t)b * N * 3;
    for (int i = tid; i < NP; i += FPS_THREADS) {
        float x = 0.f, y = 0.f, z = 0.f;
        if (i < N) x = __ldg(p + (size_t)i * 3), y = __ldg(p + (size_t)i * 3 + 1), z = __ldg(p + (size_t)i * 3 + 2);
        sx[i] = x, sy[i] = y, sz[i] = z;
    }
    float run[PPT];
#pragma unroll
    for (int k = 0; k < PPT / 4; ++k)
#pragma unroll
        for (int j = 0; j < 4; ++j) run[k * 4 + j] = (k * 4 * FPS_THREADS + 4 * tid + j < N) ? __int_as_float(0x7f800000) : -1.f;
    __syncthreads();
    const float4 *sx4 = reinterpret_cast<const float4 *>(sx), *sy4 = reinterpret_cast<const float4 *>(sy);
    const float4 *sz4 = reinterpret_cast<const float4 *>(sz);
    int cur = start;
    for (int s = 0; s < K; ++s) {
        if (tid == 0) idx_out[(size_t)b * K + s] = cur;
        const float lx = sx[cur], ly = sy[cur], lz = sz[cur];
        float best = -2.f;
        int best_i = 0x7fffffff;
#pragma unroll
        for (int k = 0; k < PPT / 4; ++k) {
            const float4 X = sx4[k * FPS_THREADS + tid], Y = sy4[k * FPS_THREADS + tid], Z = sz4[k * FPS_THREADS + tid];
            const float xs[4] = {X.x, X.y, X.z, X.w}, ys[4] = {Y.x, Y.y, Y.z, Y.w}, zs[4] = {Z.x, Z.y, Z.z, Z.w};
#pragma unroll
            for (int j = 0; j < 4; ++j) {
                const float dx = __fsub_rn(xs[j], lx), dy = __fsub_rn(ys[j], ly), dz = __fsub_rn(zs[j], lz);
                const float d = __fmaf_rn(dz, dz, __fmaf_rn(dx, dx, __fmul_rn(dy, dy)));
                const float r = (run[k * 4 + j] < d) ? run[k * 4 + j] : d;  // as oracle_fps; padding (-1) stays -1
                run[k * 4 + j] = r;
                if (r > best) {  // strict: (k, j) ascending == index ascending -> lowest index kept
                    best = r;
                    best_i = k * 4 * FPS_THREADS + 4 * tid + j;
                }
            }
        }
        int winner;
        fps_block_argmax(best, best_i, sval, sidx, s & 1, lane, warp, winner);
        if (seq_out != nullptr && tid == 0) {
            if (s == 0) seq_out[(size_t)b * K] = __int_as_float(0x7f800000);
            if (s + 1 < K) {
                unsigned m = 0;
#pragma unroll
                for (int w = 0; w < FPS_WARPS; ++w) m = max(m, sval[s & 1][w]);
                seq_out[(size_t)b * K + s + 1] = __uint_as_float(m - 1u);
            }
        }
        cur = winner;
    }
}

template <int PPT>
static cudaError_t launch_fps_smem(const float *xyz, int B, int N, int K, int start, int *idx_out, float *seq_out,
                                   cudaStream_t stream) {
    const size_t bytes = (size_t)3 * PPT * FPS_THREADS * sizeof(float);
    cudaError_t e = cudaFuncSetAttribute(fps_smem_kernel<PPT>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)bytes);
    if (e != cudaSuccess) return e;
    fps_smem_kernel<PPT><<<B, FPS_THREADS, bytes, stream>>>(xyz, N, K, start, idx_out, seq_out);
    return cudaGetLastError();
}

// Generic fallback: points and running distances stay in global memory (L2-resident for any realistic N).
__global__ void __launch_bounds__(FPS_THREADS, 1) fps_gmem_kernel(const float *__restrict__ xyz, int N, int K,
                                                                    int start, int *__restrict__ idx_out,
                                                                    float *__restrict__ seq_out,
                                                                    float *__restrict__ run_ws) {
    __shared__ unsigned sval[2][FPS_WARPS];
    __shared__ int sidx[2][FPS_WARPS];
    const int b = blockIdx.x, tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const float *p = xyz + (size_t)b * N * 3;
    float *run = run_ws + (size_t)b * N;
    for (int i = tid; i < N; i += FPS_THREADS) run[i] = __int_as_float(0x7f800000);
    int cur = start;
    for (int s = 0; s < K; ++s) {
        if (tid == 0) idx_out[(size_t)b * K + s] = cur;
        const float lx = __ldg(p + cur * 3), ly = __ldg(p + cur * 3 + 1), lz = __ldg(p + cur * 3 + 2);
        float best = -2.f;
        int best_i = 0x7fffffff;
        for (int i = tid; i < N; i += FPS_THREADS) {
            const float dx = __fsub_rn(__ldg(p + i * 3), lx), dy = __fsub_rn(__ldg(p + i * 3 + 1), ly),
                        dz = __fsub_rn(__ldg(p + i * 3 + 2), lz);
            const float d = __fmaf_rn(dz, dz, __fmaf_rn(dx, dx, __fmul_rn(dy, dy)));
            const float rv = run[i];
            const float r = (rv < d) ? rv : d;
            run[i] = r;
            if (r > best) {
                best = r;
                best_i = i;
            }
        }
        int winner;
        fps_block_argmax(best, best_i, sval, sidx, s & 1, lane, warp, winner);
        if (seq_out != nullptr && tid == 0) {
            if (s == 0) seq_out[(size_t)b * K] = __int_as_float(0x7f800000);
            if (s + 1 < K) {
                unsigned m = 0;
                for (int w = 0; w < FPS_WARPS; ++w) m = max(m, sval[s & 1][w]);
                seq_out[(size_t)b * K + s + 1] = __uint_as_float(m - 1u);
            }
        }
        cur = winner;
    }
}

// ---- thread-block-cluster version: the cloud lives in the REGISTERS of CS CTAs (CS SMs) ----------------------
// Large single clouds are bound by one SM's L1/LSU when a lone CTA re-reads 12 B/point every pick (16384 points:
// 3.6 ms for 2048 picks).  A cluster of CS CTAs keeps PPT <= 4 points per thread entirely in registers; per pick
// each CTA finds its own candidate (REDUX + one __syncthreads), the owning thread posts (value, index, x, y, z) into
// slot [parity][rank] of EVERY CTA of the cluster with asynchronous distributed-shared-memory stores that complete
// transaction bytes on the receiver's mbarrier (st.async ... mbarrier::complete_tx::bytes); one thread per CTA waits
// on the local mbarrier.  (Measured per pick: barrier.cluster 1.41 us, st.shared::cluster + release/acquire arrive
// 1.45 us, st.async + complete_tx: see profiles/.)  Every thread then reads the CS candidates locally -- the winner's coordinates travel with it, so the loop has no global
// loads at all.  Same arithmetic and tie rule as fps_reg_kernel (bit-identical results).
namespace cg = cooperative_groups;

__device__ __forceinline__ unsigned smem_u32(const void *p) { return (unsigned)__cvta_generic_to_shared(p); }
__device__ __forceinline__ unsigned mapa_u32(unsigned addr, unsigned rank) {
    unsigned r;
    asm volatile("mapa.shared::cluster.u32 %0, %1, %2;" : "=r"(r) : "r"(addr), "r"(rank));
    return r;
}
__device__ __forceinline__ void st_cluster_v4(unsigned addr, int4 v) {
    asm volatile("st.shared::cluster.v4.b32 [%0], {%1, %2, %3, %4};" ::"r"(addr), "r"(v.x), "r"(v.y), "r"(v.z), "r"(v.w) : "memory");
}
__device__ __forceinline__ void mbar_init(unsigned addr, unsigned count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(addr), "r"(count) : "memory");
}
__device__ __forceinline__ void mbar_arrive_remote(unsigned cluster_addr) {
    asm volatile("mbarrier.arrive.release.cluster.shared::cluster.b64 _, [%0];" ::"r"(cluster_addr) : "memory");
}
// asynchronous DSMEM store that signals the receiver's mbarrier by transaction bytes (no release fence on the sender,
// no cluster-scope acquire on the receiver: visibility comes with the phase completion)
__device__ __forceinline__ void st_async_v4(unsigned cluster_addr, int4 v, unsigned cluster_mbar) {
    asm volatile("st.async.weak.shared::cluster.mbarrier::complete_tx::bytes.v4.b32 [%0], {%1, %2, %3, %4}, [%5];" ::"r"(cluster_addr),
                 "r"(v.x), "r"(v.y), "r"(v.z), "r"(v.w), "r"(cluster_mbar)
                 : "memory");
}
__device__ __forceinline__ void mbar_arm(unsigned addr, unsigned tx_bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(addr), "r"(tx_bytes) : "memory");
}
__device__ __forceinline__ void mbar_wait(unsigned addr, unsigned parity) {
    asm volatile(
        "{\n"
        ".reg .pred P1;\n"
        "LAB_WAIT:\n"
        "mbarrier.try_wait.parity.shared::cta.b64 P1, [%0], %1;\n"
        "@P1 bra DONE;\n"
        "bra LAB_WAIT;\n"
        "DONE:\n"
        "}\n" ::"r"(addr), "r"(parity) : "memory");
}
// post one 32-byte candidate into slot [par][rank] of every CTA of the cluster, then release-arrive on each receiver
template <int CS>
__device__ __forceinline__ void fps_post(const void *cand_slot, const void *bar, const int4 lo, const int4 hi) {
    const unsigned slot = smem_u32(cand_slot), b = smem_u32(bar);
#pragma unroll
    for (int r = 0; r < CS; ++r) {
        const unsigned dst = mapa_u32(slot, r);
        st_cluster_v4(dst, lo);
        st_cluster_v4(dst + 16, hi);
        mbar_arrive_remote(mapa_u32(b, r));
    }
}

struct FpsCand {
    unsigned val;  // running distance bits + 1 (0 = padding)
    int idx;
    float x, y, z;
    int pad[3];
};

// SMEMC = false: coordinates in registers (PPT <= 4, N <= 32768).  SMEMC = true: coordinates in the CTA's dynamic
// shared memory as SoA (3 * PPT * 1024 floats, up to 221 KB at PPT = 18 -> clouds of up to 147 456 points, the sizes
// the reference feeds to fpsample: 45 K - 170 K), running distances still in registers.
template <int PPT, int CS, bool SMEMC>
__global__ void __launch_bounds__(FPS_THREADS, 1) fps_cluster_kernel(const float *__restrict__ xyz, int N, int K, int start,
                                                                      int *__restrict__ idx_out, float *__restrict__ seq_out) {
    extern __shared__ __align__(16) float dyn_xyz[];
    float *sx = dyn_xyz, *sy = dyn_xyz + PPT * FPS_THREADS, *sz = dyn_xyz + 2 * PPT * FPS_THREADS;
    __shared__ unsigned sval[2][FPS_WARPS];
    __shared__ int sidx[2][FPS_WARPS];
    __shared__ __align__(16) FpsCand cand[2][CS];
    __shared__ __align__(8) unsigned long long mbar[2];  // one mbarrier per parity, CS remote arrivals per pick
    cg::cluster_group cluster = cg::this_cluster();
    const int rank = (int)cluster.block_rank();
    const int b = blockIdx.x / CS, tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const float *p = xyz + (size_t)b * N * 3;
    // point i of the cloud lives in CTA (i / 1024) % CS ... interleaved so that every CTA holds a similar share:
    // global index of (rank, k, tid) = (k * CS + rank) * 1024 + tid
    constexpr int RP = SMEMC ? 1 : PPT;
    float px[RP], py[RP], pz[RP], run[PPT];
#pragma unroll
    for (int k = 0; k < PPT; ++k) {
        const int i = (k * CS + rank) * FPS_THREADS + tid;
        float x = 0.f, y = 0.f, z = 0.f;
        if (i < N) {
            x = __ldg(p + (size_t)i * 3), y = __ldg(p + (size_t)i * 3 + 1), z = __ldg(p + (size_t)i * 3 + 2);
            run[k] = __int_as_float(0x7f800000);
        } else {
            run[k] = -1.f;
        }
        if (SMEMC) sx[k * FPS_THREADS + tid] = x, sy[k * FPS_THREADS + tid] = y, sz[k * FPS_THREADS + tid] = z;
        else px[k % RP] = x, py[k % RP] = y, pz[k % RP] = z;
    }
    float lx = __ldg(p + (size_t)start * 3), ly = __ldg(p + (size_t)start * 3 + 1), lz = __ldg(p + (size_t)start * 3 + 2);
    int cur = start;
    if (tid == 0) {
        mbar_init(smem_u32(&mbar[0]), 1);
        mbar_init(smem_u32(&mbar[1]), 1);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
        mbar_arm(smem_u32(&mbar[0]), CS * 32);  // each pick delivers CS candidates of 32 bytes into this CTA
        mbar_arm(smem_u32(&mbar[1]), CS * 32);
    }
    cluster.sync();
    for (int s = 0; s < K; ++s) {
        if (rank == 0 && tid == 0) idx_out[(size_t)b * K + s] = cur;
        float best = -2.f;
        int best_i = 0x7fffffff, best_k = 0;
#pragma unroll
        for (int k = 0; k < PPT; ++k) {
            const float qx = SMEMC ? sx[k * FPS_THREADS + tid] : px[k % RP];
            const float qy = SMEMC ? sy[k * FPS_THREADS + tid] : py[k % RP];
            const float qz = SMEMC ? sz[k * FPS_THREADS + tid] : pz[k % RP];
            const float dx = __fsub_rn(qx, lx), dy = __fsub_rn(qy, ly), dz = __fsub_rn(qz, lz);
            const float d = __fmaf_rn(dz, dz, __fmaf_rn(dx, dx, __fmul_rn(dy, dy)));
            const float r = (run[k] < d) ? run[k] : d;
            run[k] = r;
            if (r > best) {  // k ascending == index ascending inside a thread
                best = r;
                best_i = (k * CS + rank) * FPS_THREADS + tid;
                best_k = k;
            }
        }
        const int par = s & 1;
        int winner;
        unsigned bmax;
        fps_block_argmax(best, best_i, sval, sidx, par, lane, warp, winner, &bmax);
        // the warp that owns the CTA's candidate posts it: the owner lane's coordinates are shuffled to lanes 0..CS-1,
        // which write to one destination CTA each (a single thread posting to all CS CTAs serialises CS release-arrives,
        // each waiting for its own DSMEM stores: measured 2.5 us per pick)
        const int wtid = (winner == 0x7fffffff) ? 0 : (winner & (FPS_THREADS - 1));
        if (warp == (wtid >> 5)) {
            float cx, cy, cz;
            if (SMEMC) {
                cx = sx[best_k * FPS_THREADS + tid], cy = sy[best_k * FPS_THREADS + tid], cz = sz[best_k * FPS_THREADS + tid];
            } else {
                cx = px[0], cy = py[0], cz = pz[0];
#pragma unroll
                for (int k = 1; k < RP; ++k)
                    if (best_k == k) cx = px[k], cy = py[k], cz = pz[k];
            }
            const int src = wtid & 31;
            cx = __shfl_sync(0xffffffffu, cx, src), cy = __shfl_sync(0xffffffffu, cy, src), cz = __shfl_sync(0xffffffffu, cz, src);
            if (lane < CS) {
                const int4 lo = make_int4((int)bmax, winner, __float_as_int(cx), __float_as_int(cy));
                const int4 hi = make_int4(__float_as_int(cz), 0, 0, 0);
                const unsigned dst = mapa_u32(smem_u32(&cand[par][rank]), lane);
                const unsigned bar = mapa_u32(smem_u32(&mbar[par]), lane);
                st_async_v4(dst, lo, bar);
                st_async_v4(dst + 16, hi, bar);
            }
        }
        if (tid == 0) {
            mbar_wait(smem_u32(&mbar[par]), (unsigned)((s >> 1) & 1));  // all CS candidates of this pick have landed
            mbar_arm(smem_u32(&mbar[par]), CS * 32);                      // re-arm for pick s+2 (nobody can post it yet)
        }
        __syncthreads();
        unsigned bv = 0;
        int bi = 0x7fffffff;
#pragma unroll
        for (int r = 0; r < CS; ++r) {
            const FpsCand c = cand[par][r];
            if (c.val > bv || (c.val == bv && c.idx < bi)) {
                bv = c.val, bi = c.idx;
                lx = c.x, ly = c.y, lz = c.z;
            }
        }
        if (seq_out != nullptr && rank == 0 && tid == 0) {
            if (s == 0) seq_out[(size_t)b * K] = __int_as_float(0x7f800000);
            if (s + 1 < K) seq_out[(size_t)b * K + s + 1] = __uint_as_float(bv - 1u);
        }
        cur = bi;
    }
    cluster.sync();  // no CTA may exit while a sibling can still write into its shared memory
}

// A cluster needs CS SMs of ONE GPC: the number of clusters that can be resident together is what the occupancy query
// says, not SMs / CS (B200: 16 clusters of 8 CTAs, although 148 / 8 = 18).  More clouds than that run in waves -- the
// single-CTA kernel is the better choice then.  Returns 0 when the query fails.
template <int PPT, int CS, bool SMEMC>
static int fps_max_active_clusters() {
    static int cache[64];     // per template instance AND per device ordinal (0 = not asked yet, -1 = query failed)
    int dev = 0;
    if (cudaGetDevice(&dev) != cudaSuccess || dev < 0 || dev >= 64) dev = 0;
    int &cached = cache[dev];
    if (cached != 0) return cached < 0 ? 0 : cached;
    cudaLaunchConfig_t cfg = {};
    cfg.gridDim = dim3((unsigned)(CS * 64));
    cfg.blockDim = dim3(FPS_THREADS);
    cfg.dynamicSmemBytes = SMEMC ? (size_t)3 * PPT * FPS_THREADS * sizeof(float) : 0;
    if (SMEMC && cudaFuncSetAttribute(fps_cluster_kernel<PPT, CS, SMEMC>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                      (int)cfg.dynamicSmemBytes) != cudaSuccess) {
        (void)cudaGetLastError();
        cached = -1;
        return 0;
    }
    if (CS > 8 && cudaFuncSetAttribute(fps_cluster_kernel<PPT, CS, SMEMC>, cudaFuncAttributeNonPortableClusterSizeAllowed, 1) !=
                      cudaSuccess) {
        (void)cudaGetLastError();
        return 0;
    }
    cudaLaunchAttribute attr[1];
    attr[0].id = cudaLaunchAttributeClusterDimension;
    attr[0].val.clusterDim.x = CS, attr[0].val.clusterDim.y = 1, attr[0].val.clusterDim.z = 1;
    cfg.attrs = attr, cfg.numAttrs = 1;
    int n = 0;
    if (cudaOccupancyMaxActiveClusters(&n, fps_cluster_kernel<PPT, CS, SMEMC>, &cfg) != cudaSuccess) {
        (void)cudaGetLastError();
        n = 0;
    }
    cached = n > 0 ? n : -1;
    return n;
}

template <int PPT, int CS, bool SMEMC>
static cudaError_t launch_fps_cluster(const float *xyz, int B, int N, int K, int start, int *idx_out, float *seq_out,
                                      cudaStream_t stream) {
    // clouds of up to 16384 points have a fast single-CTA kernel (2.9 ms for 16384 -> 2048 whatever B): when the clusters
    // would run in waves (4.7 ms at B = 18) the caller falls through to it; larger clouds stay here (waves of clusters
    // still beat the L1 / global-memory single-CTA forms)
    if (PPT * CS <= 16 && tunable("GENPC_FPS_MODE") == nullptr && B > fps_max_active_clusters<PPT, CS, SMEMC>())
        return cudaErrorLaunchOutOfResources;
    cudaLaunchConfig_t cfg = {};
    cfg.gridDim = dim3((unsigned)(B * CS));
    cfg.blockDim = dim3(FPS_THREADS);
    cfg.dynamicSmemBytes = SMEMC ? (size_t)3 * PPT * FPS_THREADS * sizeof(float) : 0;
    if (SMEMC) {
        cudaError_t e = cudaFuncSetAttribute(fps_cluster_kernel<PPT, CS, SMEMC>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                             (int)cfg.dynamicSmemBytes);
        if (e != cudaSuccess) return e;
    }
    if (CS > 8) {  // 16-CTA clusters are beyond the portable size
        cudaError_t e = cudaFuncSetAttribute(fps_cluster_kernel<PPT, CS, SMEMC>, cudaFuncAttributeNonPortableClusterSizeAllowed, 1);
        if (e != cudaSuccess) return e;
    }
    cfg.stream = stream;
    cudaLaunchAttribute attr[1];
    attr[0].id = cudaLaunchAttributeClusterDimension;
    attr[0].val.clusterDim.x = CS, attr[0].val.clusterDim.y = 1, attr[0].val.clusterDim.z = 1;
    cfg.attrs = attr, cfg.numAttrs = 1;
    return cudaLaunchKernelEx(&cfg, fps_cluster_kernel<PPT, CS, SMEMC>, xyz, N, K, start, idx_out, seq_out);
}

}  // namespace genpc

using namespace genpc;

extern "C" size_t genpc_fps_workspace_bytes(int B, int N, int K) {
    if (B < 0 || N < 0) return 0;
    return (N > FPS_THREADS * 32) ? (size_t)B * N * sizeof(float) : 0;
}

extern "C" int genpc_fps(const float *xyz, int B, int N, int K, int start, int *idx_out, float *seq_out,
                         void *workspace, size_t workspace_bytes, genpc_stream_t stream_) {
    cudaStream_t stream = (cudaStream_t)stream_;
    if (B < 0 || N <= 0 || K < 0 || K > N || start < 0 || start >= N) return GENPC_ERR_SHAPE;
    if (B == 0 || K == 0) return GENPC_OK;
    const int ppt = (N + FPS_THREADS - 1) / FPS_THREADS;
    // clusters of 8 CTAs when the batch alone cannot fill the chip and the cloud is big enough to be LSU/L1 bound.
    // Measured on B200 (profiles/r01d_fps.txt): a pick costs 1.15 us through the cluster exchange whatever N is, and
    // 0.35 / 0.51 / 1.02 / 1.78 / 4.23 us in the single CTA at N = 1024 / 4096 / 8192 / 16384 / 32768.
    const char *fm = tunable("GENPC_FPS_MODE");
    const bool want_cluster = (fm == nullptr) ? (ppt > 8 && ppt <= 144 && B * 8 <= GENPC_NUM_SMS) : (strcmp(fm, "cluster") == 0 && ppt <= 144);
    if (want_cluster) {
        cudaError_t e = cudaErrorUnknown;
        // big clouds (N > 49152, the 45 K - 170 K scans the reference feeds to fpsample): a 16-CTA cluster holds the whole
        // cloud in registers (<= 9 points per thread) and halves the per-pick distance update of the 8-CTA form, whose
        // coordinates live in shared memory; the exchange grows from 8 to 16 candidates.  Falls back to 8 CTAs when the
        // device cannot co-schedule 16 (profiles/r01j_fps_cluster16.txt).
        const char *c16 = tunable("GENPC_FPS_CLUSTER16");
        // measured: 16 CTAs win from ~45 K points on (1.73 vs 2.06 us per pick at 71 372, 2.11 vs 2.56 at 139 138)
        const bool try16 = (c16 == nullptr) ? (ppt > 48 && B * 16 <= GENPC_NUM_SMS) : (atoi(c16) != 0 && ppt > 8);
        if (try16) {
            if (ppt <= 16) e = launch_fps_cluster<1, 16, false>(xyz, B, N, K, start, idx_out, seq_out, stream);
            else if (ppt <= 32) e = launch_fps_cluster<2, 16, false>(xyz, B, N, K, start, idx_out, seq_out, stream);
            else if (ppt <= 64) e = launch_fps_cluster<4, 16, false>(xyz, B, N, K, start, idx_out, seq_out, stream);
            else if (ppt <= 96) e = launch_fps_cluster<6, 16, false>(xyz, B, N, K, start, idx_out, seq_out, stream);
            else e = launch_fps_cluster<9, 16, false>(xyz, B, N, K, start, idx_out, seq_out, stream);
            if (e == cudaSuccess) return GENPC_OK;
            (void)cudaGetLastError();  // not schedulable here: clear the error and use the portable size
        }
        if (ppt <= 8) e = launch_fps_cluster<1, 8, false>(xyz, B, N, K, start, idx_out, seq_out, stream);
        else if (ppt <= 16) e = launch_fps_cluster<2, 8, false>(xyz, B, N, K, start, idx_out, seq_out, stream);
        else if (ppt <= 32) e = launch_fps_cluster<4, 8, false>(xyz, B, N, K, start, idx_out, seq_out, stream);
        else if (ppt <= 64) e = launch_fps_cluster<8, 8, true>(xyz, B, N, K, start, idx_out, seq_out, stream);
        else if (ppt <= 96) e = launch_fps_cluster<12, 8, true>(xyz, B, N, K, start, idx_out, seq_out, stream);
        else e = launch_fps_cluster<18, 8, true>(xyz, B, N, K, start, idx_out, seq_out, stream);
        if (e == cudaSuccess) return GENPC_OK;
        if (e != cudaErrorLaunchOutOfResources || fm != nullptr) return (int)e;
        (void)cudaGetLastError();  // more clouds than co-resident clusters: one CTA per cloud below
    }
#define FPS_LAUNCH(P) fps_reg_kernel<P, (P <= 4)><<<B, FPS_THREADS, 0, stream>>>(xyz, N, K, start, idx_out, seq_out)
    // 4096 < N <= 16384: coordinates in shared memory (LDS.128) instead of re-reads through L1; GENPC_FPS_SMEM=0 keeps
    // the L1 form (measured: profiles/r01k_fps_smem.txt)
    const char *fs = tunable("GENPC_FPS_SMEM");
    const bool use_smem = (fs == nullptr) ? true : (atoi(fs) != 0);
    if (use_smem && ppt > 4 && ppt <= 16) {
        const cudaError_t e = (ppt <= 8) ? launch_fps_smem<8>(xyz, B, N, K, start, idx_out, seq_out, stream)
                                         : launch_fps_smem<16>(xyz, B, N, K, start, idx_out, seq_out, stream);
        if (e != cudaSuccess) return (int)e;
        return GENPC_OK;
    }
    if (ppt <= 1) FPS_LAUNCH(1);
    else if (ppt <= 2) FPS_LAUNCH(2);
    else if (ppt <= 4) FPS_LAUNCH(4);
    else if (ppt <= 8) FPS_LAUNCH(8);
    else if (ppt <= 16) FPS_LAUNCH(16);
    else if (ppt <= 32) FPS_LAUNCH(32);
    else {
        if (workspace == nullptr || workspace_bytes < genpc_fps_workspace_bytes(B, N, K)) return GENPC_ERR_WORKSPACE;
        fps_gmem_kernel<<<B, FPS_THREADS, 0, stream>>>(xyz, N, K, start, idx_out, seq_out, (float *)workspace);
    }
#undef FPS_LAUNCH
    GENPC_CHECK_LAUNCH();
    return GENPC_OK;
}
