// nn_prune.cuh -- exact nearest neighbours with spatial pruning (r02).  The default Chamfer scan for batches (device-resident or
// host-fed) of >= 2^30 distance evaluations (BASELINE C2) and inside the registration loop from ~20 scans up; the sort kernel also serves the
// EMD auction's pruned Bid.  GENPC_CHAMFER_PRUNE / GENPC_REGISTER_PRUNE = 0 / 1 forbid / force it.
//
// The symmetric scan (nn_sym.cuh) evaluates every (row, column) distance of a cloud pair: at its issue limit that is
// 7.9e12 pairs/s and nothing in its instruction stream is left to remove.  What is left is not to evaluate most pairs:
//   1. nn_bin_sort_kernel<CS> orders every cloud along a Hilbert curve (counting sort over 8^m grid cells in shared memory;
//      one CTA per cloud, or a thread-block cluster of CS CTAs that read each other's histograms through distributed shared
//      memory when one CTA per cloud would leave most SMs idle) into (x, y, z, original index) records and stores the
//      bounding box of every PR_BLOCK consecutive ones;
//   2. nn_prune_kernel: one warp owns PR_GROUP consecutive sorted queries (one per lane -- neighbours in space).  It
//      computes the squared distance between the group's box and every target block's box, visits the blocks nearest
//      first (REDUX picks them) and stops at the first block farther than the group's worst running minimum: no target
//      in it, or in any later block, can improve (or tie) any lane.  A block is fetched with one coalesced read into the
//      warp's shared-memory slice (SoA) and walked there with the packed FP32 distance of the exhaustive kernels: the
//      same bits.  nn_prune_coop_kernel is the form for a direction with few query groups and many target blocks: a CTA
//      per group, the blocks partitioned over its eight warps, running minima shared through shared memory.
// Exactness: the minimum of a set does not depend on the visiting order; the reported index is the LOWEST original index
// among the targets at the minimum, as in the exhaustive kernels (the winner's 8-target chunk is re-evaluated; a lane
// that saw the same minimum in two chunks takes a second pass over the surviving blocks).  The box distance is scaled
// down by 1e-5, far more than the rounding of the reference's arithmetic, and compared with <= so that ties are kept.
// The sort kernel doubles as the range check (NaN / Inf / |x| > 1e15 -> the exhaustive kernel runs instead, selected on
// the device through the same flag as the tensor-core filter).
#pragma once
#include <cooperative_groups.h>

#include "common.cuh"
#include "nn_core.cuh"   // Similarity / apply_similarity (registration: the moving cloud is sorted in its current pose)

#ifndef GENPC_PR_UNROLL
#define GENPC_PR_UNROLL 4   // 8-target chunks per unrolled step of a block walk (sweep 1/2/4/8: profiles/r02w_unroll_sweep.txt)
#endif

namespace genpc {

constexpr int PR_UNROLL = GENPC_PR_UNROLL;
constexpr int PR_BLOCK = 64;
constexpr int PR_GROUP = 32;
constexpr int PR_MAX_N = 32768;       // 512 blocks: the block id fits the 9 low bits of the selection key
constexpr int PR_SORT_THREADS = 1024;
constexpr int PR_MAX_CELLS = 4096;
constexpr int PR_THREADS = 256;

__host__ __device__ inline int pr_nblk(int n) { return (n + PR_BLOCK - 1) / PR_BLOCK; }
__host__ __device__ inline int pr_npad(int n) { return pr_nblk(n) * PR_BLOCK; }

// 3 * BITS-bit Hilbert index of a grid cell (Skilling, "Programming the Hilbert curve", 2004: axes -> transpose -> interleave).
// Consecutive indices are always neighbouring cells, so ANY run of consecutive sorted points is spatially connected; a
// Morton (Z-order) run that crosses a high-level cell boundary joins two far-apart regions in one bounding box, and every
// query group inside that box has to open it (the first r02 form: 59 instead of ~10 block visits per group on the LiDAR scene).
template <int BITS>
__device__ __forceinline__ unsigned hilbert_key(unsigned x, unsigned y, unsigned z) {
    unsigned X[3] = {x, y, z};
    constexpr unsigned M = 1u << (BITS - 1);
#pragma unroll
    for (unsigned Q = M; Q > 1; Q >>= 1) {
        const unsigned P = Q - 1;
#pragma unroll
        for (int i = 0; i < 3; ++i) {
            if (X[i] & Q) {
                X[0] ^= P;
            } else {
                const unsigned t = (X[0] ^ X[i]) & P;
                X[0] ^= t, X[i] ^= t;
            }
        }
    }
    X[1] ^= X[0], X[2] ^= X[1];
    unsigned t = 0;
#pragma unroll
    for (unsigned Q = M; Q > 1; Q >>= 1)
        if (X[2] & Q) t ^= Q - 1;
    X[0] ^= t, X[1] ^= t, X[2] ^= t;
    unsigned key = 0;
#pragma unroll
    for (int j = BITS - 1; j >= 0; --j)
#pragma unroll
        for (int i = 0; i < 3; ++i) key = (key << 1) | ((X[i] >> j) & 1u);
    return key;
}

struct PruneSortParams {
    const float *xyz[2];   // [B][n][3]
    float4 *sorted[2];     // [B][npad]   (x, y, z, original index); NaN records pad the last block
    float4 *boxes[2];      // [B][2][nblk] lower corners, then upper corners
    int n[2];
    int B;
    float limit;
    int *ctl;              // [1] selection flag (0: pruned kernels run), [2] accumulator, [3] ticket -- as nn_tc_precheck_kernel
    int hilbert;           // cell order: Hilbert curve (1) or Z-order (0)
    float *bbx;            // optional [2][B][8]: every cloud's bounding box, for the overlap test below (nullptr: none)
    int accumulate;        // 1: OR the verdict into ctl[1] (a later chunk of the same batch), 0: overwrite it
    int *flag;             // optional: atomicOr the verdict here instead (chunks sorted concurrently on several streams; ctl then
                           // only provides the chunk's own accumulator / ticket words)
    // registration: only one side is sorted per launch (side0 = its index, grid = B), its points taken from cloud b / src_div
    // and moved by the similarity sim[b] while they are read (the same rounding as the exhaustive scan's staging)
    int side0, single_side;
    int mixed;             // cluster launches only: the larger side's clouds take a whole cluster each, the other side's one CTA each
    int src_div[2];
    const Similarity *sim[2];
};

// <<<ctas, 256>>>: all-ones into the packed words when the selection flag asks for the exhaustive kernels AFTER pruned launches
// have already stored exact (dist, index) words (chunked host-fed batches: a later chunk may be the one out of range)
static __global__ void prune_rearm_kernel(unsigned long long *words, size_t n, const int *select) {
    if (*select == 0) return;
    for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (size_t)gridDim.x * blockDim.x) words[i] = ~0ull;
}

// grid = 2 * B * CS CTAs: cloud = blockIdx.x / CS = side * B + b (or the mixed layout below).  CS > 1 (launched as thread-block
// clusters of CS CTAs): the CTAs of a cluster share one cloud -- every CTA takes each CS-th slab of 1024 points, keeps its own
// histogram, and reads its siblings'
// bounding boxes and histograms through distributed shared memory (a point's slot = cells before it + the same cell's counts in
// the lower-ranked CTAs + the CTA's own atomic counter); the block boxes are split between the CTAs after the last cluster
// barrier.  One CTA per 16384-point cloud was 55 us of C2's 170 us forward with 84 of the 148 SMs idle.
template <int CS>
static __global__ void __launch_bounds__(PR_SORT_THREADS) nn_bin_sort_kernel(const PruneSortParams p) {
    __shared__ __align__(16) int hist[PR_MAX_CELLS];
    __shared__ __align__(16) int slot[CS > 1 ? PR_MAX_CELLS : 4];   // CS > 1: next free slot per cell (hist stays readable for the siblings)
    __shared__ float sred[6][PR_SORT_THREADS / 32];
    __shared__ float cbox[8];
    __shared__ int swarp[PR_SORT_THREADS / 32];
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    // cr = rank of this CTA among the `ceff` CTAs that share its cloud, r0 = cluster rank of the first of them
    int cr = 0, r0 = 0, ceff = 1, cloud = (int)blockIdx.x;
    bool idle = false;
    if constexpr (CS > 1) {
        const int rank = (int)cooperative_groups::this_cluster().block_rank();
        if (p.mixed) {
            // clusters [0, B): one cloud of the LARGER side each; the CTAs behind them: one cloud of the other side each, on their
            // own (they only keep the cluster barriers company); CTAs that round the grid up to whole clusters idle
            const int big = p.n[1] > p.n[0] ? 1 : 0;
            if ((int)blockIdx.x < p.B * CS) {
                cloud = big * p.B + (int)blockIdx.x / CS, cr = rank, ceff = CS;
            } else {
                const int j = (int)blockIdx.x - p.B * CS;
                idle = j >= p.B;
                cloud = (1 - big) * p.B + (idle ? 0 : j), r0 = rank;
            }
        } else {
            cloud = (int)blockIdx.x / CS, cr = rank, ceff = CS;
        }
    }
    const int side = p.single_side ? p.side0 : cloud / p.B, b = cloud % p.B;
    const int STRIDE = ceff * PR_SORT_THREADS;
    const int n = idle ? 0 : p.n[side], nblk = pr_nblk(n), npad = nblk * PR_BLOCK;
    const float *src = p.xyz[side] + (size_t)(p.src_div[side] > 1 ? b / p.src_div[side] : b) * n * 3;
    __shared__ Similarity sT;
    const bool moved = p.sim[side] != nullptr;
    if (moved && tid == 0) sT = p.sim[side][b];
    if (moved) __syncthreads();
    auto load_point = [&](int k, float &x, float &y, float &z) {
        x = __ldg(src + k * 3), y = __ldg(src + k * 3 + 1), z = __ldg(src + k * 3 + 2);
        if (moved) apply_similarity(sT, x, y, z);
    };
    float4 *dst = p.sorted[side] + (size_t)b * npad;
    float4 *bx = p.boxes[side] + (size_t)b * 2 * nblk;
    const float inf = __int_as_float(0x7f800000);
    // ---- bounding box + range check ----
    float lo[3] = {inf, inf, inf}, hi[3] = {-inf, -inf, -inf};
    bool bad = false;
    for (int k = cr * PR_SORT_THREADS + tid; k < n; k += STRIDE) {
        float v[3];
        load_point(k, v[0], v[1], v[2]);
#pragma unroll
        for (int c = 0; c < 3; ++c) {
            bad |= !(fabsf(v[c]) <= p.limit);
            lo[c] = fminf(lo[c], v[c]), hi[c] = fmaxf(hi[c], v[c]);
        }
    }
#pragma unroll
    for (int c = 0; c < 3; ++c) {
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) {
            lo[c] = fminf(lo[c], __shfl_xor_sync(0xffffffffu, lo[c], o));
            hi[c] = fmaxf(hi[c], __shfl_xor_sync(0xffffffffu, hi[c], o));
        }
        if (lane == 0) sred[c][warp] = lo[c], sred[3 + c][warp] = hi[c];
    }
#ifndef GENPC_SORT_BITS4_FROM
#define GENPC_SORT_BITS4_FROM 2048   // 16^3 cells from 2048 points up (5 % fewer block visits on C2 than 8^3 cells for the 2048-point clouds)
#endif
    int mbits = n >= GENPC_SORT_BITS4_FROM ? 4 : (n >= 1024 ? 3 : (n >= 128 ? 2 : 1));
    const int ncell = 1 << (3 * mbits);
    for (int k = tid; k < ncell; k += PR_SORT_THREADS) hist[k] = 0;
    const int anybad = __syncthreads_or(bad ? 1 : 0);
    float scale[3];
    if constexpr (CS > 1) {   // the cloud's box = the union of the CTAs' boxes
        if (tid < 6) {
            float r = tid < 3 ? inf : -inf;
            for (int w = 0; w < PR_SORT_THREADS / 32; ++w) r = tid < 3 ? fminf(r, sred[tid][w]) : fmaxf(r, sred[tid][w]);
            cbox[tid] = r;
        }
        cooperative_groups::cluster_group cluster = cooperative_groups::this_cluster();
        cluster.sync();
#pragma unroll
        for (int c = 0; c < 3; ++c) {
            float l = inf, h = -inf;
            for (int r = r0; r < r0 + ceff; ++r) {
                const float *o = cluster.map_shared_rank(cbox, r);
                l = fminf(l, o[c]), h = fmaxf(h, o[3 + c]);
            }
            lo[c] = l, hi[c] = h;
        }
    } else {
#pragma unroll
        for (int c = 0; c < 3; ++c) {
            float l = inf, h = -inf;
            for (int w = 0; w < PR_SORT_THREADS / 32; ++w) l = fminf(l, sred[c][w]), h = fmaxf(h, sred[3 + c][w]);
            lo[c] = l, hi[c] = h;
        }
    }
    if (tid == 0 && cr == 0 && !idle && p.bbx != nullptr) {   // read by the last CTA of the grid, behind its ticket (below)
        float *o = p.bbx + ((size_t)side * p.B + b) * 8;
        for (int c = 0; c < 3; ++c) o[c] = lo[c], o[3 + c] = hi[c];
    }
#pragma unroll
    for (int c = 0; c < 3; ++c) {
        const float ext = hi[c] - lo[c];
        scale[c] = (ext > 0.f && ext < inf) ? (float)(1 << mbits) / ext : 0.f;
    }
    const int cmax = (1 << mbits) - 1;
    const bool hilbert = p.hilbert != 0;
    auto cell_of = [&](int k, float &x, float &y, float &z) {
        load_point(k, x, y, z);
        const float v[3] = {x, y, z};
        unsigned code = 0;
        unsigned cc[3];
#pragma unroll
        for (int c = 0; c < 3; ++c) {
            const float f = (v[c] - lo[c]) * scale[c];
            cc[c] = f >= 0.f ? (unsigned)min((int)fminf(f, 2e9f), cmax) : 0u;
        }
        if (hilbert) {   // connected runs: tighter block boxes than the Z-order below
            code = mbits == 4 ? hilbert_key<4>(cc[0], cc[1], cc[2]) : mbits == 3 ? hilbert_key<3>(cc[0], cc[1], cc[2])
                 : mbits == 2 ? hilbert_key<2>(cc[0], cc[1], cc[2]) : hilbert_key<1>(cc[0], cc[1], cc[2]);
            return (int)code;
        }
        for (int bit = 0; bit < mbits; ++bit)
#pragma unroll
            for (int c = 0; c < 3; ++c) code |= (unsigned)((cc[c] >> bit) & 1) << (3 * bit + c);
        return (int)code;
    };
    // ---- histogram; clouds of up to 16384 points keep each point's cell in registers for the scatter pass (the Hilbert key is
    // a third of this kernel's instructions when it is computed in both passes) ----
    constexpr int KEEP = CS == 1 ? 16 : (PR_MAX_N + CS * PR_SORT_THREADS - 1) / (CS * PR_SORT_THREADS);   // CS > 1: every size up to PR_MAX_N
    const bool keep = n <= KEEP * STRIDE;
    unsigned short cells[KEEP];
    if (keep) {
#pragma unroll
        for (int i = 0; i < KEEP; ++i) {
            const int k = cr * PR_SORT_THREADS + tid + i * STRIDE;
            cells[i] = 0;
            if (k < n) {
                float x, y, z;
                cells[i] = (unsigned short)cell_of(k, x, y, z);
                atomicAdd(&hist[cells[i]], 1);
            }
        }
    } else {
        for (int k = cr * PR_SORT_THREADS + tid; k < n; k += STRIDE) {
            float x, y, z;
            atomicAdd(&hist[cell_of(k, x, y, z)], 1);
        }
    }
    if constexpr (CS > 1) cooperative_groups::this_cluster().sync();   // every sibling's histogram is complete
    else __syncthreads();
    // ---- exclusive scan of the histogram (4 entries per thread; CS > 1: of the sum of the CS histograms) ----
    int *next = CS > 1 ? slot : hist;
    {
        constexpr int per = PR_MAX_CELLS / PR_SORT_THREADS;
        static_assert(per == 4, "one int4 of cells per thread");
        int v[per], below[per], s = 0;
#pragma unroll
        for (int i = 0; i < per; ++i) v[i] = below[i] = 0;
        if (tid * per < ncell) {   // ncell is 8, 64, 512 or 4096: a thread's four cells are all inside or all outside
            if constexpr (CS > 1) {
                cooperative_groups::cluster_group cluster = cooperative_groups::this_cluster();
                for (int r = r0; r < r0 + ceff; ++r) {
                    const int4 t = *reinterpret_cast<const int4 *>(cluster.map_shared_rank(hist, r) + tid * per);
                    const int tv[per] = {t.x, t.y, t.z, t.w};
#pragma unroll
                    for (int i = 0; i < per; ++i) {
                        v[i] += tv[i];
                        if (r < r0 + cr) below[i] += tv[i];
                    }
                }
            } else {
                const int4 t = *reinterpret_cast<const int4 *>(hist + tid * per);
                v[0] = t.x, v[1] = t.y, v[2] = t.z, v[3] = t.w;
            }
        }
#pragma unroll
        for (int i = 0; i < per; ++i) s += v[i];
        int incl = s;
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) {
            const int t = __shfl_up_sync(0xffffffffu, incl, o);
            if (lane >= o) incl += t;
        }
        if (lane == 31) swarp[warp] = incl;
        __syncthreads();
        int base = 0;
        for (int w = 0; w < warp; ++w) base += swarp[w];
        int run = base + incl - s;
#pragma unroll
        for (int i = 0; i < per; ++i) {
            const int e = tid * per + i;
            if (e < ncell) next[e] = run + below[i];
            run += v[i];
        }
    }
    __syncthreads();
    // ---- scatter (the order inside a cell is whatever the atomics give: it only shapes the blocks, never a result) ----
    if (keep) {
#pragma unroll
        for (int i = 0; i < KEEP; ++i) {
            const int k = cr * PR_SORT_THREADS + tid + i * STRIDE;
            if (k < n) {
                float x, y, z;
                load_point(k, x, y, z);
                const int pos = atomicAdd(&next[cells[i]], 1);
                dst[pos] = make_float4(x, y, z, __int_as_float(k));
            }
        }
    } else {
        for (int k = cr * PR_SORT_THREADS + tid; k < n; k += STRIDE) {
            float x, y, z;
            const int cell = cell_of(k, x, y, z);
            const int pos = atomicAdd(&next[cell], 1);
            dst[pos] = make_float4(x, y, z, __int_as_float(k));
        }
    }
    const float qnan = __int_as_float(0x7fc00000);
    if (cr == 0)
        for (int k = n + tid; k < npad; k += PR_SORT_THREADS) dst[k] = make_float4(qnan, qnan, qnan, __int_as_float(0));
    if constexpr (CS > 1) {
        // the siblings' records are read below: fence + cluster barrier, and the loads bypass L1 (__ldcg).  The barrier also is
        // the last point at which a sibling reads this CTA's shared memory: nobody exits before it
        __threadfence();
        cooperative_groups::this_cluster().sync();
    } else {
        __syncthreads();   // the CTA's own global writes are visible to it after the barrier
    }
    // ---- block boxes ----
    for (int blk = cr * (PR_SORT_THREADS / 32) + warp; blk < nblk; blk += ceff * (PR_SORT_THREADS / 32)) {
        float l[3] = {inf, inf, inf}, h[3] = {-inf, -inf, -inf};
#pragma unroll
        for (int e = 0; e < PR_BLOCK / 32; ++e) {
            const float4 t = CS > 1 ? __ldcg(dst + blk * PR_BLOCK + e * 32 + lane) : dst[blk * PR_BLOCK + e * 32 + lane];
            l[0] = fminf(l[0], t.x), h[0] = fmaxf(h[0], t.x);   // NaN padding is ignored
            l[1] = fminf(l[1], t.y), h[1] = fmaxf(h[1], t.y);
            l[2] = fminf(l[2], t.z), h[2] = fmaxf(h[2], t.z);
        }
#pragma unroll
        for (int c = 0; c < 3; ++c)
#pragma unroll
            for (int o = 16; o > 0; o >>= 1) {
                l[c] = fminf(l[c], __shfl_xor_sync(0xffffffffu, l[c], o));
                h[c] = fmaxf(h[c], __shfl_xor_sync(0xffffffffu, h[c], o));
            }
        if (lane == 0) bx[blk] = make_float4(l[0], l[1], l[2], 0.f), bx[nblk + blk] = make_float4(h[0], h[1], h[2], 0.f);
    }
    // ---- selection flag: the last CTA publishes "any cloud out of range, or a pair of clouds that does not overlap" (pruning
    // needs neighbours to be near: when the two boxes of a pair are further apart than a quarter of the smaller one's diagonal,
    // or one box is less than a quarter of the other across, every query group would open every block, 2.8x the cost of the
    // exhaustive scan) ----
    if (tid == 0) {
        if (anybad) atomicOr(p.ctl + 2, 1);
        __threadfence();
        if (atomicAdd(p.ctl + 3, 1) == (int)gridDim.x - 1) {
            __threadfence();
            int apart = 0;
            if (p.bbx != nullptr && !p.single_side) {
                for (int i = 0; i < p.B; ++i) {
                    const volatile float *u = p.bbx + (size_t)i * 8, *v = p.bbx + ((size_t)p.B + i) * 8;
                    float gap2 = 0.f, du = 0.f, dv = 0.f;
                    for (int c = 0; c < 3; ++c) {
                        const float g = fmaxf(fmaxf(u[c] - v[3 + c], v[c] - u[3 + c]), 0.f);
                        gap2 += g * g, du += (u[3 + c] - u[c]) * (u[3 + c] - u[c]), dv += (v[3 + c] - v[c]) * (v[3 + c] - v[c]);
                    }
                    if (gap2 > 0.0625f * fminf(du, dv)) apart = 1;
                    // ... and a cloud collapsed to a blob next to a spread-out one (an untrained generator's output against its
                    // target): every group of the large cloud finds all blocks of the blob equally near and opens them all
                    if (fminf(du, dv) < 0.0625f * fmaxf(du, dv)) apart = 1;
                }
            }
            const int verdict = atomicExch(p.ctl + 2, 0) | apart;
            if (p.flag != nullptr) {
                if (verdict) atomicOr(p.flag, 1);
            } else {
                p.ctl[1] = p.accumulate ? (p.ctl[1] | verdict) : verdict;
            }
            p.ctl[3] = 0;
        }
    }
}

// `ctas` CTAs (rounded up to whole clusters) as thread-block clusters of CS
template <int CS>
static void launch_bin_sort_cs(PruneSortParams sp, int ctas, cudaStream_t stream) {
    cudaLaunchConfig_t cfg = {};
    cfg.gridDim = dim3((unsigned)((ctas + CS - 1) / CS * CS));
    cfg.blockDim = dim3(PR_SORT_THREADS);
    cfg.stream = stream;
    cudaLaunchAttribute attr[1];
    attr[0].id = cudaLaunchAttributeClusterDimension;
    attr[0].val.clusterDim.x = CS, attr[0].val.clusterDim.y = 1, attr[0].val.clusterDim.z = 1;
    cfg.attrs = attr, cfg.numAttrs = 1;
    if (cudaLaunchKernelEx(&cfg, nn_bin_sort_kernel<CS>, sp) != cudaSuccess) {
        // a part whose GPCs cannot co-schedule CS 1024-thread CTAs (MIG slices, harvested parts): one CTA per cloud
        (void)cudaGetLastError();
        sp.mixed = 0;
        nn_bin_sort_kernel<1><<<(sp.single_side ? 1 : 2) * sp.B, PR_SORT_THREADS, 0, stream>>>(sp);
    }
}


struct PruneParams {
    const float4 *q, *t;      // sorted queries [B][npad_q], sorted targets [B][npad_t]
    const float4 *tbox;       // [B][2][nblk_t]
    unsigned long long *out;  // [B][nq] packed (dist, original target index), addressed by the query's original index
    int nq, nt, B;
    const int *select;        // run only when *select == 0
    unsigned *stats;          // optional [4]: blocks scanned, tie passes, groups, -
    int qdiv, tdiv;           // batch entry b reads queries / targets of cloud b / qdiv, b / tdiv (0 or 1: its own; registration: the
                              // starts of one scan share the fixed cloud's sorted copy)
};

__device__ __forceinline__ int pr_f2ord(float f) {
    const int i = __float_as_int(f);
    return i ^ ((i >> 31) & 0x7fffffff);
}
__device__ __forceinline__ float pr_ord2f(int i) { return __int_as_float(i ^ ((i >> 31) & 0x7fffffff)); }

// Both directions of a cloud pair in ONE launch: the first ctas0 CTAs take direction d[0] (put the direction with the longer
// per-group chains there: it is scheduled first), the rest d[1].
struct PrunePair {
    PruneParams d[2];
    int ctas0;
    int ctas1_unused;   // (host-side bookkeeping of the registration loop)
};

// grid = sum over the two directions of B * ceil(groups / 8) CTAs of 8 warps; BOXR * 32 >= nblk of either target cloud
template <int BOXR>
__global__ void __launch_bounds__(PR_THREADS) nn_prune_kernel(const PrunePair pp) {
    __shared__ __align__(16) float stage[PR_THREADS / 32][3][PR_BLOCK];   // the block being walked, SoA: x[64] y[64] z[64]
    const int dir = (int)blockIdx.x >= pp.ctas0 ? 1 : 0;
    const PruneParams &p = pp.d[dir];
    if (p.select != nullptr && *p.select != 0) return;
    const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
    const int groups = (p.nq + PR_GROUP - 1) / PR_GROUP;
    const int ctas_per_cloud = (groups + PR_THREADS / 32 - 1) / (PR_THREADS / 32);
    const int bx = (int)blockIdx.x - dir * pp.ctas0;
    const int b = bx / ctas_per_cloud;
    const int g = (bx % ctas_per_cloud) * (PR_THREADS / 32) + wid;
    if (g >= groups) return;
    const int nblk = pr_nblk(p.nt);
    const int bt = p.tdiv > 1 ? b / p.tdiv : b, bq = p.qdiv > 1 ? b / p.qdiv : b;
    const float4 *T = p.t + (size_t)bt * pr_npad(p.nt);
    const float4 *BL = p.tbox + (size_t)bt * 2 * nblk, *BH = BL + nblk;
    const float inf = __int_as_float(0x7f800000);
    const int qi = g * PR_GROUP + lane;
    const bool valid = qi < p.nq;
    const float4 q = p.q[(size_t)bq * pr_npad(p.nq) + qi];   // the padding records are NaN: they never win, nothing is stored
    // ---- the group's box ----
    float glo[3], ghi[3];
    {
        const float v[3] = {q.x, q.y, q.z};
#pragma unroll
        for (int c = 0; c < 3; ++c) {
            // NaN (padding) must not take part: +inf / -inf are neutral
            const float a = valid ? v[c] : inf, z = valid ? v[c] : -inf;
            glo[c] = pr_ord2f(__reduce_min_sync(0xffffffffu, pr_f2ord(a)));
            ghi[c] = pr_ord2f(__reduce_max_sync(0xffffffffu, pr_f2ord(z)));
        }
    }
    // ---- squared box-to-box distances, scaled down (see the header) ----
    float bd[BOXR];
#pragma unroll
    for (int r = 0; r < BOXR; ++r) {
        const int blk = r * 32 + lane;
        bd[r] = inf;
        if (blk < nblk) {
            const float4 lo = __ldg(BL + blk), hi = __ldg(BH + blk);
            const float dx = fmaxf(fmaxf(lo.x - ghi[0], glo[0] - hi.x), 0.f);
            const float dy = fmaxf(fmaxf(lo.y - ghi[1], glo[1] - hi.y), 0.f);
            const float dz = fmaxf(fmaxf(lo.z - ghi[2], glo[2] - hi.z), 0.f);
            bd[r] = __fmul_rn(__fmaf_rn(dz, dz, __fmaf_rn(dx, dx, __fmul_rn(dy, dy))), 0.99999f);
        }
    }
    const float2 nx = make_float2(-q.x, -q.x), ny = make_float2(-q.y, -q.y), nz = make_float2(-q.z, -q.z);
    float best = inf;
    int bchunk = 0;
    bool tie = false;
    float thr = inf;   // the largest running minimum of the group's lanes
    unsigned scanned = 0;
    for (;;) {
        // nearest remaining block: (distance bits without the low 9, block id); the truncation errs on the near side
        unsigned key = 0xffffffffu;
#pragma unroll
        for (int r = 0; r < BOXR; ++r) key = min(key, (__float_as_uint(bd[r]) & 0xfffffe00u) | (unsigned)(r * 32 + lane));
        key = __reduce_min_sync(0xffffffffu, key);
        if (key >= 0x7f800000u || !(__uint_as_float(key & 0xfffffe00u) <= thr)) break;
        const int blk = (int)(key & 0x1ffu);
#pragma unroll
        for (int r = 0; r < BOXR; ++r)
            if (r * 32 + lane == blk) bd[r] = inf;
        {   // one coalesced 1 KB read per block, walked in shared memory (see nn_prune2_kernel)
            const float4 *tg = T + blk * PR_BLOCK;
            const float4 u0 = __ldg(tg + lane), u1 = __ldg(tg + 32 + lane);
            __syncwarp();
            stage[wid][0][lane] = u0.x, stage[wid][1][lane] = u0.y, stage[wid][2][lane] = u0.z;
            stage[wid][0][32 + lane] = u1.x, stage[wid][1][32 + lane] = u1.y, stage[wid][2][32 + lane] = u1.z;
            __syncwarp();
        }
        // SoA: one LDS.128 per coordinate brings four targets as two register pairs the packed FP32 instructions take as they
        // are (the AoS form spent 6 of its 15 instructions per target pair moving registers into pairs)
        const float4 *sx4 = reinterpret_cast<const float4 *>(stage[wid][0]), *sy4 = reinterpret_cast<const float4 *>(stage[wid][1]),
                     *sz4 = reinterpret_cast<const float4 *>(stage[wid][2]);
        // 256 target blocks and more (C3: 16384^2 inside the registration loop) keep the 2x form: 4x cost C3 9 % on the same box
        // (26 100 vs 28 600 scan-iters/s) while the 32- and 128-block instantiations gained 1-4 % (profiles/r02w_unroll_sweep.txt)
        constexpr int UNR = BOXR >= 8 ? 2 : PR_UNROLL;
#pragma unroll UNR
        for (int c = 0; c < PR_BLOCK / 8; ++c) {
            float cm = inf;
#pragma unroll
            for (int i = 0; i < 2; ++i) {
                const float4 X = sx4[c * 2 + i], Y = sy4[c * 2 + i], Z = sz4[c * 2 + i];
                const float2 a2 = sqdist_ref_x2(nx, ny, nz, make_float2(X.x, X.y), make_float2(Y.x, Y.y), make_float2(Z.x, Z.y));
                const float2 b2 = sqdist_ref_x2(nx, ny, nz, make_float2(X.z, X.w), make_float2(Y.z, Y.w), make_float2(Z.z, Z.w));
                cm = fmin3(cm, a2.x, a2.y);
                cm = fmin3(cm, b2.x, b2.y);
            }
            if (cm < best) {
                best = cm, bchunk = blk * (PR_BLOCK / 8) + c, tie = false;
            } else if (cm == best && cm < inf) {
                tie = true;
            }
        }
        ++scanned;
        thr = pr_ord2f(__reduce_max_sync(0xffffffffu, valid ? pr_f2ord(best) : (int)0x80000000));
    }
    // ---- the lowest original index at the minimum: inside the winning chunk ... ----
    int bidx = 0x7fffffff;
    {
        const float4 *tc = T + bchunk * 8;
#pragma unroll
        for (int i = 0; i < 8; ++i) {
            const float4 t = __ldg(tc + i);
            const float d = sqdist_ref(q.x, q.y, q.z, t.x, t.y, t.z);
            if (d == best) bidx = min(bidx, __float_as_int(t.w));
        }
    }
    // ---- ... and, for a lane that met its minimum in two chunks, over every block that could hold it ----
    const unsigned ties = __ballot_sync(0xffffffffu, tie && valid);
    if (ties != 0u) {
        for (int blk = 0; blk < nblk; ++blk) {
            const float4 lo = __ldg(BL + blk), hi = __ldg(BH + blk);
            const float dx = fmaxf(fmaxf(lo.x - ghi[0], glo[0] - hi.x), 0.f);
            const float dy = fmaxf(fmaxf(lo.y - ghi[1], glo[1] - hi.y), 0.f);
            const float dz = fmaxf(fmaxf(lo.z - ghi[2], glo[2] - hi.z), 0.f);
            const float s = __fmul_rn(__fmaf_rn(dz, dz, __fmaf_rn(dx, dx, __fmul_rn(dy, dy))), 0.99999f);
            if (!(s <= thr)) continue;   // warp-uniform
            const float4 *tb = T + blk * PR_BLOCK;
            for (int i = 0; i < PR_BLOCK; ++i) {
                const float4 t = __ldg(tb + i);
                const float d = sqdist_ref(q.x, q.y, q.z, t.x, t.y, t.z);
                if (tie && d == best) bidx = min(bidx, __float_as_int(t.w));
            }
        }
    }
    if (valid) p.out[(size_t)b * p.nq + __float_as_int(q.w)] = pack_dist_idx(best, bidx);
    if (p.stats != nullptr && lane == 0) {
        atomicAdd(p.stats + 0, scanned);
        if (ties != 0u) atomicAdd(p.stats + 1, 1u);
        atomicAdd(p.stats + 2, 1u);
    }
}

// Cooperative form for a direction with FEW query groups and many target blocks (C2: 2048 groups against 256 blocks each): one
// CTA per group, each of its eight warps walks an eighth of the blocks nearest first; the running minima are shared through
// shared-memory atomics, so every warp stops where a single warp would (its bound is never below the final one: no block that
// matters is skipped), and warp 0 merges the eight partial results.  A group's dependent chain -- up to 256
// block visits of 0.35 us in one warp, which set the duration of the whole launch (118 us on one batch of the bench's generator,
// 167 us on others with the same instruction count) -- becomes an eighth as long.  Launched on a side stream next to the other
// direction's nn_prune_kernel.
template <int BOXR>
__global__ void __launch_bounds__(PR_THREADS) nn_prune_coop_kernel(const PruneParams p) {
    __shared__ __align__(16) float stage[PR_THREADS / 32][3][PR_BLOCK];
    __shared__ unsigned s_best[PR_GROUP];                       // running minimum per query over all warps (float bits, >= 0)
    __shared__ float m_best[PR_THREADS / 32][PR_GROUP];
    __shared__ int m_chunk[PR_THREADS / 32][PR_GROUP], m_tie[PR_THREADS / 32][PR_GROUP];
    if (p.select != nullptr && *p.select != 0) return;
    const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
    const int groups = (p.nq + PR_GROUP - 1) / PR_GROUP;
    const int b = (int)blockIdx.x / groups;
    const int g = (int)blockIdx.x % groups;
    if (wid == 0) s_best[lane] = 0x7f800000u;
    __syncthreads();
    const int nblk = pr_nblk(p.nt);
    const int bt = p.tdiv > 1 ? b / p.tdiv : b, bq = p.qdiv > 1 ? b / p.qdiv : b;
    const float4 *T = p.t + (size_t)bt * pr_npad(p.nt);
    const float4 *BL = p.tbox + (size_t)bt * 2 * nblk, *BH = BL + nblk;
    const float inf = __int_as_float(0x7f800000);
    const int qi = g * PR_GROUP + lane;
    const bool valid = qi < p.nq;
    const float4 q = p.q[(size_t)bq * pr_npad(p.nq) + qi];   // the padding records are NaN: they never win, nothing is stored
    // ---- the group's box ----
    float glo[3], ghi[3];
    {
        const float v[3] = {q.x, q.y, q.z};
#pragma unroll
        for (int c = 0; c < 3; ++c) {
            // NaN (padding) must not take part: +inf / -inf are neutral
            const float a = valid ? v[c] : inf, z = valid ? v[c] : -inf;
            glo[c] = pr_ord2f(__reduce_min_sync(0xffffffffu, pr_f2ord(a)));
            ghi[c] = pr_ord2f(__reduce_max_sync(0xffffffffu, pr_f2ord(z)));
        }
    }
    // ---- squared box-to-box distances, scaled down (see the header) ----
    // warp w owns the blocks w, w + 8, w + 16, ... (consecutive Hilbert blocks are neighbours: every warp gets an even sample
    // of the cloud) and runs the nearest-first walk over ITS blocks only -- no selection work is repeated
    constexpr int WARPS = PR_THREADS / 32;
    constexpr int CBOXR = (BOXR + WARPS - 1) / WARPS;
    float bd[CBOXR];
#pragma unroll
    for (int r = 0; r < CBOXR; ++r) {
        const int blk = wid + WARPS * (r * 32 + lane);
        bd[r] = inf;
        if (blk < nblk) {
            const float4 lo = __ldg(BL + blk), hi = __ldg(BH + blk);
            const float dx = fmaxf(fmaxf(lo.x - ghi[0], glo[0] - hi.x), 0.f);
            const float dy = fmaxf(fmaxf(lo.y - ghi[1], glo[1] - hi.y), 0.f);
            const float dz = fmaxf(fmaxf(lo.z - ghi[2], glo[2] - hi.z), 0.f);
            bd[r] = __fmul_rn(__fmaf_rn(dz, dz, __fmaf_rn(dx, dx, __fmul_rn(dy, dy))), 0.99999f);
        }
    }
    const float2 nx = make_float2(-q.x, -q.x), ny = make_float2(-q.y, -q.y), nz = make_float2(-q.z, -q.z);
    float best = inf;
    int bchunk = 0;
    bool tie = false;
    float thr = inf;   // the largest running minimum of the group's lanes
    unsigned scanned = 0;
    for (;;) {
        unsigned key = 0xffffffffu;   // (distance bits without the low 9, slot = r * 32 + lane of the owned block)
#pragma unroll
        for (int r = 0; r < CBOXR; ++r) key = min(key, (__float_as_uint(bd[r]) & 0xfffffe00u) | (unsigned)(r * 32 + lane));
        key = __reduce_min_sync(0xffffffffu, key);
        // the bound uses what ANY warp has found so far for each query
        thr = pr_ord2f(__reduce_max_sync(0xffffffffu, valid ? pr_f2ord(fminf(best, __uint_as_float(*(volatile unsigned *)&s_best[lane])))
                                                            : (int)0x80000000));
        if (key >= 0x7f800000u || !(__uint_as_float(key & 0xfffffe00u) <= thr)) break;
        const int slot = (int)(key & 0x1ffu);
#pragma unroll
        for (int r = 0; r < CBOXR; ++r)
            if (r * 32 + lane == slot) bd[r] = inf;
        const int blk = wid + WARPS * slot;
        {   // one coalesced 1 KB read per block, walked in shared memory (see nn_prune2_kernel)
            const float4 *tg = T + blk * PR_BLOCK;
            const float4 u0 = __ldg(tg + lane), u1 = __ldg(tg + 32 + lane);
            __syncwarp();
            stage[wid][0][lane] = u0.x, stage[wid][1][lane] = u0.y, stage[wid][2][lane] = u0.z;
            stage[wid][0][32 + lane] = u1.x, stage[wid][1][32 + lane] = u1.y, stage[wid][2][32 + lane] = u1.z;
            __syncwarp();
        }
        // SoA: one LDS.128 per coordinate brings four targets as two register pairs the packed FP32 instructions take as they
        // are (the AoS form spent 6 of its 15 instructions per target pair moving registers into pairs)
        const float4 *sx4 = reinterpret_cast<const float4 *>(stage[wid][0]), *sy4 = reinterpret_cast<const float4 *>(stage[wid][1]),
                     *sz4 = reinterpret_cast<const float4 *>(stage[wid][2]);
#pragma unroll PR_UNROLL
        for (int c = 0; c < PR_BLOCK / 8; ++c) {
            float cm = inf;
#pragma unroll
            for (int i = 0; i < 2; ++i) {
                const float4 X = sx4[c * 2 + i], Y = sy4[c * 2 + i], Z = sz4[c * 2 + i];
                const float2 a2 = sqdist_ref_x2(nx, ny, nz, make_float2(X.x, X.y), make_float2(Y.x, Y.y), make_float2(Z.x, Z.y));
                const float2 b2 = sqdist_ref_x2(nx, ny, nz, make_float2(X.z, X.w), make_float2(Y.z, Y.w), make_float2(Z.z, Z.w));
                cm = fmin3(cm, a2.x, a2.y);
                cm = fmin3(cm, b2.x, b2.y);
            }
            if (cm < best) {
                best = cm, bchunk = blk * (PR_BLOCK / 8) + c, tie = false;
            } else if (cm == best && cm < inf) {
                tie = true;
            }
        }
        ++scanned;
        if (valid) atomicMin(&s_best[lane], __float_as_uint(best));
    }
    // ---- merge the eight warps' results (warp 0 finishes the group) ----
    m_best[wid][lane] = best, m_chunk[wid][lane] = bchunk, m_tie[wid][lane] = tie ? 1 : 0;
    if (p.stats != nullptr && lane == 0 && wid != 0) atomicAdd(p.stats + 0, scanned);
    __syncthreads();
    if (wid != 0) return;
#pragma unroll
    for (int w = 1; w < PR_THREADS / 32; ++w) {
        const float mb = m_best[w][lane];
        const int mc = m_chunk[w][lane];
        const bool mt = m_tie[w][lane] != 0;
        if (mb < best) {
            best = mb, bchunk = mc, tie = mt;
        } else if (mb == best && mb < inf) {
            tie = tie || mt || mc != bchunk;
        }
    }
    thr = pr_ord2f(__reduce_max_sync(0xffffffffu, valid ? pr_f2ord(best) : (int)0x80000000));
    // ---- the lowest original index at the minimum: inside the winning chunk ... ----
    int bidx = 0x7fffffff;
    {
        const float4 *tc = T + bchunk * 8;
#pragma unroll
        for (int i = 0; i < 8; ++i) {
            const float4 t = __ldg(tc + i);
            const float d = sqdist_ref(q.x, q.y, q.z, t.x, t.y, t.z);
            if (d == best) bidx = min(bidx, __float_as_int(t.w));
        }
    }
    // ---- ... and, for a lane that met its minimum in two chunks, over every block that could hold it ----
    const unsigned ties = __ballot_sync(0xffffffffu, tie && valid);
    if (ties != 0u) {
        for (int blk = 0; blk < nblk; ++blk) {
            const float4 lo = __ldg(BL + blk), hi = __ldg(BH + blk);
            const float dx = fmaxf(fmaxf(lo.x - ghi[0], glo[0] - hi.x), 0.f);
            const float dy = fmaxf(fmaxf(lo.y - ghi[1], glo[1] - hi.y), 0.f);
            const float dz = fmaxf(fmaxf(lo.z - ghi[2], glo[2] - hi.z), 0.f);
            const float s = __fmul_rn(__fmaf_rn(dz, dz, __fmaf_rn(dx, dx, __fmul_rn(dy, dy))), 0.99999f);
            if (!(s <= thr)) continue;   // warp-uniform
            const float4 *tb = T + blk * PR_BLOCK;
            for (int i = 0; i < PR_BLOCK; ++i) {
                const float4 t = __ldg(tb + i);
                const float d = sqdist_ref(q.x, q.y, q.z, t.x, t.y, t.z);
                if (tie && d == best) bidx = min(bidx, __float_as_int(t.w));
            }
        }
    }
    if (valid) p.out[(size_t)b * p.nq + __float_as_int(q.w)] = pack_dist_idx(best, bidx);
    if (p.stats != nullptr && lane == 0) {
        atomicAdd(p.stats + 0, scanned);
        if (ties != 0u) atomicAdd(p.stats + 1, 1u);
        atomicAdd(p.stats + 2, 1u);
    }
}

}  // namespace genpc
