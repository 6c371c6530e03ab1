// nn_grid.cuh -- spatially pruned exact nearest neighbours for LARGE clouds (r02): the multi-CTA form of nn_prune.cuh.
//
// nn_prune.cuh sorts a cloud inside one CTA and keeps every block's box distance in registers: good for <= 32768 points.
// The exhaustive scan is O(N * M): BASELINE C5 (1M x 1M) takes 228 ms on one GPU; this path 2.2 ms, same bits.  Default from
// 2^32 evaluations with more than 32768 points on a side (B <= 8); a sampled probe hands clouds that do not overlap back to
// the exhaustive kernels.  Here:
//   * grid_* kernels: bounding box (ordered-int atomics) -> 30-bit Hilbert keys -> radix sort of (key, index) pairs
//     (cub::DeviceRadixSort, the one library call of the path) -> gather into
//     (x, y, z, original index) records -> boxes of every 64 records (block) and of every 64 blocks (superblock);
//   * nn_prune2_kernel: one warp owns 32 consecutive sorted queries; superblocks are visited nearest first (their box
//     distances live in SBR registers per lane), inside a superblock its blocks nearest first, and both loops stop at the
//     first box farther than the group's worst running minimum.  The block scan, the index recovery and the tie pass are
//     those of nn_prune_kernel: same arithmetic, same bits as the exhaustive kernels.
#pragma once
#include <cub/device/device_radix_sort.cuh>

#include "nn_prune.cuh"

namespace genpc {

constexpr int GR_SUPER = 64;              // blocks per superblock
constexpr int GR_MAX_N = 2 * 1024 * 1024; // 512 superblocks: the id fits the 9 low bits of the selection key
constexpr int GR_THREADS = 256;

__host__ __device__ inline int gr_nsb(int n) { return (pr_nblk(n) + GR_SUPER - 1) / GR_SUPER; }

struct GridParams {
    const float *xyz[2];   // [B][n][3]
    float4 *sorted[2];     // [B][npad]
    float4 *boxes[2];      // [B][2][nblk]
    float4 *sboxes[2];     // [B][2][nsb]
    int *bb[2];            // [B][8]: ordered-int min xyz, max xyz, out-of-range flag, -
    unsigned *keys;        // [max n] Morton keys of the cloud being sorted (one cloud at a time)
    int *vals;             // [max n] 0 .. n-1
    const int *order;      // [max n] the sorted permutation (gather kernel)
    int n[2];
    int B;
    float limit;
    int *ctl;              // [1] selection flag as in nn_prune.cuh
};

static __global__ void grid_init_kernel(const GridParams p) {
    for (int i = threadIdx.x; i < 2 * p.B * 8; i += blockDim.x) {
        const int side = i / (p.B * 8), r = i % (p.B * 8), k = r & 7;
        p.bb[side][r] = k < 3 ? 0x7fffffff : (k < 6 ? (int)0x80000000 : 0);
    }
}

// grid (ctas, B, 2)
static __global__ void __launch_bounds__(GR_THREADS) grid_bbox_kernel(const GridParams p) {
    const int side = blockIdx.z, b = blockIdx.y, n = p.n[side];
    const float *src = p.xyz[side] + (size_t)b * n * 3;
    int lo[3] = {0x7fffffff, 0x7fffffff, 0x7fffffff}, hi[3] = {(int)0x80000000, (int)0x80000000, (int)0x80000000};
    bool bad = false;
    for (int k = blockIdx.x * GR_THREADS + threadIdx.x; k < n; k += gridDim.x * GR_THREADS) {
#pragma unroll
        for (int c = 0; c < 3; ++c) {
            const float v = __ldg(src + (size_t)k * 3 + c);
            if (fabsf(v) <= p.limit) {
                const int o = pr_f2ord(v);
                lo[c] = min(lo[c], o), hi[c] = max(hi[c], o);
            } else {
                bad = true;
            }
        }
    }
    int *bb = p.bb[side] + b * 8;
#pragma unroll
    for (int c = 0; c < 3; ++c) {
        const int l = __reduce_min_sync(0xffffffffu, lo[c]), h = __reduce_max_sync(0xffffffffu, hi[c]);
        if ((threadIdx.x & 31) == 0) atomicMin(bb + c, l), atomicMax(bb + 3 + c, h);
    }
    if (__any_sync(0xffffffffu, bad) && (threadIdx.x & 31) == 0) atomicOr(bb + 6, 1);
}

// one cloud (side, b): 30-bit Hilbert keys over its bounding box, values 0 .. n-1
static __global__ void __launch_bounds__(GR_THREADS) grid_key_kernel(const GridParams p, int side, int b) {
    const int n = p.n[side];
    const float *src = p.xyz[side] + (size_t)b * n * 3;
    const int *bb = p.bb[side] + b * 8;
    float lo[3], scale[3];
#pragma unroll
    for (int c = 0; c < 3; ++c) {
        const float l = pr_ord2f(bb[c]), h = pr_ord2f(bb[3 + c]);
        const float ext = h - l;
        lo[c] = l;
        scale[c] = (ext > 0.f && ext < 3.0e38f) ? 1024.f / ext : 0.f;
    }
    for (int k = blockIdx.x * GR_THREADS + threadIdx.x; k < n; k += gridDim.x * GR_THREADS) {
        unsigned cc[3];
#pragma unroll
        for (int c = 0; c < 3; ++c) {
            const float f = (__ldg(src + (size_t)k * 3 + c) - lo[c]) * scale[c];
            cc[c] = f >= 0.f ? (unsigned)min((int)fminf(f, 2e9f), 1023) : 0u;   // NaN -> 0
        }
        p.keys[k] = hilbert_key<10>(cc[0], cc[1], cc[2]);
        p.vals[k] = k;
    }
}

// one cloud (side, b): records in sorted order + NaN padding of the last block; the first launch publishes the selection flag
static __global__ void __launch_bounds__(GR_THREADS) grid_gather_kernel(const GridParams p, int side, int b, int publish) {
    const int n = p.n[side], npad = pr_npad(n);
    const float *src = p.xyz[side] + (size_t)b * n * 3;
    float4 *dst = p.sorted[side] + (size_t)b * npad;
    for (int k = blockIdx.x * GR_THREADS + threadIdx.x; k < npad; k += gridDim.x * GR_THREADS) {
        if (k < n) {
            const int o = __ldg(p.order + k);
            dst[k] = make_float4(__ldg(src + (size_t)o * 3), __ldg(src + (size_t)o * 3 + 1), __ldg(src + (size_t)o * 3 + 2), __int_as_float(o));
        } else {
            const float qnan = __int_as_float(0x7fc00000);
            dst[k] = make_float4(qnan, qnan, qnan, __int_as_float(0));
        }
    }
    if (publish && blockIdx.x == 0 && threadIdx.x == 0) {
        int bad = 0;
        for (int sd = 0; sd < 2; ++sd)
            for (int i = 0; i < p.B; ++i) bad |= p.bb[sd][i * 8 + 6];
        p.ctl[1] = bad;
    }
}

// grid (ceil(count / 8), B, 2): SUPER = false: boxes of 64 records; true: boxes of 64 block boxes
template <bool SUPER>
static __global__ void __launch_bounds__(GR_THREADS) grid_boxes_kernel(const GridParams p) {
    const int side = blockIdx.z, b = blockIdx.y, n = p.n[side];
    const int nblk = pr_nblk(n), nsb = gr_nsb(n);
    const int lane = threadIdx.x & 31;
    const int item = blockIdx.x * (GR_THREADS / 32) + (threadIdx.x >> 5);
    const float inf = __int_as_float(0x7f800000);
    float l[3] = {inf, inf, inf}, h[3] = {-inf, -inf, -inf};
    if (!SUPER) {
        if (item >= nblk) return;
        const float4 *src = p.sorted[side] + (size_t)b * pr_npad(n) + (size_t)item * PR_BLOCK;
#pragma unroll
        for (int e = 0; e < PR_BLOCK / 32; ++e) {
            const float4 t = src[e * 32 + lane];
            l[0] = fminf(l[0], t.x), h[0] = fmaxf(h[0], t.x);
            l[1] = fminf(l[1], t.y), h[1] = fmaxf(h[1], t.y);
            l[2] = fminf(l[2], t.z), h[2] = fmaxf(h[2], t.z);
        }
    } else {
        if (item >= nsb) return;
        const float4 *BL = p.boxes[side] + (size_t)b * 2 * nblk, *BH = BL + nblk;
#pragma unroll
        for (int e = 0; e < GR_SUPER / 32; ++e) {
            const int blk = item * GR_SUPER + e * 32 + lane;
            if (blk < nblk) {
                const float4 a = BL[blk], c = BH[blk];
                l[0] = fminf(l[0], a.x), l[1] = fminf(l[1], a.y), l[2] = fminf(l[2], a.z);
                h[0] = fmaxf(h[0], c.x), h[1] = fmaxf(h[1], c.y), h[2] = fmaxf(h[2], c.z);
            }
        }
    }
#pragma unroll
    for (int c = 0; c < 3; ++c)
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) {
            l[c] = fminf(l[c], __shfl_xor_sync(0xffffffffu, l[c], o));
            h[c] = fmaxf(h[c], __shfl_xor_sync(0xffffffffu, h[c], o));
        }
    if (lane == 0) {
        float4 *out = SUPER ? p.sboxes[side] + (size_t)b * 2 * nsb : p.boxes[side] + (size_t)b * 2 * nblk;
        const int cnt = SUPER ? nsb : nblk;
        out[item] = make_float4(l[0], l[1], l[2], 0.f), out[cnt + item] = make_float4(h[0], h[1], h[2], 0.f);
    }
}

struct Prune2Params {
    const float4 *q, *t;       // sorted queries [B][npad_q], sorted targets [B][npad_t]
    const float4 *tbox, *tsbox;
    unsigned long long *out;   // [B][nq] packed (dist, original target index), addressed by the query's original index
    int nq, nt, B;
    const int *select;
    unsigned *stats;           // optional [4]: blocks scanned, tie passes, groups, superblocks opened
    // probe launch (PROBE = true): only every probe_stride-th group runs, gives up after probe_cap block visits and adds its
    // visits to *probe_acc -- grid_decide_kernel turns the sum into the selection flag before the real launches
    int probe_stride, probe_cap;
    int *probe_acc;
};

// <<<1, 1>>>: the pruned scan only pays while a query group visits a small share of the target blocks; the probe measured it
// on a sample.  ctl[1] |= 1 (exhaustive kernels) when the sampled groups visited more than `limit` blocks in all; the
// accumulator ctl[2] is left zero.
static __global__ void grid_decide_kernel(int *ctl, int limit) {
    const int v = atomicExch(ctl + 2, 0);
    if (v > limit) ctl[1] = 1;
}

__device__ __forceinline__ float gr_box_dist(const float4 &lo, const float4 &hi, const float (&glo)[3], const float (&ghi)[3]) {
    const float dx = fmaxf(fmaxf(lo.x - ghi[0], glo[0] - hi.x), 0.f);
    const float dy = fmaxf(fmaxf(lo.y - ghi[1], glo[1] - hi.y), 0.f);
    const float dz = fmaxf(fmaxf(lo.z - ghi[2], glo[2] - hi.z), 0.f);
    return __fmul_rn(__fmaf_rn(dz, dz, __fmaf_rn(dx, dx, __fmul_rn(dy, dy))), 0.99999f);
}

// grid (ceil(groups / 8), B); SBR * 32 >= number of superblocks of the target cloud
template <int SBR, bool PROBE = false>
__global__ void __launch_bounds__(PR_THREADS) nn_prune2_kernel(const Prune2Params p) {
    __shared__ __align__(16) float stage[PR_THREADS / 32][3][PR_BLOCK];   // the block being walked, SoA: x[64] y[64] z[64]
    if (p.select != nullptr && *p.select != 0) return;
    const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
    const int groups = (p.nq + PR_GROUP - 1) / PR_GROUP;
    const int b = blockIdx.y;
    const int g = ((int)blockIdx.x * (PR_THREADS / 32) + wid) * (PROBE ? p.probe_stride : 1);
    if (g >= groups) return;
    const int nblk = pr_nblk(p.nt), nsb = gr_nsb(p.nt);
    const float4 *T = p.t + (size_t)b * pr_npad(p.nt);
    const float4 *BL = p.tbox + (size_t)b * 2 * nblk, *BH = BL + nblk;
    const float4 *SL = p.tsbox + (size_t)b * 2 * nsb, *SH = SL + nsb;
    const float inf = __int_as_float(0x7f800000);
    const int qi = g * PR_GROUP + lane;
    const bool valid = qi < p.nq;
    const float4 q = p.q[(size_t)b * pr_npad(p.nq) + qi];
    float glo[3], ghi[3];
    {
        const float v[3] = {q.x, q.y, q.z};
#pragma unroll
        for (int c = 0; c < 3; ++c) {
            const float a = valid ? v[c] : inf, z = valid ? v[c] : -inf;
            glo[c] = pr_ord2f(__reduce_min_sync(0xffffffffu, pr_f2ord(a)));
            ghi[c] = pr_ord2f(__reduce_max_sync(0xffffffffu, pr_f2ord(z)));
        }
    }
    float sbd[SBR];
#pragma unroll
    for (int r = 0; r < SBR; ++r) {
        const int sb = r * 32 + lane;
        sbd[r] = sb < nsb ? gr_box_dist(__ldg(SL + sb), __ldg(SH + sb), glo, ghi) : inf;
    }
    const float2 nx = make_float2(-q.x, -q.x), ny = make_float2(-q.y, -q.y), nz = make_float2(-q.z, -q.z);
    float best = inf;
    int bchunk = 0;
    bool tie = false;
    float thr = inf;
    unsigned scanned = 0, opened = 0;
    for (;;) {
        // nearest remaining superblock
        unsigned skey = 0xffffffffu;
#pragma unroll
        for (int r = 0; r < SBR; ++r) skey = min(skey, (__float_as_uint(sbd[r]) & 0xfffffe00u) | (unsigned)(r * 32 + lane));
        skey = __reduce_min_sync(0xffffffffu, skey);
        if (skey >= 0x7f800000u || !(__uint_as_float(skey & 0xfffffe00u) <= thr)) break;
        const int sb = (int)(skey & 0x1ffu);
#pragma unroll
        for (int r = 0; r < SBR; ++r)
            if (r * 32 + lane == sb) sbd[r] = inf;
        ++opened;
        // its blocks, nearest first
        float bd[GR_SUPER / 32];
#pragma unroll
        for (int r = 0; r < GR_SUPER / 32; ++r) {
            const int blk = sb * GR_SUPER + r * 32 + lane;
            bd[r] = blk < nblk ? gr_box_dist(__ldg(BL + blk), __ldg(BH + blk), glo, ghi) : inf;
        }
        for (;;) {
            unsigned key = 0xffffffffu;
#pragma unroll
            for (int r = 0; r < GR_SUPER / 32; ++r) key = min(key, (__float_as_uint(bd[r]) & 0xffffffc0u) | (unsigned)(r * 32 + lane));
            key = __reduce_min_sync(0xffffffffu, key);
            if (key >= 0x7f800000u || !(__uint_as_float(key & 0xffffffc0u) <= thr)) break;
            const int bl = (int)(key & 0x3fu);
#pragma unroll
            for (int r = 0; r < GR_SUPER / 32; ++r)
                if (r * 32 + lane == bl) bd[r] = inf;
            const int blk = sb * GR_SUPER + bl;
            // the warp fetches the block with ONE coalesced 1 KB read (a uniform load per target would be an L2 round trip
            // per 8-target chunk: measured 0.15 instructions per cycle and SM) and walks it in shared memory
            {
                const float4 *tg = T + (size_t)blk * PR_BLOCK;
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
            thr = pr_ord2f(__reduce_max_sync(0xffffffffu, valid ? pr_f2ord(best) : (int)0x80000000));
            if (PROBE && scanned >= (unsigned)p.probe_cap) break;
        }
        if (PROBE && scanned >= (unsigned)p.probe_cap) break;
    }
    if (PROBE) {
        if (lane == 0) atomicAdd(p.probe_acc, (int)scanned);
        return;
    }
    // lowest original index at the minimum: inside the winning chunk ...
    int bidx = 0x7fffffff;
    {
        const float4 *tc = T + (size_t)bchunk * 8;
#pragma unroll
        for (int i = 0; i < 8; ++i) {
            const float4 t = __ldg(tc + i);
            const float d = sqdist_ref(q.x, q.y, q.z, t.x, t.y, t.z);
            if (d == best) bidx = min(bidx, __float_as_int(t.w));
        }
    }
    // ... and, for a lane that met its minimum in two chunks, over every block that could hold it
    const unsigned ties = __ballot_sync(0xffffffffu, tie && valid);
    if (ties != 0u) {
        for (int sb = 0; sb < nsb; ++sb) {
            if (!(gr_box_dist(__ldg(SL + sb), __ldg(SH + sb), glo, ghi) <= thr)) continue;   // warp-uniform
            for (int blk = sb * GR_SUPER; blk < min((sb + 1) * GR_SUPER, nblk); ++blk) {
                if (!(gr_box_dist(__ldg(BL + blk), __ldg(BH + blk), glo, ghi) <= thr)) continue;
                const float4 *tb = T + (size_t)blk * PR_BLOCK;
                for (int i = 0; i < PR_BLOCK; ++i) {
                    const float4 t = __ldg(tb + i);
                    const float d = sqdist_ref(q.x, q.y, q.z, t.x, t.y, t.z);
                    if (tie && d == best) bidx = min(bidx, __float_as_int(t.w));
                }
            }
        }
    }
    if (valid) p.out[(size_t)b * p.nq + __float_as_int(q.w)] = pack_dist_idx(best, bidx);
    if (p.stats != nullptr && lane == 0) {
        atomicAdd(p.stats + 0, scanned);
        if (ties != 0u) atomicAdd(p.stats + 1, 1u);
        atomicAdd(p.stats + 2, 1u);
        atomicAdd(p.stats + 3, opened);
    }
}

}  // namespace genpc
