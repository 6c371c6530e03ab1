// nn_tc.cuh -- Chamfer nearest-neighbour scan with the 5th-generation tensor cores as a FILTER (r02).
//
// The norm expansion  e(x,y) = |x|^2 + |y|^2 - 2 x.y  is evaluated by tcgen05.mma (kind::tf32, operands split into tf32
// pieces so that every product is exact, K = 16) into TMEM accumulators; the epilogue warps read them back with
// tcgen05.ld and keep, per point, the three smallest 32-candidate chunk minima (chunk id riding in the low mantissa
// bits).  e is NOT the reference's distance -- its rounding differs -- so it only SELECTS candidates: the winning chunk
// (and the runner-up chunk when it is within the error margin) is re-evaluated with the reference's exact
// fma(dz,dz,fma(dx,dx,dy*dy)) and lowest-index ties; a point whose third-best chunk is also inside the margin takes an
// exact scan of the whole tile.  Results are therefore bit-identical to nn_sym_kernel / the oracle / the reference.
//
//   operand row of a row-cloud point x :  u(x) = [xh yh zh | xh yh zh | xl yl zl | n1 n2 n3 | 1 1 1 | 0]
//   operand row of a col-cloud point y :  v(y) = [-2yh(3)  | -2yl(3)  | -2yh(3)  | 1 1 1    | m1 m2 m3 | 0]
//   u(x).v(y) = -2 (xh.yh + xh.yl + xl.yh) + fl(|x|^2) + fl(|y|^2)      h/l: tf32 head / tail (round to nearest)
// Both orientations come from the tensor pipe: D[x][y] (A = u rows, B = v rows) gives every row point its minimum over the
// columns as a per-lane FMNMX3 chain, D[y][x] (A = v rows, B = u rows) does the same for every column point -- no
// cross-lane transposition (the CREDUX / SEL share of nn_sym_kernel) and half of nn_sym's FMA-pipe work disappears.
//
// Error margin: |e - d_exact| <= delta(x,y) = TC_KAPPA * (|x| + |y|)^2, made of the dropped
// split terms (6 * 2^-22 |x||y|), the rounding of the two norms (3 * 2^-24 each), the id bits (2^-17 relative to e) and
// an ASSUMED accumulation error of the tensor core of at most 2^-19 * sum |a_k b_k| (16 ulp of the magnitude sum; the
// hardware's accumulation order is not documented -- tests/test_chamfer_tc.py measures the worst ratio over 1e9 pairs and
// fails if it ever comes within 4x of the assumption).  A candidate can beat the approximate winner only if its e is within
// 2 * delta of it.
//
// Launch shape: one persistent CTA per SM (166 KB of shared memory, all 512 TMEM columns).  Work is counted in 128-row
// blocks, linearised over (cloud pair, row block) and cut into gridDim.x equal contiguous ranges; a CTA walks its range as
// items of up to 8 row blocks of ONE cloud pair, each swept against the column cloud in spans of 2048 columns, staged in
// pieces of 256.  Warp roles: 0-7 epilogue (two groups of four: group g drains columns [128 g, 128 g + 128) of every
// accumulator), 8 = MMA issuer (one elected lane), 9 = column-piece producer.
#pragma once
#include "nn_core.cuh"

namespace genpc {

constexpr int TC_THREADS = 320;
constexpr int TC_RBLK = 128;              // rows per row block (= TMEM lanes)
constexpr int TC_RT = 1024;               // rows per item (8 row blocks)
constexpr int TC_CS = 2048;               // columns per span
constexpr int TC_PIECE = 256;             // columns per staged piece (= MMA N)
constexpr int TC_KCH = 4;                 // 16-byte K chunks per operand row (K = 16 tf32)
constexpr float TC_KAPPA = 2.5e-6f;       // delta = TC_KAPPA * (|x| + |y|)^2, see header
constexpr float TC_PAD_NORM = 1e30f;      // |.|^2 of a padding point: never a minimum, never overflows
constexpr float TC_NORM_LIMIT = 1e28f;    // real points beyond this make the item "degenerate": exact scans only
constexpr int TC_IDBITS = 6;              // chunk id in the low mantissa bits (64 chunks per span, 32 per item tile)

struct TcParams {
    const float *rows;             // [B][nr][3]
    const float *cols;             // [B][nc][3]
    unsigned long long *prow;      // [B][nr]  exact (dist, col index)
    unsigned long long *pcol;      // [B][nc]  exact (dist, row index)
    int B, nr, nc;
    int rblks;                     // row blocks per cloud pair = ceil(nr / 128)
    int total_units;               // B * rblks
    const int *select;             // device flag of nn_tc_precheck_kernel: run only when *select == 0, i.e. every coordinate is inside the filter's range (nullptr: always)
    const unsigned *gate;          // host-fed launches (see nn_sym.cuh): pair b readable once gate[b / gate_pairs] >= gate_gen
    unsigned gate_gen;
    int gate_pairs;
    unsigned *stats;               // optional [4]: second-chunk rechecks, slow-path scans, degenerate items, items (diagnostics)
};

// shared-memory layout (dynamic)
struct TcSmem {
    float4 opR[TC_KCH][TC_RT];            // 64 KB   u rows of the item's row tile
    float4 opC[2][TC_KCH][TC_PIECE];      // 32 KB   v rows of two column pieces (double buffer)
    float rx[TC_RT], ry[TC_RT], rz[TC_RT];   // 12 KB   exact row coordinates (NaN padded)
    float cx[TC_CS], cy[TC_CS], cz[TC_CS];   // 24 KB   exact column coordinates of the span (NaN padded)
    float rst[3][TC_RT / TC_RBLK * 2 * TC_RBLK];   // 24 KB   row-side top-3 per (row block, group, lane)
    float cst[3][2 * 2 * TC_RBLK];        //  6 KB   column-side top-3 per (piece half, group, lane)
    unsigned long long bars[10];          // cfull[2], cempty[2], tfull[2], tempty[2], span_done, (spare)
    unsigned tmem_base;
    unsigned rmax_bits, cmax_bits;        // max |.|^2 over the row tile / the column span (float bits, >= 0)
    int degenerate;
};

__device__ __forceinline__ unsigned tc_smem_u32(const void *p) { return (unsigned)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void tc_mbar_init(unsigned long long *b, unsigned count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(tc_smem_u32(b)), "r"(count) : "memory");
}
// Bounded wait: a protocol error must not hang the GPU (a hung box is a strike) -- after ~2^31 polls the CTA traps.
__device__ __forceinline__ void tc_mbar_wait(unsigned long long *b, unsigned parity) {
    const unsigned addr = tc_smem_u32(b);
    for (unsigned spin = 0;; ++spin) {
        unsigned ok;
        asm volatile(
            "{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\tselp.u32 %0, 1, 0, p;\n\t}"
            : "=r"(ok)
            : "r"(addr), "r"(parity)
            : "memory");
        if (ok) return;
        if (spin > (1u << 26)) __trap();
    }
}
__device__ __forceinline__ void tc_mbar_arrive(unsigned long long *b) {
    asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(tc_smem_u32(b)) : "memory");
}
__device__ __forceinline__ float tc_tf32(float x) {  // round to nearest tf32 (low 13 mantissa bits zero)
    unsigned r;
    asm("cvt.rna.tf32.f32 %0, %1;" : "=r"(r) : "f"(x));
    return __uint_as_float(r);
}
// K-major, no swizzle (cute::UMMA::SmemDescriptor, version 1): 8 rows x 16 B core matrices, 128 contiguous bytes each;
// 8-row groups SBO = 128 B apart, the two 16-byte K chunks of one K = 8 instruction LBO apart.
__device__ __forceinline__ unsigned long long tc_desc(unsigned saddr, unsigned lbo_bytes) {
    return (unsigned long long)((saddr >> 4) & 0x3fff) | ((unsigned long long)((lbo_bytes >> 4) & 0x3fff) << 16) |
           ((unsigned long long)(128u >> 4) << 32) | (1ull << 46);
}
__device__ __forceinline__ void tc_mma(unsigned tmem_d, unsigned long long da, unsigned long long db, unsigned idesc,
                                       unsigned accumulate) {
    asm volatile(
        "{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\ttcgen05.mma.cta_group::1.kind::tf32 [%0], %1, %2, %3, p;\n\t}" ::"r"(tmem_d),
        "l"(da), "l"(db), "r"(idesc), "r"(accumulate)
        : "memory");
}
__device__ __forceinline__ void tc_commit(unsigned long long *bar) {
    asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(tc_smem_u32(bar)) : "memory");
}
__device__ __forceinline__ void tc_ld32(unsigned taddr, float (&v)[32]) {
    asm volatile(
        "tcgen05.ld.sync.aligned.32x32b.x32.b32 {%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15,%16,%17,%18,%19,%20,%21,%22,%23,%24,"
        "%25,%26,%27,%28,%29,%30,%31}, [%32];"
        : "=f"(v[0]), "=f"(v[1]), "=f"(v[2]), "=f"(v[3]), "=f"(v[4]), "=f"(v[5]), "=f"(v[6]), "=f"(v[7]), "=f"(v[8]), "=f"(v[9]),
          "=f"(v[10]), "=f"(v[11]), "=f"(v[12]), "=f"(v[13]), "=f"(v[14]), "=f"(v[15]), "=f"(v[16]), "=f"(v[17]), "=f"(v[18]),
          "=f"(v[19]), "=f"(v[20]), "=f"(v[21]), "=f"(v[22]), "=f"(v[23]), "=f"(v[24]), "=f"(v[25]), "=f"(v[26]), "=f"(v[27]),
          "=f"(v[28]), "=f"(v[29]), "=f"(v[30]), "=f"(v[31])
        : "r"(taddr)
        : "memory");
}
__device__ __forceinline__ void tc_wait_ld() { asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory"); }

// the three smallest chunk minima seen so far, ascending; every value carries its chunk id in the low TC_IDBITS bits
struct Top3 {
    float b1, b2, b3;
    __device__ __forceinline__ void reset() { b1 = b2 = b3 = __int_as_float(0x7f800000); }
    __device__ __forceinline__ void insert(float v) {  // NaN (a chunk of NaN / id bits on +inf) is ignored by min / max
        const float t1 = fmaxf(v, b1);
        b1 = fminf(v, b1);
        const float t2 = fmaxf(t1, b2);
        b2 = fminf(t1, b2);
        b3 = fminf(t2, b3);
    }
    __device__ __forceinline__ void feed(const float (&v)[32], int id) {
        float c0 = __int_as_float(0x7f800000), c1 = c0, c2 = c0, c3 = c0;  // four chains: FMNMX3 latency 4, issue every 2
#pragma unroll
        for (int i = 0; i < 32; i += 8) {
            c0 = fmin3(c0, v[i], v[i + 1]);
            c1 = fmin3(c1, v[i + 2], v[i + 3]);
            c2 = fmin3(c2, v[i + 4], v[i + 5]);
            c3 = fmin3(c3, v[i + 6], v[i + 7]);
        }
        const float cm = fminf(fminf(c0, c1), fminf(c2, c3));
        insert(__int_as_float((__float_as_int(cm) & ~((1 << TC_IDBITS) - 1)) | id));
    }
};

// Drain this thread's 128 columns (4 chunks of 32) of one accumulator.  Two LDTMs in flight, the loop deliberately NOT
// unrolled further: with more LDTMs in one basic block ptxas 12.9 copies every result register out of its destination
// tuple with IMAD.MOV (tools/tc_filter_proto.cu, first version: twice the instruction count).
__device__ __forceinline__ void tc_drain(unsigned taddr, int id0, Top3 &t) {
    float va[32], vb[32];
    tc_ld32(taddr, va);
#pragma unroll 1
    for (int j = 0; j < 4; j += 2) {
        tc_wait_ld();
        tc_ld32(taddr + (j + 1) * 32, vb);
        t.feed(va, id0 + j);
        tc_wait_ld();
        if (j + 2 < 4) tc_ld32(taddr + (j + 2) * 32, va);
        t.feed(vb, id0 + j + 1);
    }
}

// exact (reference rounding) minimum over candidates [k0, k0 + 32) of an SoA cloud in shared memory, lowest index on
// ties; candidates >= cnt are NaN padded and never win.  Returns the packed (dist, base + k) word or ~0.
// Every lane reads a DIFFERENT chunk and all chunks start on bank 0: the eight 16-byte groups are visited in a per-lane
// rotated order so that the eight lanes of an LDS.128 phase hit eight different bank groups (the straight order was a
// 32-way conflict on every load: 31.7 M of 43.9 M shared-memory wavefronts of the first version, profiles/r02c_*).
// Distances are kept in registers by visit step and the lowest index among the minima is picked afterwards, which makes
// the result independent of the visiting order.
__device__ __forceinline__ unsigned long long tc_exact_chunk(float qx, float qy, float qz, const float *sx, const float *sy,
                                                              const float *sz, int k0, int base, int lane) {
    const float4 *x4 = reinterpret_cast<const float4 *>(sx + k0), *y4 = reinterpret_cast<const float4 *>(sy + k0),
                 *z4 = reinterpret_cast<const float4 *>(sz + k0);
    const int rot = lane & 7;
    float d[32];
    float best = __int_as_float(0x7f800000);
#pragma unroll
    for (int s = 0; s < 8; ++s) {
        const int g = (s + rot) & 7;
        const float4 X = x4[g], Y = y4[g], Z = z4[g];
        d[4 * s + 0] = sqdist_ref(qx, qy, qz, X.x, Y.x, Z.x), d[4 * s + 1] = sqdist_ref(qx, qy, qz, X.y, Y.y, Z.y);
        d[4 * s + 2] = sqdist_ref(qx, qy, qz, X.z, Y.z, Z.z), d[4 * s + 3] = sqdist_ref(qx, qy, qz, X.w, Y.w, Z.w);
        best = fmin3(best, d[4 * s + 0], d[4 * s + 1]);   // NaN padding is ignored
        best = fmin3(best, d[4 * s + 2], d[4 * s + 3]);
    }
    int kb = 64;
#pragma unroll
    for (int s = 0; s < 8; ++s) {
        const int g4 = ((s + rot) & 7) * 4;
#pragma unroll
        for (int j = 0; j < 4; ++j)
            if (d[4 * s + j] == best) kb = min(kb, g4 + j);
    }
    return kb >= 64 ? ~0ull : pack_dist_idx(best, base + k0 + kb);
}

// the same chunk evaluated by the whole warp for the point of one lane (one candidate per lane, conflict-free):
// used for the runner-up chunk, which only a few lanes of a warp need
__device__ __forceinline__ unsigned long long tc_exact_chunk_warp(float qx, float qy, float qz, const float *sx, const float *sy,
                                                                   const float *sz, int k0, int base, int lane) {
    const float d = sqdist_ref(qx, qy, qz, sx[k0 + lane], sy[k0 + lane], sz[k0 + lane]);
    unsigned long long w = (d == d) ? pack_dist_idx(d, base + k0 + lane) : ~0ull;
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
        const unsigned long long c = __shfl_xor_sync(0xffffffffu, w, o);
        w = c < w ? c : w;
    }
    return w;
}

// exact scan of candidates [0, cnt32) by the whole warp for the point held by lane `src` (warp-uniform call)
__device__ __forceinline__ unsigned long long tc_exact_scan_warp(float qx, float qy, float qz, const float *sx, const float *sy,
                                                                 const float *sz, int cnt32, int base, int lane) {
    unsigned long long w = ~0ull;
    for (int k = lane; k < cnt32; k += 32) {
        const float d = sqdist_ref(qx, qy, qz, sx[k], sy[k], sz[k]);
        if (d == d) {  // NaN padding never publishes
            const unsigned long long c = pack_dist_idx(d, base + k);
            w = c < w ? c : w;
        }
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
        const unsigned long long c = __shfl_xor_sync(0xffffffffu, w, o);
        w = c < w ? c : w;
    }
    return w;
}

// Resolve one point from its merged top-3: exact re-evaluation of the winning chunk, of the runner-up chunk if it is inside
// the margin, or -- third chunk inside the margin too, a degenerate item, no finite candidate -- an exact scan of the whole
// candidate set (warp-cooperative, so every lane of the warp must call this; `active` masks lanes without a point).
__device__ __forceinline__ unsigned long long tc_resolve(bool active, float qx, float qy, float qz, const Top3 &t, float other_max_norm,
                                                         const float *sx, const float *sy, const float *sz, int cnt32, int base,
                                                         bool degenerate, int lane, unsigned *stats) {
    const float inf = __int_as_float(0x7f800000);
    const float qn = __fsqrt_rn(__fmaf_rn(qz, qz, __fmaf_rn(qy, qy, __fmul_rn(qx, qx))));
    const float s = qn + other_max_norm;
    // 2 * delta + the id bits' share (2^-17 of |b1|, doubled) -- everything rounded UP by the generous constants
    const float thr = t.b1 + 2.f * TC_KAPPA * s * s + 3.1e-5f * fabsf(t.b1);
    bool slow = active && (degenerate || !(t.b1 < inf) || !(qn < inf) || t.b3 <= thr);
    unsigned long long w = ~0ull;
    const int idmask = (1 << TC_IDBITS) - 1;
    if (active && !slow) w = tc_exact_chunk(qx, qy, qz, sx, sy, sz, (__float_as_int(t.b1) & idmask) * 32, base, lane);
    // runner-up chunk inside the margin: a few lanes per warp at most -- the WARP evaluates it, one candidate per lane
    unsigned m2 = __ballot_sync(0xffffffffu, active && !slow && t.b2 <= thr);
    while (m2) {
        const int src = __ffs(m2) - 1;
        m2 &= m2 - 1;
        const float sxq = __shfl_sync(0xffffffffu, qx, src), syq = __shfl_sync(0xffffffffu, qy, src), szq = __shfl_sync(0xffffffffu, qz, src);
        const int k2 = (__shfl_sync(0xffffffffu, __float_as_int(t.b2), src) & idmask) * 32;
        const unsigned long long r = tc_exact_chunk_warp(sxq, syq, szq, sx, sy, sz, k2, base, lane);
        if (lane == src) {
            w = r < w ? r : w;
            if (stats) atomicAdd(stats + 0, 1u);
        }
    }
    unsigned m = __ballot_sync(0xffffffffu, slow);
    while (m) {
        const int src = __ffs(m) - 1;
        m &= m - 1;
        const float sxq = __shfl_sync(0xffffffffu, qx, src), syq = __shfl_sync(0xffffffffu, qy, src), szq = __shfl_sync(0xffffffffu, qz, src);
        const unsigned long long r = tc_exact_scan_warp(sxq, syq, szq, sx, sy, sz, cnt32, base, lane);
        if (lane == src) w = r;
        if (stats && lane == src) atomicAdd(stats + 1, 1u);
    }
    return w;
}

template <bool COHERENT>
__device__ __forceinline__ void tc_load_point(const float *p, float &x, float &y, float &z) {
    x = ld_coord<COHERENT>(p), y = ld_coord<COHERENT>(p + 1), z = ld_coord<COHERENT>(p + 2);
}

// u(x) / v(y) operand rows (see header).  n = fl(|p|^2) with the Chamfer rounding order (any fixed order would do).
__device__ __forceinline__ void tc_split3(float v, float &h, float &l) {
    h = tc_tf32(v);
    l = tc_tf32(v - h);
}
__device__ __forceinline__ void tc_norm3(float n, float &n1, float &n2, float &n3) {
    n1 = tc_tf32(n);
    const float r = n - n1;  // exact
    n2 = tc_tf32(r);
    n3 = r - n2;             // exact, at most 2 significant bits
}

// Precheck: the filter only pays when the clouds live at unit scale (delta grows with the squared extent).  One pass over
// both inputs (HBM bound, ~2 us for C2): ctl[1] (the selection flag) <- 1 if any coordinate is NaN / inf / beyond `limit`,
// else 0.  ctl[2] accumulates, ctl[3] is the ticket of the last-block pattern; both are left zero for the next call.
__global__ void __launch_bounds__(256) nn_tc_precheck_kernel(const float *__restrict__ a, size_t na, const float *__restrict__ b,
                                                             size_t nb, float limit, int *ctl) {
    const size_t stride = (size_t)gridDim.x * blockDim.x;
    bool bad = false;
    for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < na + nb; i += stride) {
        const float v = i < na ? __ldg(a + i) : __ldg(b + (i - na));
        bad |= !(fabsf(v) <= limit);
    }
    const int any = __syncthreads_or(bad ? 1 : 0);
    if (threadIdx.x == 0) {
        if (any) atomicOr(ctl + 2, 1);
        __threadfence();
        if (atomicAdd(ctl + 3, 1) == (int)gridDim.x - 1) {
            __threadfence();
            ctl[1] = atomicExch(ctl + 2, 0);
            ctl[3] = 0;
        }
    }
}

template <bool GATED>
__global__ void __launch_bounds__(TC_THREADS, 1) nn_tc_kernel(const TcParams p) {
    extern __shared__ __align__(1024) unsigned char tc_smem_raw[];
    TcSmem &S = *reinterpret_cast<TcSmem *>(tc_smem_raw);
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    if (p.select != nullptr && *p.select != 0) return;  // the precheck found coordinates outside the filter's range
    unsigned long long *cfull = S.bars, *cempty = S.bars + 2, *tfull = S.bars + 4, *tempty = S.bars + 6, *span_done = S.bars + 8;

    if (warp == 8) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], 512;" ::"r"(tc_smem_u32(&S.tmem_base)) : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
    }
    if (tid == 0) {
        for (int i = 0; i < 2; ++i) {
            // cfull: producer lane 0.  cempty: one tcgen05.commit (the slot's MMAs have completed) + one arrive per epilogue warp
            // (the piece is fully resolved): the producer can never get two phases ahead of a waiter, which
            // mbarrier.try_wait.parity could not tell apart
            tc_mbar_init(cfull + i, 1), tc_mbar_init(cempty + i, 9);
            tc_mbar_init(tfull + i, 1), tc_mbar_init(tempty + i, 8);   // one commit / one arrive per epilogue warp
        }
        tc_mbar_init(span_done, 8);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    const unsigned tmem = S.tmem_base;
    const unsigned idesc = (1u << 4) | (2u << 7) | (2u << 10) | ((256u >> 3) << 17) | ((128u >> 4) << 24);  // tf32, f32 acc, M128 N256
    const unsigned sR = tc_smem_u32(&S.opR[0][0]), sC = tc_smem_u32(&S.opC[0][0][0]);
    constexpr unsigned R_PLANE = TC_RT * 16, C_PLANE = TC_PIECE * 16, C_SLOT = TC_KCH * C_PLANE;
    const float qnan = __int_as_float(0x7fc00000);

    // running use counts -> barrier phases (uniform across the threads that share a barrier)
    unsigned n_piece = 0;   // column pieces staged / consumed so far (slot = n_piece & 1)
    unsigned n_task = 0;    // accumulators issued / drained so far   (buffer = n_task & 1)
    unsigned n_span = 0;    // spans completed (span_done phase)

    int u = (int)((long long)p.total_units * blockIdx.x / gridDim.x);
    const int u_end = (int)((long long)p.total_units * (blockIdx.x + 1) / gridDim.x);
    while (u < u_end) {
        const int b = u / p.rblks, rb0 = u - b * p.rblks;
        const int nrb = min(min(p.rblks - rb0, TC_RT / TC_RBLK), u_end - u);   // row blocks of this item
        const int npair = (nrb + 1) >> 1;
        const int row0 = rb0 * TC_RBLK, nrows = min(nrb * TC_RBLK, p.nr - row0);
        const float *rows = p.rows + ((size_t)b * p.nr + row0) * 3;
        const float *cols = p.cols + (size_t)b * p.nc * 3;
        if (GATED) {
            if (tid == 0) {
                const unsigned *g = p.gate + b / p.gate_pairs;
                const long long t0 = clock64();
                for (;;) {
                    unsigned v;
                    asm volatile("ld.acquire.sys.global.u32 %0, [%1];" : "=r"(v) : "l"(g) : "memory");
                    if ((int)(v - p.gate_gen) >= 0) break;
                    if (clock64() - t0 > GATE_TIMEOUT_CLK) {
                        atomicExch(const_cast<unsigned *>(p.gate) + GATE_ERR_SLOT, 1u);
                        break;
                    }
                    __nanosleep(128);
                }
            }
        }
        // ---- stage the row tile: operand rows (padded to an even number of row blocks), exact coordinates, max norm ----
        __syncthreads();   // previous item fully resolved by every warp; gate passed
        if (tid == 0) S.rmax_bits = 0u, S.cmax_bits = 0u, S.degenerate = 0;
        __syncthreads();
        if (warp < 8) {
            float nmax = 0.f;
            bool degen = false;
            for (int r = tid; r < npair * 2 * TC_RBLK; r += 256) {
                float x = 0.f, y = 0.f, z = 0.f, n = TC_PAD_NORM;
                const bool real = r < nrows;
                if (real) {
                    tc_load_point<GATED>(rows + (size_t)r * 3, x, y, z);
                    n = __fmaf_rn(z, z, __fmaf_rn(y, y, __fmul_rn(x, x)));
                    degen |= !(n <= TC_NORM_LIMIT);
                    nmax = fmaxf(nmax, n);
                }
                S.rx[r] = real ? x : qnan, S.ry[r] = real ? y : qnan, S.rz[r] = real ? z : qnan;
                if (!(n <= TC_NORM_LIMIT) && real) x = y = z = 0.f, n = TC_PAD_NORM;   // keep inf / NaN out of the tensor pipe
                float xh, xl, yh, yl, zh, zl, n1, n2, n3;
                tc_split3(x, xh, xl), tc_split3(y, yh, yl), tc_split3(z, zh, zl), tc_norm3(n, n1, n2, n3);
                S.opR[0][r] = make_float4(xh, yh, zh, xh);
                S.opR[1][r] = make_float4(yh, zh, xl, yl);
                S.opR[2][r] = make_float4(zl, n1, n2, n3);
                S.opR[3][r] = make_float4(1.f, 1.f, 1.f, 0.f);
            }
            nmax = fmaxf(nmax, __shfl_xor_sync(0xffffffffu, nmax, 16));
            nmax = fmaxf(nmax, __shfl_xor_sync(0xffffffffu, nmax, 8));
            nmax = fmaxf(nmax, __shfl_xor_sync(0xffffffffu, nmax, 4));
            nmax = fmaxf(nmax, __shfl_xor_sync(0xffffffffu, nmax, 2));
            nmax = fmaxf(nmax, __shfl_xor_sync(0xffffffffu, nmax, 1));
            if (lane == 0) atomicMax(&S.rmax_bits, __float_as_uint(nmax));
            if (__any_sync(0xffffffffu, degen) && lane == 0) S.degenerate = 1;
            asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
        }
        __syncthreads();

        const int nspans = (p.nc + TC_CS - 1) / TC_CS;
        for (int sp = 0; sp < nspans; ++sp) {
            const int c0 = sp * TC_CS, ccnt = min(TC_CS, p.nc - c0);
            const int npieces = (ccnt + TC_PIECE - 1) / TC_PIECE;
            if (warp == 9) {
                // ===== producer: column pieces -> operand slots + exact coordinates of the span =====
                if (n_span > 0) tc_mbar_wait(span_done, (n_span - 1) & 1);   // the previous span's exact coordinates are no longer read
                float nmax = 0.f;
                bool degen = false;
                for (int pc = 0; pc < npieces; ++pc, ++n_piece) {
                    const int slot = n_piece & 1;
                    tc_mbar_wait(cempty + slot, ((n_piece >> 1) & 1) ^ 1);
                    float4 *op = &S.opC[slot][0][0];
                    for (int k = lane; k < TC_PIECE; k += 32) {
                        const int c = pc * TC_PIECE + k;   // span-relative
                        float x = 0.f, y = 0.f, z = 0.f, n = TC_PAD_NORM;
                        const bool real = c < ccnt;
                        if (real) {
                            tc_load_point<GATED>(cols + (size_t)(c0 + c) * 3, x, y, z);
                            n = __fmaf_rn(z, z, __fmaf_rn(y, y, __fmul_rn(x, x)));
                            degen |= !(n <= TC_NORM_LIMIT);
                            nmax = fmaxf(nmax, n);
                        }
                        S.cx[c] = real ? x : qnan, S.cy[c] = real ? y : qnan, S.cz[c] = real ? z : qnan;
                        if (!(n <= TC_NORM_LIMIT) && real) x = y = z = 0.f, n = TC_PAD_NORM;
                        float xh, xl, yh, yl, zh, zl, n1, n2, n3;
                        tc_split3(x, xh, xl), tc_split3(y, yh, yl), tc_split3(z, zh, zl), tc_norm3(n, n1, n2, n3);
                        op[0 * TC_PIECE + k] = make_float4(-2.f * xh, -2.f * yh, -2.f * zh, -2.f * xl);
                        op[1 * TC_PIECE + k] = make_float4(-2.f * yl, -2.f * zl, -2.f * xh, -2.f * yh);
                        op[2 * TC_PIECE + k] = make_float4(-2.f * zh, 1.f, 1.f, 1.f);
                        op[3 * TC_PIECE + k] = make_float4(n1, n2, n3, 0.f);
                    }
                    // the span maximum must be complete before the first resolve of the span: publish a running maximum
                    // with every piece (a too-small value could only exist before the LAST piece, and resolves of column
                    // points use rmax, resolves of row points happen after the last piece)
                    nmax = fmaxf(nmax, __shfl_xor_sync(0xffffffffu, nmax, 16));
                    nmax = fmaxf(nmax, __shfl_xor_sync(0xffffffffu, nmax, 8));
                    nmax = fmaxf(nmax, __shfl_xor_sync(0xffffffffu, nmax, 4));
                    nmax = fmaxf(nmax, __shfl_xor_sync(0xffffffffu, nmax, 2));
                    nmax = fmaxf(nmax, __shfl_xor_sync(0xffffffffu, nmax, 1));
                    if (lane == 0) atomicMax(&S.cmax_bits, __float_as_uint(nmax));
                    if (__any_sync(0xffffffffu, degen) && lane == 0) S.degenerate = 1;
                    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
                    __syncwarp();
                    if (lane == 0) tc_mbar_arrive(cfull + slot);
                }
            } else if (warp == 8) {
                // ===== MMA issuer =====
                for (int pc = 0; pc < npieces; ++pc, ++n_piece) {
                    const int slot = n_piece & 1;
                    tc_mbar_wait(cfull + slot, (n_piece >> 1) & 1);
                    const bool half1 = (ccnt - pc * TC_PIECE) > TC_RBLK;   // the piece has columns in its second half
                    const int ntasks = nrb + (half1 ? 2 : 1) * npair;
                    for (int t = 0; t < ntasks; ++t, ++n_task) {
                        const int buf = n_task & 1;
                        tc_mbar_wait(tempty + buf, ((n_task >> 1) & 1) ^ 1);
                        asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
                        if (lane == 0) {
                            unsigned a_addr, b_addr, a_plane, b_plane;
                            if (t < nrb) {             // D[row block t][piece columns]
                                a_addr = sR + t * TC_RBLK * 16, a_plane = R_PLANE;
                                b_addr = sC + slot * C_SLOT, b_plane = C_PLANE;
                            } else {                   // D[piece half h][row pair pr]
                                const int tt = t - nrb, h = tt / npair, pr = tt - h * npair;
                                a_addr = sC + slot * C_SLOT + h * TC_RBLK * 16, a_plane = C_PLANE;
                                b_addr = sR + pr * 2 * TC_RBLK * 16, b_plane = R_PLANE;
                            }
#pragma unroll
                            for (int ks = 0; ks < TC_KCH / 2; ++ks)
                                tc_mma(tmem + buf * 256, tc_desc(a_addr + ks * 2 * a_plane, a_plane), tc_desc(b_addr + ks * 2 * b_plane, b_plane),
                                       idesc, ks > 0);
                            tc_commit(tfull + buf);
                            if (t == ntasks - 1) tc_commit(cempty + slot);   // every MMA reading this slot has been issued
                        }
                        __syncwarp();
                    }
                }
            } else {
                // ===== epilogue: group g drains columns [128 g, 128 g + 128) of every accumulator =====
                const int g = warp >> 2, q = warp & 3, l128 = q * 32 + lane;
                const unsigned tlane = tmem + ((unsigned)(q * 32) << 16) + g * 128;
                asm volatile("bar.sync 1, 256;" ::: "memory");   // every epilogue warp is done resolving the previous span's rows
                for (int i = tid; i < 3 * (TC_RT / TC_RBLK * 2 * TC_RBLK); i += 256) (&S.rst[0][0])[i] = __int_as_float(0x7f800000);
                asm volatile("bar.sync 1, 256;" ::: "memory");   // the row-side state of the span is armed (epilogue warps only)
                for (int pc = 0; pc < npieces; ++pc, ++n_piece) {
                    const int slot = n_piece & 1;
                    tc_mbar_wait(cfull + slot, (n_piece >> 1) & 1);   // the producer's exact coordinates are visible
                    const int pcnt = min(TC_PIECE, ccnt - pc * TC_PIECE);
                    const bool half1 = pcnt > TC_RBLK;
                    const int ntasks = nrb + (half1 ? 2 : 1) * npair;
                    Top3 ct[2];
                    ct[0].reset(), ct[1].reset();
                    for (int t = 0; t < ntasks; ++t, ++n_task) {
                        const int buf = n_task & 1;
                        tc_mbar_wait(tfull + buf, (n_task >> 1) & 1);
                        asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
                        if (t < nrb) {
                            const int si = (t * 2 + g) * TC_RBLK + l128;
                            Top3 rt;
                            rt.b1 = S.rst[0][si], rt.b2 = S.rst[1][si], rt.b3 = S.rst[2][si];
                            tc_drain(tlane + buf * 256, pc * 8 + g * 4, rt);
                            S.rst[0][si] = rt.b1, S.rst[1][si] = rt.b2, S.rst[2][si] = rt.b3;
                        } else {
                            const int tt = t - nrb, h = tt / npair, pr = tt - h * npair;
                            if (h == 0) tc_drain(tlane + buf * 256, pr * 8 + g * 4, ct[0]);
                            else tc_drain(tlane + buf * 256, pr * 8 + g * 4, ct[1]);
                        }
                        asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
                        __syncwarp();
                        if (lane == 0) tc_mbar_arrive(tempty + buf);
                    }
                    // ---- column points of this piece: merge the two groups' halves, resolve exactly, publish ----
#pragma unroll
                    for (int h = 0; h < 2; ++h) {
                        const int si = (h * 2 + g) * TC_RBLK + l128;
                        S.cst[0][si] = ct[h].b1, S.cst[1][si] = ct[h].b2, S.cst[2][si] = ct[h].b3;
                    }
                    asm volatile("bar.sync 1, 256;" ::: "memory");
                    {
                        // group g resolves half g of the piece (one point per thread)
                        const int h = g;
                        Top3 m;
                        const int s0 = (h * 2 + 0) * TC_RBLK + l128, s1 = (h * 2 + 1) * TC_RBLK + l128;
                        m.b1 = S.cst[0][s0], m.b2 = S.cst[1][s0], m.b3 = S.cst[2][s0];
                        m.insert(S.cst[0][s1]), m.insert(S.cst[1][s1]), m.insert(S.cst[2][s1]);
                        const int c = pc * TC_PIECE + h * TC_RBLK + l128;   // span-relative column
                        const bool active = (h == 0 || half1) && c < ccnt;
                        const float qx = S.cx[min(c, TC_CS - 1)], qy = S.cy[min(c, TC_CS - 1)], qz = S.cz[min(c, TC_CS - 1)];
                        const float rmax = __fsqrt_rn(__uint_as_float(S.rmax_bits));
                        const unsigned long long w = tc_resolve(active, qx, qy, qz, m, rmax, S.rx, S.ry, S.rz, npair * 2 * TC_RBLK, row0,
                                                                S.degenerate != 0, lane, p.stats);
                        if (active && w != ~0ull) atomicMin(p.pcol + (size_t)b * p.nc + c0 + c, w);
                    }
                    asm volatile("bar.sync 1, 256;" ::: "memory");   // cst is free for the next piece
                    if (lane == 0) tc_mbar_arrive(cempty + slot);
                }
                // ---- row points of the item against this span ----
                {
                    const float cmax = __fsqrt_rn(__uint_as_float(S.cmax_bits));
                    const int cnt32 = (ccnt + 31) & ~31;
                    for (int r = g; r < nrb; r += 2) {   // group g resolves the row blocks of its parity
                        Top3 m;
                        const int s0 = (r * 2 + 0) * TC_RBLK + l128, s1 = (r * 2 + 1) * TC_RBLK + l128;
                        m.b1 = S.rst[0][s0], m.b2 = S.rst[1][s0], m.b3 = S.rst[2][s0];
                        m.insert(S.rst[0][s1]), m.insert(S.rst[1][s1]), m.insert(S.rst[2][s1]);
                        const int rr = r * TC_RBLK + l128;
                        const bool active = rr < nrows;
                        const unsigned long long w = tc_resolve(active, S.rx[rr], S.ry[rr], S.rz[rr], m, cmax, S.cx, S.cy, S.cz, cnt32, c0,
                                                                S.degenerate != 0, lane, p.stats);
                        if (active && w != ~0ull) atomicMin(p.prow + (size_t)b * p.nr + row0 + rr, w);
                    }
                }
                __syncwarp();
                if (lane == 0) tc_mbar_arrive(span_done);
            }
            ++n_span;
        }
        if (p.stats != nullptr && tid == 0) {
            atomicAdd(p.stats + 3, 1u);
            if (S.degenerate) atomicAdd(p.stats + 2, 1u);
        }
        u += nrb;
    }
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    if (warp == 8) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, 512;" ::"r"(tmem) : "memory");
}


// Probe (tests only): e(x, y) of ONE 128 x 256 tile exactly as nn_tc_kernel's tensor pipe produces it, dumped to global
// memory -- tests/test_chamfer_tc.py compares it with float64 to pin the error assumption of the header.
__global__ void __launch_bounds__(128, 1) nn_tc_probe_kernel(const float *__restrict__ rows128, const float *__restrict__ cols256,
                                                             float *__restrict__ e_out /* [128][256] */) {
    __shared__ __align__(128) float4 opR[TC_KCH][TC_RBLK];
    __shared__ __align__(128) float4 opC[TC_KCH][TC_PIECE];
    __shared__ unsigned long long bar;
    __shared__ unsigned tmem_base;
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    if (warp == 0) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], 256;" ::"r"(tc_smem_u32(&tmem_base)) : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
    }
    if (tid == 0) {
        tc_mbar_init(&bar, 1);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    for (int r = tid; r < TC_RBLK + TC_PIECE; r += 128) {
        const bool isr = r < TC_RBLK;
        const float *pt = isr ? rows128 + r * 3 : cols256 + (r - TC_RBLK) * 3;
        const float x = pt[0], y = pt[1], z = pt[2];
        const float n = __fmaf_rn(z, z, __fmaf_rn(y, y, __fmul_rn(x, x)));
        float xh, xl, yh, yl, zh, zl, n1, n2, n3;
        tc_split3(x, xh, xl), tc_split3(y, yh, yl), tc_split3(z, zh, zl), tc_norm3(n, n1, n2, n3);
        if (isr) {
            opR[0][r] = make_float4(xh, yh, zh, xh), opR[1][r] = make_float4(yh, zh, xl, yl);
            opR[2][r] = make_float4(zl, n1, n2, n3), opR[3][r] = make_float4(1.f, 1.f, 1.f, 0.f);
        } else {
            const int k = r - TC_RBLK;
            opC[0][k] = make_float4(-2.f * xh, -2.f * yh, -2.f * zh, -2.f * xl), opC[1][k] = make_float4(-2.f * yl, -2.f * zl, -2.f * xh, -2.f * yh);
            opC[2][k] = make_float4(-2.f * zh, 1.f, 1.f, 1.f), opC[3][k] = make_float4(n1, n2, n3, 0.f);
        }
    }
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    const unsigned tmem = tmem_base;
    if (tid == 0) {
        const unsigned idesc = (1u << 4) | (2u << 7) | (2u << 10) | ((256u >> 3) << 17) | ((128u >> 4) << 24);
        const unsigned a = tc_smem_u32(&opR[0][0]), b = tc_smem_u32(&opC[0][0]);
        for (int ks = 0; ks < 2; ++ks)
            tc_mma(tmem, tc_desc(a + ks * 2 * TC_RBLK * 16, TC_RBLK * 16), tc_desc(b + ks * 2 * TC_PIECE * 16, TC_PIECE * 16), idesc, ks > 0);
        tc_commit(&bar);
    }
    tc_mbar_wait(&bar, 0);
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
#pragma unroll 1
    for (int j = 0; j < 8; ++j) {
        float v[32];
        tc_ld32(tmem + ((unsigned)(warp * 32) << 16) + j * 32, v);
        tc_wait_ld();
#pragma unroll
        for (int i = 0; i < 32; ++i) e_out[(size_t)(warp * 32 + lane) * 256 + j * 32 + i] = v[i];
    }
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    if (warp == 0) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, 256;" ::"r"(tmem) : "memory");
}

}  // namespace genpc
