// issue_mix.cu -- does a packed FP32 instruction (FFMA2: 2 FMA-pipe cycles per warp) leave the issue port free for ALU-pipe
// work in its second cycle?  The symmetric Chamfer scan mixes 384 packed FMA-pipe instructions with ~330 ALU / LSU
// instructions per 32-column block; whether it is bound by the FMA pipe (768 cycles) or by dispatch (768 + 330) decides
// what is left to gain.  Every mode runs independent dependency chains (8 packed chains, 8 min accumulators, 8 integer
// chains per thread), 256 threads, `ctas` CTAs per SM; reports SM cycles per unit per scheduler (SMSP).
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o tools/issue_mix tools/issue_mix.cu
#include <cuda_runtime.h>
#include <cstdio>
#include <cstdlib>

enum Mode { F2_ONLY = 0, M3_ONLY, M2_ONLY, LOP_ONLY, F2_M3_1_1, F2_M3_3_1, F2_M3_LOP_6_2_3, F1_M3_2_1, F1_ONLY, F2_LOP_1_1, NMODES };
static const char *names[NMODES] = {"ffma2_only", "fmnmx3_only", "fmnmx_only", "lop3_only", "ffma2_fmnmx3_1to1", "ffma2_fmnmx3_3to1",
                                    "ffma2_fmnmx3_lop3_6to2to3", "ffma_fmnmx3_2to1", "ffma_only", "ffma2_lop3_1to1"};
// instructions of each kind per unit
static const int nf2[NMODES] = {8, 0, 0, 0, 8, 6, 6, 0, 0, 8};
static const int nf1[NMODES] = {0, 0, 0, 0, 0, 0, 0, 8, 8, 0};
static const int nm3[NMODES] = {0, 8, 0, 0, 8, 2, 2, 4, 0, 0};
static const int nm2[NMODES] = {0, 0, 8, 0, 0, 0, 0, 0, 0, 0};
static const int nlop[NMODES] = {0, 0, 0, 8, 0, 0, 3, 0, 0, 8};

#define F2(i) asm volatile("fma.rn.f32x2 %0, %0, %1, %2;" : "+l"(x[i]) : "l"(y), "l"(z))
#define F1(i) asm volatile("fma.rn.f32 %0, %0, %1, %2;" : "+f"(s[i]) : "f"(sy), "f"(sz))
#define M3(i) asm volatile("min.f32 %0, %0, %1, %2;" : "+f"(m[i]) : "f"(sy), "f"(sz))
#define M2(i) asm volatile("min.f32 %0, %0, %1;" : "+f"(m[i]) : "f"(sy))
#define LOP(i) asm volatile("lop3.b32 %0, %0, %1, %2, 0x96;" : "+r"(q[i]) : "r"(qa), "r"(qb))

template <int MODE>
__global__ void __launch_bounds__(256) mix_kernel(float *out, int iters, float seed, long long *cycles) {
    unsigned long long x[8];
    float s[8], m[8];
    unsigned q[8];
    const unsigned long long y = ((unsigned long long)__float_as_uint(seed) << 32) | __float_as_uint(seed * 0.5f);
    const unsigned long long z = ((unsigned long long)__float_as_uint(seed * 0.25f) << 32) | __float_as_uint(seed * 0.125f);
    const float sy = seed * 0.999f, sz = seed * 0.001f;
    const unsigned qa = threadIdx.x * 2654435761u, qb = (unsigned)(seed * 1000.f);
#pragma unroll
    for (int i = 0; i < 8; ++i) x[i] = y + i + threadIdx.x, s[i] = seed + i, m[i] = 3e38f - i - threadIdx.x, q[i] = i + threadIdx.x;
    const long long t0 = clock64();
    for (int it = 0; it < iters; ++it) {
#pragma unroll
        for (int k = 0; k < 16; ++k) {
            if (MODE == F2_ONLY) { F2(0); F2(1); F2(2); F2(3); F2(4); F2(5); F2(6); F2(7); }
            if (MODE == F1_ONLY) { F1(0); F1(1); F1(2); F1(3); F1(4); F1(5); F1(6); F1(7); }
            if (MODE == M3_ONLY) { M3(0); M3(1); M3(2); M3(3); M3(4); M3(5); M3(6); M3(7); }
            if (MODE == M2_ONLY) { M2(0); M2(1); M2(2); M2(3); M2(4); M2(5); M2(6); M2(7); }
            if (MODE == LOP_ONLY) { LOP(0); LOP(1); LOP(2); LOP(3); LOP(4); LOP(5); LOP(6); LOP(7); }
            if (MODE == F2_M3_1_1) { F2(0); M3(0); F2(1); M3(1); F2(2); M3(2); F2(3); M3(3); F2(4); M3(4); F2(5); M3(5); F2(6); M3(6); F2(7); M3(7); }
            if (MODE == F2_LOP_1_1) { F2(0); LOP(0); F2(1); LOP(1); F2(2); LOP(2); F2(3); LOP(3); F2(4); LOP(4); F2(5); LOP(5); F2(6); LOP(6); F2(7); LOP(7); }
            if (MODE == F2_M3_3_1) { F2(0); F2(1); F2(2); M3(0); F2(3); F2(4); F2(5); M3(1); }
            if (MODE == F2_M3_LOP_6_2_3) { F2(0); F2(1); LOP(0); F2(2); M3(0); F2(3); LOP(1); F2(4); F2(5); M3(1); LOP(2); }
            if (MODE == F1_M3_2_1) { F1(0); F1(1); M3(0); F1(2); F1(3); M3(1); F1(4); F1(5); M3(2); F1(6); F1(7); M3(3); }
        }
    }
    const long long t1 = clock64();
    float r = 0.f;
#pragma unroll
    for (int i = 0; i < 8; ++i) r += __uint_as_float((unsigned)(x[i] >> 32)) + __uint_as_float((unsigned)x[i]) + s[i] + m[i] + (float)q[i];
    if (r == 123.456f) out[0] = r;
    if (threadIdx.x == 0) cycles[blockIdx.x] = t1 - t0;
}

template <int MODE>
static void run(int sms, int ctas, int iters, float *dout, long long *dcyc) {
    const int grid = sms * ctas;
    mix_kernel<MODE><<<grid, 256>>>(dout, iters / 8 + 1, 1.0f, dcyc);
    cudaDeviceSynchronize();
    long long best = 1LL << 62;
    long long *h = (long long *)malloc(sizeof(long long) * grid);
    for (int rep = 0; rep < 5; ++rep) {
        mix_kernel<MODE><<<grid, 256>>>(dout, iters, 1.0f, dcyc);
        cudaDeviceSynchronize();
        cudaMemcpy(h, dcyc, sizeof(long long) * grid, cudaMemcpyDeviceToHost);
        long long cyc = 0;
        for (int i = 0; i < grid; ++i) cyc = h[i] > cyc ? h[i] : cyc;
        if (cyc < best) best = cyc;
    }
    free(h);
    const double units_per_warp = (double)iters * 16;
    const double warps_per_smsp = 8.0 * ctas / 4.0;
    const double cyc_per_unit = (double)best / (units_per_warp * warps_per_smsp);  // scheduler cycles per unit of one warp
    const int ninstr = nf2[MODE] + nf1[MODE] + nm3[MODE] + nm2[MODE] + nlop[MODE];
    printf("  \"%s\": {\"ffma2\": %d, \"ffma\": %d, \"fmnmx3\": %d, \"fmnmx\": %d, \"lop3\": %d, \"smsp_cycles_per_unit\": %.3f, "
           "\"issue_per_clk_per_smsp\": %.3f, \"fma_pipe_cycles_per_unit_if_2_per_packed\": %d},\n",
           names[MODE], nf2[MODE], nf1[MODE], nm3[MODE], nm2[MODE], nlop[MODE], cyc_per_unit, ninstr / cyc_per_unit,
           2 * nf2[MODE] + nf1[MODE]);
}

int main(int argc, char **argv) {
    cudaDeviceProp prop;
    cudaGetDeviceProperties(&prop, 0);
    const int sms = prop.multiProcessorCount;
    float *dout;
    long long *dcyc;
    cudaMalloc(&dout, 4);
    cudaMalloc(&dcyc, sizeof(long long) * sms * 8);
    const int iters = argc > 1 ? atoi(argv[1]) : 2000;
    const int ctas = argc > 2 ? atoi(argv[2]) : 2;
    printf("{\n  \"gpu\": \"%s\", \"sms\": %d, \"ctas_per_sm\": %d, \"threads\": 256,\n", prop.name, sms, ctas);
    run<F2_ONLY>(sms, ctas, iters, dout, dcyc);
    run<F1_ONLY>(sms, ctas, iters, dout, dcyc);
    run<M3_ONLY>(sms, ctas, iters, dout, dcyc);
    run<M2_ONLY>(sms, ctas, iters, dout, dcyc);
    run<LOP_ONLY>(sms, ctas, iters, dout, dcyc);
    run<F2_M3_1_1>(sms, ctas, iters, dout, dcyc);
    run<F2_LOP_1_1>(sms, ctas, iters, dout, dcyc);
    run<F2_M3_3_1>(sms, ctas, iters, dout, dcyc);
    run<F2_M3_LOP_6_2_3>(sms, ctas, iters, dout, dcyc);
    run<F1_M3_2_1>(sms, ctas, iters, dout, dcyc);
    printf("  \"note\": \"smsp_cycles_per_unit = max CTA clock64 span / (units per warp * warps per scheduler)\"\n}\n");
    return cudaGetLastError() != cudaSuccess;
}
