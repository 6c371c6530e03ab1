// fp32_peak.cu -- FP32 CUDA-core issue-rate microbenchmark for B200 (sm_100a).
// MEASURED_PEAKS.json has no FP32 figure; the Chamfer NN scan is FP32-pipe bound, so its roofline
// denominator is measured here: scalar FFMA/FADD, packed FFMA2/FADD2/FMUL2 and FMNMX3 issue rates (instruction MIXES are
// measured by tools/issue_mix.cu).  Prints one JSON object.  Build:
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o tools/fp32_peak tools/fp32_peak.cu
#include <cuda_runtime.h>
#include <cstdio>
#include <cstdlib>
#include <cstring>

#define CHAINS 16
#define INNER 64

enum Mode { FFMA = 0, FFMA2, FADD, FADD2, FMUL2, FMNMX3, NMODES };
static const char *names[NMODES] = {"ffma", "ffma2", "fadd", "fadd2", "fmul2", "fmnmx3"};
// lane-flops and instructions per asm statement group (per thread per inner step per chain pair): the scalar modes and
// FMNMX3 issue TWO instructions per group (one per chain), the packed modes one.
static const double flops_per_step[NMODES] = {4, 4, 2, 2, 2, 0};
static const double instr_per_step[NMODES] = {2, 1, 2, 1, 1, 2};

template <int MODE>
__global__ void __launch_bounds__(256) peak_kernel(float *out, int iters, float seed, long long *cycles) {
    float a[CHAINS], b[CHAINS];
    float m = 3.0e38f;
    const float c0 = seed * 0.5f, c1 = seed * 0.25f, c2 = seed * 0.125f;
#pragma unroll
    for (int i = 0; i < CHAINS; ++i) {
        a[i] = seed + i + threadIdx.x;
        b[i] = seed - i;
    }
    long long t0 = clock64();
    for (int it = 0; it < iters; ++it) {
#pragma unroll
        for (int k = 0; k < INNER; ++k) {
#pragma unroll
            for (int i = 0; i < CHAINS; i += 2) {
                if (MODE == FFMA) {
                    asm volatile("fma.rn.f32 %0, %0, %1, %2;" : "+f"(a[i]) : "f"(b[i]), "f"(c0));
                    asm volatile("fma.rn.f32 %0, %0, %1, %2;" : "+f"(a[i + 1]) : "f"(b[i + 1]), "f"(c1));
                } else if (MODE == FADD) {
                    asm volatile("add.rn.f32 %0, %0, %1;" : "+f"(a[i]) : "f"(b[i]));
                    asm volatile("add.rn.f32 %0, %0, %1;" : "+f"(a[i + 1]) : "f"(b[i + 1]));
                } else if (MODE == FFMA2) {
                    asm volatile(
                        "{.reg .b64 x, y, z; mov.b64 x, {%0, %1}; mov.b64 y, {%2, %3}; mov.b64 z, {%4, %5};\n"
                        "fma.rn.f32x2 x, x, y, z; mov.b64 {%0, %1}, x;}"
                        : "+f"(a[i]), "+f"(a[i + 1])
                        : "f"(b[i]), "f"(b[i + 1]), "f"(c0), "f"(c1));
                } else if (MODE == FADD2) {
                    asm volatile(
                        "{.reg .b64 x, y; mov.b64 x, {%0, %1}; mov.b64 y, {%2, %3};\n"
                        "add.rn.f32x2 x, x, y; mov.b64 {%0, %1}, x;}"
                        : "+f"(a[i]), "+f"(a[i + 1])
                        : "f"(b[i]), "f"(b[i + 1]));
                } else if (MODE == FMUL2) {
                    asm volatile(
                        "{.reg .b64 x, y; mov.b64 x, {%0, %1}; mov.b64 y, {%2, %3};\n"
                        "mul.rn.f32x2 x, x, y; mov.b64 {%0, %1}, x;}"
                        : "+f"(a[i]), "+f"(a[i + 1])
                        : "f"(b[i]), "f"(b[i + 1]));
                } else if (MODE == FMNMX3) {
                    asm volatile("min.f32 %0, %0, %1, %2;" : "+f"(a[i]) : "f"(b[i]), "f"(b[i + 1]));
                    asm volatile("min.f32 %0, %0, %1, %2;" : "+f"(a[i + 1]) : "f"(b[i]), "f"(b[i + 1]));
                }
            }
        }
    }
    long long t1 = clock64();
    float r = m;
#pragma unroll
    for (int i = 0; i < CHAINS; ++i) r += a[i] + b[i];
    if (r == 123.456f) out[0] = r;
    if (threadIdx.x == 0) cycles[blockIdx.x] = t1 - t0;
}

template <int MODE>
static void run(int sms, int ctas_per_sm, int iters, float *dout, long long *dcyc, double *tflops, double *ipc) {
    int grid = sms * ctas_per_sm;
    cudaEvent_t e0, e1;
    cudaEventCreate(&e0);
    cudaEventCreate(&e1);
    peak_kernel<MODE><<<grid, 256>>>(dout, iters / 8 + 1, 1.0f, dcyc);
    cudaDeviceSynchronize();
    float best = 1e30f;
    long long cyc = 0;
    for (int rep = 0; rep < 5; ++rep) {
        cudaEventRecord(e0);
        peak_kernel<MODE><<<grid, 256>>>(dout, iters, 1.0f, dcyc);
        cudaEventRecord(e1);
        cudaEventSynchronize(e1);
        float ms;
        cudaEventElapsedTime(&ms, e0, e1);
        if (ms < best) {
            best = ms;
            long long *h = (long long *)malloc(sizeof(long long) * grid);
            cudaMemcpy(h, dcyc, sizeof(long long) * grid, cudaMemcpyDeviceToHost);
            cyc = 0;
            for (int i = 0; i < grid; ++i) cyc = h[i] > cyc ? h[i] : cyc;
            free(h);
        }
    }
    double steps = (double)iters * INNER * (CHAINS / 2) * 256.0 * grid;  // asm groups executed (thread level)
    *tflops = steps * flops_per_step[MODE] / (best * 1e-3) / 1e12;
    // warp-instructions per clock per SM (each group = instr_per_step instrs, 8 warps/CTA)
    double winstr_per_sm = (double)iters * INNER * (CHAINS / 2) * instr_per_step[MODE] * 8.0 * ctas_per_sm;
    *ipc = winstr_per_sm / (double)cyc;
    printf("  \"%s\": {\"ms\": %.4f, \"tflops\": %.2f, \"warp_instr_per_clk_per_sm\": %.3f, \"sm_cycles\": %lld, "
           "\"implied_mhz\": %.0f},\n",
           names[MODE], best, *tflops, *ipc, cyc, cyc / (best * 1e-3) / 1e6);
}

int main(int argc, char **argv) {
    int dev = 0;
    cudaSetDevice(dev);
    cudaDeviceProp prop;
    cudaGetDeviceProperties(&prop, dev);
    int sms = prop.multiProcessorCount;
    float *dout;
    long long *dcyc;
    cudaMalloc(&dout, 4);
    cudaMalloc(&dcyc, sizeof(long long) * sms * 8);
    int iters = argc > 1 ? atoi(argv[1]) : 400;
    int cps = argc > 2 ? atoi(argv[2]) : 4;
    double tf, ipc;
    printf("{\n  \"gpu\": \"%s\", \"sms\": %d, \"clock_khz\": %d, \"ctas_per_sm\": %d,\n", prop.name, sms,
           prop.clockRate, cps);
    run<FFMA>(sms, cps, iters, dout, dcyc, &tf, &ipc);
    run<FFMA2>(sms, cps, iters, dout, dcyc, &tf, &ipc);
    run<FADD>(sms, cps, iters, dout, dcyc, &tf, &ipc);
    run<FADD2>(sms, cps, iters, dout, dcyc, &tf, &ipc);
    run<FMUL2>(sms, cps, iters, dout, dcyc, &tf, &ipc);
    run<FMNMX3>(sms, cps, iters, dout, dcyc, &tf, &ipc);
    printf("  \"note\": \"tflops counts FMA=2; instruction mixes: tools/issue_mix.cu\"\n}\n");
    cudaError_t e = cudaGetLastError();
    if (e != cudaSuccess) {
        fprintf(stderr, "CUDA error: %s\n", cudaGetErrorString(e));
        return 1;
    }
    return 0;
}
