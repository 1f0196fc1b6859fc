// microbench.cu -- measures the FP32 CUDA-core issue ceilings the scoring kernel lives under:
// FFMA (3-register), packed FFMA2 (fma.rn.f32x2, sm_100+), and the scoring loop's instruction mix
// (FMA pipe + ALU pipe + MUFU).  Prints one JSON line.  Build: nvcc -arch=sm_100a -O3 -o microbench microbench.cu
#include <cstdio>
#include <cuda_runtime.h>

#define ITERS 4096

__global__ void k_ffma(float *out, float a, float b)
{
    float r[8];
#pragma unroll
    for (int i = 0; i < 8; ++i) r[i] = threadIdx.x * 0.001f + i;
    for (int it = 0; it < ITERS; ++it) {
#pragma unroll
        for (int i = 0; i < 8; ++i) r[i] = fmaf(r[i], a, b);
    }
    float s = 0;
#pragma unroll
    for (int i = 0; i < 8; ++i) s += r[i];
    if (s == 123.456f) out[0] = s;
}

__global__ void k_ffma2(float *out, float a, float b)
{
    unsigned long long r[8], av, bv;
    asm("mov.b64 %0, {%1, %1};" : "=l"(av) : "f"(a));
    asm("mov.b64 %0, {%1, %1};" : "=l"(bv) : "f"(b));
#pragma unroll
    for (int i = 0; i < 8; ++i) { float v = threadIdx.x * 0.001f + i; asm("mov.b64 %0, {%1, %1};" : "=l"(r[i]) : "f"(v)); }
    for (int it = 0; it < ITERS; ++it) {
#pragma unroll
        for (int i = 0; i < 8; ++i) asm volatile("fma.rn.f32x2 %0, %0, %1, %2;" : "+l"(r[i]) : "l"(av), "l"(bv));
    }
    float s = 0;
#pragma unroll
    for (int i = 0; i < 8; ++i) { float lo, hi; asm("mov.b64 {%0, %1}, %2;" : "=f"(lo), "=f"(hi) : "l"(r[i])); s += lo + hi; }
    if (s == 123.456f) out[0] = s;
}

// scoring-like mix per "pair": 10 FMA-pipe ops, 1 MUFU.RCP, 1 integer add of the sign bit, 1 FMNMX
__global__ void k_mix(float *out, float a, float b, int n_alu, int use_mufu)
{
    float x[4], m[4];
    unsigned c[4];
#pragma unroll
    for (int i = 0; i < 4; ++i) { x[i] = threadIdx.x * 0.001f + i + 1.f; m[i] = 1e30f; c[i] = 0; }
    for (int it = 0; it < ITERS; ++it) {
#pragma unroll
        for (int i = 0; i < 4; ++i) {
            float r0 = fmaf(x[i], a, b), r1 = fmaf(x[i], b, a), w = fmaf(x[i], a, 1.0f);
            r0 = fmaf(r0, a, b); r1 = fmaf(r1, b, a); w = fmaf(w, b, 1.0f);
            float iw;
            if (use_mufu) asm("rcp.approx.ftz.f32 %0, %1;" : "=f"(iw) : "f"(w)); else iw = w * a;
            float dx = fmaf(r0, iw, -a), dy = fmaf(r1, iw, -b);
            float q = fmaf(dy, dy, fmaf(dx, dx, -9.0f));
            if (n_alu >= 1) c[i] += __float_as_uint(q) >> 31;
            if (n_alu >= 2) m[i] = fminf(m[i], fabsf(q));
            x[i] = q * 1e-3f + x[i];
        }
    }
    float s = 0;
#pragma unroll
    for (int i = 0; i < 4; ++i) s += x[i] + m[i] + c[i];
    if (s == 123.456f) out[0] = s;
}

template <typename F> float time_ms(F f, int reps = 5)
{
    cudaEvent_t a, b;
    cudaEventCreate(&a); cudaEventCreate(&b);
    f(); cudaDeviceSynchronize();
    float best = 1e30f;
    for (int i = 0; i < reps; ++i) {
        cudaEventRecord(a); f(); cudaEventRecord(b); cudaEventSynchronize(b);
        float ms; cudaEventElapsedTime(&ms, a, b);
        if (ms < best) best = ms;
    }
    return best;
}

int main()
{
    cudaDeviceProp p; cudaGetDeviceProperties(&p, 0);
    int sms = p.multiProcessorCount, khz = 0;
    cudaDeviceGetAttribute(&khz, cudaDevAttrClockRate, 0);
    float *out; cudaMalloc(&out, 4);
    int blocks = sms * 8, threads = 256;
    double lanes = (double)blocks * threads;
    float t1 = time_ms([&] { k_ffma<<<blocks, threads>>>(out, 1.0001f, 0.5f); });
    float t2 = time_ms([&] { k_ffma2<<<blocks, threads>>>(out, 1.0001f, 0.5f); });
    double ffma = lanes * ITERS * 8 * 2 / (t1 * 1e-3) / 1e12;
    double ffma2 = lanes * ITERS * 8 * 4 / (t2 * 1e-3) / 1e12;
    printf("{\"gpu\": \"%s\", \"sms\": %d, \"clock_khz\": %d, \"ffma_tflops\": %.2f, \"ffma2_tflops\": %.2f", p.name, sms, khz, ffma, ffma2);
    const char *names[4] = {"mix_fma_only", "mix_fma_mufu", "mix_fma_mufu_1alu", "mix_fma_mufu_2alu"};
    int cfg[4][2] = {{0, 0}, {0, 1}, {1, 1}, {2, 1}};
    for (int i = 0; i < 4; ++i) {
        float t = time_ms([&] { k_mix<<<blocks, threads>>>(out, 1.0001f, 0.5f, cfg[i][0], cfg[i][1]); });
        double pairs = lanes * ITERS * 4 / (t * 1e-3);
        printf(", \"%s_gpairs_per_s\": %.1f", names[i], pairs / 1e9);
    }
    printf("}\n");
    return 0;
}
