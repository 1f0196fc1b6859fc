// microbench2.cu -- what limits the packed-FFMA2 scoring loop (K2) on B200: register-file operand bandwidth and the ALU ops
// mixed into the FMA stream.  Prints one JSON line.  Build: nvcc -arch=sm_100a -O3 -o microbench2 microbench2.cu
#include <cstdio>
#include <cuda_runtime.h>

#define ITERS 2048
typedef unsigned long long u64;

__device__ __forceinline__ u64 pack(float a, float b) { u64 r; asm("mov.b64 %0, {%1, %2};" : "=l"(r) : "f"(a), "f"(b)); return r; }
__device__ __forceinline__ float lo(u64 v) { float a, b; asm("mov.b64 {%0, %1}, %2;" : "=f"(a), "=f"(b) : "l"(v)); (void)b; return a; }
__device__ __forceinline__ float hi(u64 v) { float a, b; asm("mov.b64 {%0, %1}, %2;" : "=f"(a), "=f"(b) : "l"(v)); (void)a; return b; }
#define FMA2(d, a, b, c) asm volatile("fma.rn.f32x2 %0, %1, %2, %3;" : "=l"(d) : "l"(a), "l"(b), "l"(c))

// MODE 0: r = r * A + B            (one fresh 64-bit operand per instruction: the other two are reused)
// MODE 1: r[i] = x[i] * y[i] + r[i] (three distinct 64-bit operands, no reuse possible)
// MODE 2: r[i] = x[i] * Y + r[i]    (one operand shared by consecutive instructions)
template <int MODE>
__global__ void k_f2(float *out, float a, float b)
{
    u64 r[8], x[8], y[8];
#pragma unroll
    for (int i = 0; i < 8; ++i) {
        r[i] = pack(threadIdx.x * 0.001f + i, 1.f + i);
        x[i] = pack(a + 1e-6f * i, a - 1e-6f * i);
        y[i] = pack(b + 1e-6f * i, b - 1e-6f * i);
    }
    const u64 A = pack(a, a), B = pack(b, b);
    for (int it = 0; it < ITERS; ++it) {
#pragma unroll
        for (int i = 0; i < 8; ++i) {
            if (MODE == 0) FMA2(r[i], r[i], A, B);
            if (MODE == 1) FMA2(r[i], x[i], y[i], r[i]);
            if (MODE == 2) FMA2(r[i], x[i], B, r[i]);
        }
    }
    float s = 0;
#pragma unroll
    for (int i = 0; i < 8; ++i) s += lo(r[i]) + hi(r[i]);
    if (s == 123.456f) out[0] = s;
}

// The scoring loop's shape: per correspondence and hypothesis pair 10 FFMA2 + 1 FMUL2, then per hypothesis one sign-bit
// accumulation and one running min |q|.   ALU = 0: one LOP3 per two values (keeps them alive); 1: LEA.HI count + min per
// correspondence; 2: LEA.HI count + one min over two correspondences (FMNMX3); 3: PRMT sign masks + IADD3 (two hypotheses per
// counter word, two correspondences per add) + FMNMX3.
template <int ALU>
__global__ void k_score(float *out, const float *__restrict__ pts, float a, float b)
{
    __shared__ float4 tile[2 * 256];
    for (int i = threadIdx.x; i < 512; i += blockDim.x) tile[i] = make_float4(pts[i & 63], pts[i & 63], pts[(i + 7) & 63], pts[(i + 7) & 63]);
    __syncthreads();
    float2 h[2][9];
#pragma unroll
    for (int g = 0; g < 2; ++g)
#pragma unroll
        for (int j = 0; j < 9; ++j) h[g][j] = make_float2(__ldg(pts + ((threadIdx.x + 18 * g + 2 * j) & 63)) * a, __ldg(pts + ((threadIdx.x + 18 * g + 2 * j + 1) & 63)) * b);
    unsigned c[4] = {0, 0, 0, 0};
    unsigned cp[2] = {0, 0};
    float m[4] = {1e30f, 1e30f, 1e30f, 1e30f};
    for (int it = 0; it < ITERS / 8; ++it) {
#pragma unroll
        for (int u = 0; u < 16; u += 2) {
            float2 q[2][2];
#pragma unroll
            for (int v = 0; v < 2; ++v) {
                const float4 pa = tile[2 * ((it * 16 + u + v) & 255)], pb = tile[2 * ((it * 16 + u + v) & 255) + 1];
                const float2 xx = make_float2(pa.x, pa.y), yy = make_float2(pa.z, pa.w), nX = make_float2(pb.x, pb.y), nY = make_float2(pb.z, pb.w);
#pragma unroll
                for (int g = 0; g < 2; ++g) {
                    const float2 w = __ffma2_rn(h[g][6], xx, __ffma2_rn(h[g][7], yy, h[g][8]));
                    const float2 aa = __ffma2_rn(h[g][0], xx, __ffma2_rn(h[g][1], yy, __ffma2_rn(nX, w, h[g][2])));
                    const float2 bb = __ffma2_rn(h[g][3], xx, __ffma2_rn(h[g][4], yy, __ffma2_rn(nY, w, h[g][5])));
                    const float2 t = __fmul2_rn(w, w);
                    q[v][g] = __ffma2_rn(aa, aa, __ffma2_rn(bb, bb, make_float2(-t.x, -t.y)));
                }
            }
#pragma unroll
            for (int g = 0; g < 2; ++g) {
                const float q0x = q[0][g].x, q0y = q[0][g].y, q1x = q[1][g].x, q1y = q[1][g].y;
                if (ALU == 0) {
                    c[2 * g] ^= __float_as_uint(q0x) ^ __float_as_uint(q1x);
                    c[2 * g + 1] ^= __float_as_uint(q0y) ^ __float_as_uint(q1y);
                }
                if (ALU == 1) {
                    c[2 * g] += __float_as_uint(q0x) >> 31; c[2 * g + 1] += __float_as_uint(q0y) >> 31;
                    c[2 * g] += __float_as_uint(q1x) >> 31; c[2 * g + 1] += __float_as_uint(q1y) >> 31;
                    asm volatile("min.f32 %0, %0, %1;" : "+f"(m[2 * g]) : "f"(fabsf(q0x)));
                    asm volatile("min.f32 %0, %0, %1;" : "+f"(m[2 * g + 1]) : "f"(fabsf(q0y)));
                    asm volatile("min.f32 %0, %0, %1;" : "+f"(m[2 * g]) : "f"(fabsf(q1x)));
                    asm volatile("min.f32 %0, %0, %1;" : "+f"(m[2 * g + 1]) : "f"(fabsf(q1y)));
                }
                if (ALU >= 2) {
                    m[2 * g] = fminf(fminf(m[2 * g], fabsf(q0x)), fabsf(q1x));
                    m[2 * g + 1] = fminf(fminf(m[2 * g + 1], fabsf(q0y)), fabsf(q1y));
                }
                if (ALU == 2) {
                    c[2 * g] += __float_as_uint(q0x) >> 31; c[2 * g + 1] += __float_as_uint(q0y) >> 31;
                    c[2 * g] += __float_as_uint(q1x) >> 31; c[2 * g + 1] += __float_as_uint(q1y) >> 31;
                }
                if (ALU == 3) {
                    const unsigned s0 = __byte_perm(__float_as_uint(q0x), __float_as_uint(q0y), 0xFFBB);
                    const unsigned s1 = __byte_perm(__float_as_uint(q1x), __float_as_uint(q1y), 0xFFBB);
                    cp[g] += s0 + s1;
                }
            }
        }
    }
    float s = 0;
#pragma unroll
    for (int i = 0; i < 4; ++i) s += m[i] + c[i];
    s += cp[0] + cp[1];
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
    const int sms = p.multiProcessorCount;
    float *out, *pts; cudaMalloc(&out, 4); cudaMalloc(&pts, 256);
    float hp[64]; for (int i = 0; i < 64; ++i) hp[i] = 0.5f + 0.01f * i;
    cudaMemcpy(pts, hp, 256, cudaMemcpyHostToDevice);
    const int blocks = sms * 2, threads = 256;
    const double lanes = (double)blocks * threads;
    printf("{\"gpu\": \"%s\"", p.name);
    const char *n2[3] = {"ffma2_reuse2", "ffma2_3distinct", "ffma2_share1"};
    float t;
    t = time_ms([&] { k_f2<0><<<blocks, threads>>>(out, 1.0001f, 0.5f); }); printf(", \"%s_tflops\": %.2f", n2[0], lanes * ITERS * 8 * 4 / (t * 1e-3) / 1e12);
    t = time_ms([&] { k_f2<1><<<blocks, threads>>>(out, 1.0001f, 0.5f); }); printf(", \"%s_tflops\": %.2f", n2[1], lanes * ITERS * 8 * 4 / (t * 1e-3) / 1e12);
    t = time_ms([&] { k_f2<2><<<blocks, threads>>>(out, 1.0001f, 0.5f); }); printf(", \"%s_tflops\": %.2f", n2[2], lanes * ITERS * 8 * 4 / (t * 1e-3) / 1e12);
    const double pairs = lanes * (ITERS / 8) * 16 * 4;
    t = time_ms([&] { k_score<0><<<blocks, threads>>>(out, pts, 1.0001f, 0.5f); }); printf(", \"score_fma_only_gpairs\": %.1f", pairs / (t * 1e-3) / 1e9);
    t = time_ms([&] { k_score<1><<<blocks, threads>>>(out, pts, 1.0001f, 0.5f); }); printf(", \"score_lea_fmnmx_gpairs\": %.1f", pairs / (t * 1e-3) / 1e9);
    t = time_ms([&] { k_score<2><<<blocks, threads>>>(out, pts, 1.0001f, 0.5f); }); printf(", \"score_lea_min3_gpairs\": %.1f", pairs / (t * 1e-3) / 1e9);
    t = time_ms([&] { k_score<3><<<blocks, threads>>>(out, pts, 1.0001f, 0.5f); }); printf(", \"score_prmt_iadd3_min3_gpairs\": %.1f", pairs / (t * 1e-3) / 1e9);
    printf("}\n");
    return 0;
}
