// Minimal check of the 2-D TMA tensor copy used by warp_pipe_kernel: encode a uint8 map over a
// pitched byte array, copy one 256 x 48 box at an arbitrary (x, y), compare with the source.
#include <cuda.h>
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>
#include <vector>
constexpr int BW = 256, BH = 48;
__global__ void k(const __grid_constant__ CUtensorMap tmap, int x, int y, uint8_t *out)
{
    __shared__ __align__(128) uint8_t buf[BW * BH];
    __shared__ __align__(8) unsigned long long mbar;
    const uint32_t sb = (uint32_t)__cvta_generic_to_shared(buf), sm = (uint32_t)__cvta_generic_to_shared(&mbar);
    if (threadIdx.x == 0) {
        asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(sm) : "memory");
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
        asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
    }
    __syncthreads();
    if (threadIdx.x == 0) {
        asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(sm), "r"(BW * BH) : "memory");
        asm volatile("cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%2, %3}], [%4];"
                     ::"r"(sb), "l"(reinterpret_cast<uint64_t>(&tmap)), "r"(x), "r"(y), "r"(sm) : "memory");
    }
    unsigned spins = 0;
    for (;;) {
        uint32_t ok;
        asm volatile("{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\tselp.u32 %0, 1, 0, p;\n\t}" : "=r"(ok) : "r"(sm), "r"(0) : "memory");
        if (ok) break;
        if (++spins > (1u << 22)) { if (threadIdx.x == 0) printf("timeout\n"); return; }
    }
    for (int i = threadIdx.x; i < BW * BH; i += blockDim.x) out[i] = buf[i];
}
int main()
{
    const long long pitch = 5760, rows = 8 * 1080;
    std::vector<uint8_t> h(pitch * rows);
    for (size_t i = 0; i < h.size(); ++i) h[i] = (uint8_t)((i * 2654435761u) >> 13);
    uint8_t *d, *o;
    cudaMalloc(&d, h.size()); cudaMalloc(&o, BW * BH);
    cudaMemcpy(d, h.data(), h.size(), cudaMemcpyHostToDevice);
    void *fn = nullptr; cudaDriverEntryPointQueryResult q;
    cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &fn, cudaEnableDefault, &q);
    printf("entry %p q %d\n", fn, (int)q);
    typedef CUresult (*encode_fn)(CUtensorMap *, CUtensorMapDataType, cuuint32_t, void *, const cuuint64_t *, const cuuint64_t *, const cuuint32_t *, const cuuint32_t *, CUtensorMapInterleave, CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);
    CUtensorMap m;
    cuuint64_t gdim[2] = {(cuuint64_t)pitch, (cuuint64_t)rows}, gstr[1] = {(cuuint64_t)pitch};
    cuuint32_t box[2] = {BW, BH}, es[2] = {1, 1};
    CUresult r = ((encode_fn)fn)(&m, CU_TENSOR_MAP_DATA_TYPE_UINT8, 2, d, gdim, gstr, box, es, CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_L2_128B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    printf("encode %d\n", (int)r);
    for (int t = 0; t < 4; ++t) {
        const int xs[4] = {0, 48, 3 * 333, 3 * 1801}, ys[4] = {0, 17, 2000, 8 * 1080 - 20};
        k<<<1, 128>>>(m, xs[t], ys[t], o);
        cudaError_t e = cudaDeviceSynchronize();
        printf("x %d y %d: %s\n", xs[t], ys[t], cudaGetErrorString(e));
        if (e != cudaSuccess) return 1;
        std::vector<uint8_t> g(BW * BH);
        cudaMemcpy(g.data(), o, g.size(), cudaMemcpyDeviceToHost);
        long bad = 0;
        for (int rr = 0; rr < BH; ++rr) for (int c = 0; c < BW; ++c) {
            long long yy = ys[t] + rr, xx = xs[t] + c;
            uint8_t want = (yy < rows && xx < pitch) ? h[yy * pitch + xx] : 0;
            bad += g[rr * BW + c] != want;
        }
        printf("  mismatches %ld\n", bad);
    }
    return 0;
}
