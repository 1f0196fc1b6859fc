#include <cstdio>
#include <cstdlib>
#include <algorithm>
__host__ __device__ inline int imin(int a, int b) { return a < b ? a : b; }
__host__ __device__ inline int imax(int a, int b) { return a > b ? a : b; }
__device__ int fast_best_dev(const int *d)
{
    int m2[16], M2[16], m4[16], M4[16], m8[16], M8[16];
#pragma unroll
    for (int i = 0; i < 16; ++i) { m2[i] = min(d[i], d[(i + 1) & 15]); M2[i] = max(d[i], d[(i + 1) & 15]); }
#pragma unroll
    for (int i = 0; i < 16; ++i) { m4[i] = min(m2[i], m2[(i + 2) & 15]); M4[i] = max(M2[i], M2[(i + 2) & 15]); }
#pragma unroll
    for (int i = 0; i < 16; ++i) { m8[i] = min(m4[i], m4[(i + 4) & 15]); M8[i] = max(M4[i], M4[(i + 4) & 15]); }
    int bright = -256, dark = 256;
#pragma unroll
    for (int i = 0; i < 16; ++i) { bright = max(bright, min(m8[i], d[(i + 8) & 15])); dark = min(dark, max(M8[i], d[(i + 8) & 15])); }
    return max(bright, -dark);
}
int fast_best_host(const int *d)
{
    int best = -256;
    for (int k = 0; k < 16; ++k) {
        int lo = d[k], hi = d[k];
        for (int j = 1; j < 9; ++j) { lo = imin(lo, d[(k + j) & 15]); hi = imax(hi, d[(k + j) & 15]); }
        best = imax(best, imax(lo, -hi));
    }
    return best;
}
__global__ void k(const int *in, int *out, int n)
{
    int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    int d[16];
    for (int j = 0; j < 16; ++j) d[j] = in[i * 16 + j];
    out[i] = fast_best_dev(d);
}
int main(){
    const int n = 200000;
    int *h = (int*)malloc(n*64), *r = (int*)malloc(n*4);
    srand(1);
    for (int i = 0; i < n*16; ++i) { int m = rand() % 3; h[i] = m == 0 ? (rand() % 511) - 255 : (m == 1 ? 40 + rand() % 60 : -(40 + rand() % 60)); }
    int d0[16]={24,21,20,17,3,98,98,96,83,-15,98,98,21,41,18,23}; for (int j=0;j<16;++j) h[j]=d0[j];
    int *di,*dout; cudaMalloc(&di,n*64); cudaMalloc(&dout,n*4); cudaMemcpy(di,h,n*64,cudaMemcpyHostToDevice);
    k<<<(n+255)/256,256>>>(di,dout,n); cudaMemcpy(r,dout,n*4,cudaMemcpyDeviceToHost);
    int bad = 0; for (int i = 0; i < n; ++i) if (r[i] != fast_best_host(h + 16*i)) ++bad;
    printf("first %d (expect 18), mismatches %d of %d, %s\n", r[0], bad, n, cudaGetErrorString(cudaGetLastError()));
}
