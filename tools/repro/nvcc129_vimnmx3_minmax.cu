#include <cstdio>
__global__ void k(const int *in, int *out)
{
    int d[16];
    for (int i = 0; i < 16; ++i) d[i] = in[i];
    int best = -256;
#pragma unroll
    for (int k = 0; k < 16; ++k) {
        int lo = d[k], hi = d[k];
#pragma unroll
        for (int j = 1; j < 9; ++j) {
            lo = min(lo, d[(k + j) & 15]);
            hi = max(hi, d[(k + j) & 15]);
        }
        best = max(best, max(lo, -hi));
    }
    out[0] = best;
}
int main(){ int d[16]={24,21,20,17,3,98,98,96,83,-15,98,98,21,41,18,23}; int *di,*dout; cudaMalloc(&di,64); cudaMalloc(&dout,4); cudaMemcpy(di,d,64,cudaMemcpyHostToDevice); k<<<1,1>>>(di,dout); int r; cudaMemcpy(&r,dout,4,cudaMemcpyDeviceToHost); printf("device %d (expect 18) %s\n", r, cudaGetErrorString(cudaGetLastError())); }
