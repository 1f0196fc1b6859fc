// match.cuh -- K6: brute-force Hamming matcher with cross-check.
//
// Replaces cv2.BFMatcher(cv2.NORM_HAMMING, crossCheck=True).match(featuresA, featuresB) in the reference's
// stitching() (ransac.py:258-259), the step between ORB and RANSAC.run.  OpenCV's semantics (batchDistance,
// K = 1): for every query descriptor i the train descriptor with the smallest Hamming distance, the FIRST such
// index on ties (strict '<' over ascending j); with cross-check the roles are swapped for a second pass and
// match (i, j) survives iff best_of_train[j] == i.  Integer work: results are bit-exact.
#pragma once

#include "common.cuh"

namespace b200geo {

constexpr int HM_THREADS = 128;      // queries per CTA (one per thread)
constexpr int HM_TILE = 128;         // candidate descriptors staged per shared-memory tile
constexpr int HM_MAX_WORDS = 64;     // descriptor length limit: 256 bytes

// Grid: (ceil(max(nA, nB) / HM_THREADS), cross ? 2 : 1).  blockIdx.y = 0: queries = A, candidates = B;
// blockIdx.y = 1: queries = B, candidates = A.  best[y][q] = first index of the minimum, dist[y][q] its distance.
template <int WORDS>
__global__ void __launch_bounds__(HM_THREADS)
hamming_best_kernel(const uint32_t *__restrict__ A, int nA, const uint32_t *__restrict__ B, int nB, int words_rt,
                    int32_t *__restrict__ bestA, int32_t *__restrict__ distA, int32_t *__restrict__ bestB)
{
    const int words = WORDS ? WORDS : words_rt;
    const uint32_t *Q = blockIdx.y ? B : A, *C = blockIdx.y ? A : B;
    const int nQ = blockIdx.y ? nB : nA, nC = blockIdx.y ? nA : nB;
    if ((int)(blockIdx.x * HM_THREADS) >= nQ) return;                  // uniform per CTA
    extern __shared__ uint32_t hm_tile[];                              // HM_TILE x words
    const int q = blockIdx.x * HM_THREADS + threadIdx.x;
    uint32_t qd[WORDS ? WORDS : HM_MAX_WORDS];
    if (q < nQ)
#pragma unroll
        for (int w = 0; w < (WORDS ? WORDS : HM_MAX_WORDS); ++w)
            if (w < words) qd[w] = __ldg(Q + (size_t)q * words + w);
    int best = -1, bestd = 0x7fffffff;
    for (int c0 = 0; c0 < nC; c0 += HM_TILE) {
        const int m = min(HM_TILE, nC - c0);
        __syncthreads();
        for (int t = threadIdx.x; t < m * words; t += HM_THREADS) hm_tile[t] = __ldg(C + (size_t)c0 * words + t);
        __syncthreads();
        if (q < nQ)
            for (int j = 0; j < m; ++j) {
                int d = 0;
#pragma unroll
                for (int w = 0; w < (WORDS ? WORDS : HM_MAX_WORDS); ++w)
                    if (w < words) d += __popc(qd[w] ^ hm_tile[j * words + w]);      // broadcast reads
                if (d < bestd) { bestd = d; best = c0 + j; }
            }
    }
    if (q < nQ) {
        if (blockIdx.y == 0) { bestA[q] = best; distA[q] = bestd; }
        else bestB[q] = best;
    }
}

// cross-check: trainIdx[i] = -1 unless best_of_train[bestA[i]] == i
__global__ void hamming_cross_kernel(int32_t *__restrict__ bestA, const int32_t *__restrict__ bestB, int nA)
{
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= nA) return;
    const int j = bestA[i];
    if (j >= 0 && bestB[j] != i) bestA[i] = -1;
}

}  // namespace b200geo
