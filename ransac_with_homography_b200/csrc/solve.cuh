// solve.cuh -- K1: hypothesis generation + 4-point DLT solve, one homography per thread.
//
// Replaces calc_corresp + calcHomography (homography.py:4-14, 71-88) as called per RANSAC iteration
// (ransac.py:177-180).  The reference takes the null vector of the raw 8x9 float32 matrix by LAPACK
// SVD; for four points in general position that vector, scaled to h22 = 1, is the solution of the
// 8x8 system of homography.py:16-28.  Per hypothesis that system is solved in closed form
// (dlt4_closed: two square-to-quadrilateral maps composed, ~200 flops, registers only) in
// similarity-normalised coordinates (translate to the centroid, scale to unit box), then
// de-normalised and divided by h22.  dlt4_solve (Gaussian elimination with partial pivoting on the
// 8x9 augmented matrix) serves the single explicit Scanner system and the refit.
// T = double is the default (FP64 is 1:2 on B200); T = float is the literal "FP32 registers" variant.
#pragma once

#include "common.cuh"

namespace b200geo {

template <typename T> struct Real;
template <> struct Real<float> {
    __device__ static __forceinline__ float abs(float v) { return fabsf(v); }
    __device__ static __forceinline__ float max(float a, float b) { return fmaxf(a, b); }
};
template <> struct Real<double> {
    __device__ static __forceinline__ double abs(double v) { return fabs(v); }
    __device__ static __forceinline__ double max(double a, double b) { return fmax(a, b); }
};

// Solve the n x n system held in A[n][n+1] (augmented) in place; returns false on a zero pivot.
template <typename T, int N>
__device__ __forceinline__ bool gauss_solve(T (&A)[N][N + 1], T (&x)[N])
{
    bool ok = true;
#pragma unroll
    for (int c = 0; c < N; ++c) {
        // partial pivoting: bubble the largest |A[r][c]|, r >= c, into row c
#pragma unroll
        for (int r = c + 1; r < N; ++r) {
            bool sw = Real<T>::abs(A[r][c]) > Real<T>::abs(A[c][c]);
#pragma unroll
            for (int j = c; j <= N; ++j) {
                T a = A[c][j], b = A[r][j];
                A[c][j] = sw ? b : a;
                A[r][j] = sw ? a : b;
            }
        }
        ok = ok && (Real<T>::abs(A[c][c]) > T(0));
        T inv = T(1) / A[c][c];
#pragma unroll
        for (int r = c + 1; r < N; ++r) {
            T f = A[r][c] * inv;
#pragma unroll
            for (int j = c + 1; j <= N; ++j) A[r][j] -= f * A[c][j];
        }
    }
#pragma unroll
    for (int c = N - 1; c >= 0; --c) {
        T s = A[c][N];
#pragma unroll
        for (int j = c + 1; j < N; ++j) s -= A[c][j] * x[j];
        x[c] = s / A[c][c];
    }
    return ok;
}

// 4 correspondences (x, y) -> (X, Y)  =>  H (row major, h22 = 1).  Returns false if degenerate.
template <typename T>
__device__ __forceinline__ bool dlt4_solve(const T (&x)[4], const T (&y)[4], const T (&X)[4], const T (&Y)[4],
                                           T (&H)[9])
{
    // similarity normalisation of both point sets
    T cx = (x[0] + x[1] + x[2] + x[3]) * T(0.25), cy = (y[0] + y[1] + y[2] + y[3]) * T(0.25);
    T cX = (X[0] + X[1] + X[2] + X[3]) * T(0.25), cY = (Y[0] + Y[1] + Y[2] + Y[3]) * T(0.25);
    T ds = T(0), dd = T(0);
#pragma unroll
    for (int i = 0; i < 4; ++i) {
        ds = Real<T>::max(ds, Real<T>::max(Real<T>::abs(x[i] - cx), Real<T>::abs(y[i] - cy)));
        dd = Real<T>::max(dd, Real<T>::max(Real<T>::abs(X[i] - cX), Real<T>::abs(Y[i] - cY)));
    }
    bool ok = (ds > T(0)) && (dd > T(0));
    T ss = T(1) / ds, sd = T(1) / dd;

    T A[8][9];
#pragma unroll
    for (int i = 0; i < 4; ++i) {
        T u = (x[i] - cx) * ss, v = (y[i] - cy) * ss;
        T U = (X[i] - cX) * sd, V = (Y[i] - cY) * sd;
        A[2 * i][0] = u;  A[2 * i][1] = v;  A[2 * i][2] = T(1);
        A[2 * i][3] = T(0); A[2 * i][4] = T(0); A[2 * i][5] = T(0);
        A[2 * i][6] = -u * U; A[2 * i][7] = -v * U; A[2 * i][8] = U;
        A[2 * i + 1][0] = T(0); A[2 * i + 1][1] = T(0); A[2 * i + 1][2] = T(0);
        A[2 * i + 1][3] = u;  A[2 * i + 1][4] = v;  A[2 * i + 1][5] = T(1);
        A[2 * i + 1][6] = -u * V; A[2 * i + 1][7] = -v * V; A[2 * i + 1][8] = V;
    }
    T g[8];
    ok = gauss_solve<T, 8>(A, g) && ok;

    // H = Td^-1 * G * Ts,  Ts = [ss 0 -ss*cx; 0 ss -ss*cy; 0 0 1],  Td^-1 = [dd 0 cX; 0 dd cY; 0 0 1]
    T m[9];
    m[0] = g[0] * ss; m[1] = g[1] * ss; m[2] = g[2] - (g[0] * cx + g[1] * cy) * ss;
    m[3] = g[3] * ss; m[4] = g[4] * ss; m[5] = g[5] - (g[3] * cx + g[4] * cy) * ss;
    m[6] = g[6] * ss; m[7] = g[7] * ss; m[8] = T(1) - (g[6] * cx + g[7] * cy) * ss;
    T n[9];
#pragma unroll
    for (int j = 0; j < 3; ++j) {
        n[j] = m[j] * dd + cX * m[6 + j];
        n[3 + j] = m[3 + j] * dd + cY * m[6 + j];
        n[6 + j] = m[6 + j];
    }
    T inv = T(1) / n[8];
#pragma unroll
    for (int j = 0; j < 8; ++j) H[j] = n[j] * inv;
    H[8] = T(1);
#pragma unroll
    for (int j = 0; j < 8; ++j) ok = ok && (Real<T>::abs(H[j]) <= T(3.0e38));   // false for NaN / inf
    return ok;
}

// Unit square (0,0) (1,0) (1,1) (0,1) -> quadrilateral p0 p1 p2 p3 (Heckbert's closed form of the
// same 8x8 system): S = [[a b c] [d e f] [g h 1]].  Returns the denominator whose vanishing means
// p1, p2, p3 are collinear.
template <typename T>
__device__ __forceinline__ T square_to_quad(const T (&x)[4], const T (&y)[4], T (&S)[9])
{
    const T dx1 = x[1] - x[2], dx2 = x[3] - x[2], dx3 = x[0] - x[1] + x[2] - x[3];
    const T dy1 = y[1] - y[2], dy2 = y[3] - y[2], dy3 = y[0] - y[1] + y[2] - y[3];
    const T den = dx1 * dy2 - dy1 * dx2;
    const T r = T(1) / den;
    const T g = (dx3 * dy2 - dy3 * dx2) * r, h = (dx1 * dy3 - dy1 * dx3) * r;
    S[0] = x[1] - x[0] + g * x[1]; S[1] = x[3] - x[0] + h * x[3]; S[2] = x[0];
    S[3] = y[1] - y[0] + g * y[1]; S[4] = y[3] - y[0] + h * y[3]; S[5] = y[0];
    S[6] = g; S[7] = h; S[8] = T(1);
    return den;
}

// 4 correspondences -> H by composing two square-to-quad maps: H = S_dst * adj(S_src), h22 = 1.
// The solution of the same 8x8 system dlt4_solve() eliminates (~1200 flops, 72-entry matrix), in
// ~200 flops and a few dozen registers; used per RANSAC hypothesis.  Similarity normalisation as in
// dlt4_solve keeps the differences well scaled.
template <typename T>
__device__ __forceinline__ bool dlt4_closed(const T (&x)[4], const T (&y)[4], const T (&X)[4], const T (&Y)[4], T (&H)[9])
{
    T cx = (x[0] + x[1] + x[2] + x[3]) * T(0.25), cy = (y[0] + y[1] + y[2] + y[3]) * T(0.25);
    T cX = (X[0] + X[1] + X[2] + X[3]) * T(0.25), cY = (Y[0] + Y[1] + Y[2] + Y[3]) * T(0.25);
    T ds = T(0), dd = T(0);
#pragma unroll
    for (int i = 0; i < 4; ++i) {
        ds = Real<T>::max(ds, Real<T>::max(Real<T>::abs(x[i] - cx), Real<T>::abs(y[i] - cy)));
        dd = Real<T>::max(dd, Real<T>::max(Real<T>::abs(X[i] - cX), Real<T>::abs(Y[i] - cY)));
    }
    bool ok = (ds > T(0)) && (dd > T(0));
    const T ss = T(1) / ds, sd = T(1) / dd;
    T u[4], v[4], U[4], V[4];
#pragma unroll
    for (int i = 0; i < 4; ++i) {
        u[i] = (x[i] - cx) * ss; v[i] = (y[i] - cy) * ss;
        U[i] = (X[i] - cX) * sd; V[i] = (Y[i] - cY) * sd;
    }
    T A[9], B[9];
    const T da = square_to_quad<T>(u, v, A), db = square_to_quad<T>(U, V, B);
    ok = ok && (Real<T>::abs(da) > T(0)) && (Real<T>::abs(db) > T(0));
    // adj(A)
    T a[9];
    a[0] = A[4] * A[8] - A[5] * A[7]; a[1] = A[2] * A[7] - A[1] * A[8]; a[2] = A[1] * A[5] - A[2] * A[4];
    a[3] = A[5] * A[6] - A[3] * A[8]; a[4] = A[0] * A[8] - A[2] * A[6]; a[5] = A[2] * A[3] - A[0] * A[5];
    a[6] = A[3] * A[7] - A[4] * A[6]; a[7] = A[1] * A[6] - A[0] * A[7]; a[8] = A[0] * A[4] - A[1] * A[3];
    // g = B * adj(A): normalised-coordinate homography (any scale)
    T g[9];
#pragma unroll
    for (int r = 0; r < 3; ++r)
#pragma unroll
        for (int c = 0; c < 3; ++c) g[3 * r + c] = B[3 * r] * a[c] + B[3 * r + 1] * a[3 + c] + B[3 * r + 2] * a[6 + c];
    // H = Td^-1 * g * Ts,  Ts = [ss 0 -ss*cx; 0 ss -ss*cy; 0 0 1],  Td^-1 = [dd 0 cX; 0 dd cY; 0 0 1]
    T m[9];
#pragma unroll
    for (int r = 0; r < 3; ++r) {
        m[3 * r] = g[3 * r] * ss; m[3 * r + 1] = g[3 * r + 1] * ss;
        m[3 * r + 2] = g[3 * r + 2] - (g[3 * r] * cx + g[3 * r + 1] * cy) * ss;
    }
    T n[9];
#pragma unroll
    for (int j = 0; j < 3; ++j) {
        n[j] = m[j] * dd + cX * m[6 + j];
        n[3 + j] = m[3 + j] * dd + cY * m[6 + j];
        n[6 + j] = m[6 + j];
    }
    const T inv = T(1) / n[8];
#pragma unroll
    for (int j = 0; j < 8; ++j) H[j] = n[j] * inv;
    H[8] = T(1);
#pragma unroll
    for (int j = 0; j < 8; ++j) ok = ok && (Real<T>::abs(H[j]) <= T(3.0e38));   // false for NaN / inf
    return ok;
}

// 3x3 inverse by the adjugate (np.linalg.inv(self.val), ransac.py:74)
template <typename T>
__device__ __forceinline__ void invert3(const T (&h)[9], T (&o)[9])
{
    T c0 = h[4] * h[8] - h[5] * h[7];
    T c1 = h[5] * h[6] - h[3] * h[8];
    T c2 = h[3] * h[7] - h[4] * h[6];
    T det = h[0] * c0 + h[1] * c1 + h[2] * c2;
    T r = T(1) / det;
    o[0] = c0 * r; o[1] = (h[2] * h[7] - h[1] * h[8]) * r; o[2] = (h[1] * h[5] - h[2] * h[4]) * r;
    o[3] = c1 * r; o[4] = (h[0] * h[8] - h[2] * h[6]) * r; o[5] = (h[2] * h[3] - h[0] * h[5]) * r;
    o[6] = c2 * r; o[7] = (h[1] * h[6] - h[0] * h[7]) * r; o[8] = (h[0] * h[4] - h[1] * h[3]) * r;
}

// Solve the sample set `id` over the correspondence rows `pts`; writes float32 H (and Hinv).
// Returns the degenerate flag.
template <typename T>
__device__ __forceinline__ bool solve_sample(const float4 *__restrict__ pts, int4 id, float (&Hf)[9],
                                             float (&Hif)[9], bool want_inv)
{
    float4 p0 = __ldg(pts + id.x), p1 = __ldg(pts + id.y), p2 = __ldg(pts + id.z), p3 = __ldg(pts + id.w);
    T x[4] = {T(p0.x), T(p1.x), T(p2.x), T(p3.x)};
    T y[4] = {T(p0.y), T(p1.y), T(p2.y), T(p3.y)};
    T X[4] = {T(p0.z), T(p1.z), T(p2.z), T(p3.z)};
    T Y[4] = {T(p0.w), T(p1.w), T(p2.w), T(p3.w)};
    bool dup = (id.x == id.y) | (id.x == id.z) | (id.x == id.w) | (id.y == id.z) | (id.y == id.w) | (id.z == id.w);
    T H[9];
    bool ok = dlt4_closed<T>(x, y, X, Y, H) && !dup;
    const float qnan = __int_as_float(0x7fc00000);
    if (want_inv) {
        T Hi[9];
        invert3<T>(H, Hi);
#pragma unroll
        for (int j = 0; j < 9; ++j) Hif[j] = ok ? (float)Hi[j] : qnan;
    }
#pragma unroll
    for (int j = 0; j < 9; ++j) Hf[j] = ok ? (float)H[j] : qnan;
    return !ok;
}

// K1: one thread per sample set.  Grid (ceil(K/128), B).
template <typename T>
__global__ void __launch_bounds__(128)
dlt4_kernel(const float4 *__restrict__ pts, int M, const int4 *__restrict__ idx, int k_begin, int k_end,
            int K_stride, float *__restrict__ H, float *__restrict__ Hinv, uint8_t *__restrict__ flags)
{
    int b = blockIdx.y;
    int k = k_begin + blockIdx.x * blockDim.x + threadIdx.x;
    if (k >= k_end) return;
    pts += (size_t)b * M;
    size_t o = (size_t)b * K_stride + k;
    int4 id = __ldg(idx + o);
    // indices outside [0, M) cannot come from randint(0, M); clamp so the gather stays in bounds
    bool bad = ((unsigned)id.x >= (unsigned)M) | ((unsigned)id.y >= (unsigned)M) | ((unsigned)id.z >= (unsigned)M) |
               ((unsigned)id.w >= (unsigned)M);
    if (bad) id = make_int4(0, 0, 0, 0);
    float Hf[9], Hif[9];
    bool deg = solve_sample<T>(pts, id, Hf, Hif, Hinv != nullptr);
#pragma unroll
    for (int j = 0; j < 9; ++j) H[o * 9 + j] = Hf[j];
    if (Hinv) {
#pragma unroll
        for (int j = 0; j < 9; ++j) Hinv[o * 9 + j] = Hif[j];
    }
    if (flags) flags[o] = (uint8_t)(deg || bad);
}

// One explicit system (Scanner: calcHomographyLinear(u, v), homography.py:90-105), float64.
__global__ void homography4_kernel(const double *__restrict__ uv, double *__restrict__ H, int *__restrict__ status)
{
    if (threadIdx.x != 0 || blockIdx.x != 0) return;
    double x[4], y[4], X[4], Y[4], h[9];
#pragma unroll
    for (int i = 0; i < 4; ++i) { x[i] = uv[2 * i]; y[i] = uv[2 * i + 1]; X[i] = uv[8 + 2 * i]; Y[i] = uv[8 + 2 * i + 1]; }
    bool ok = dlt4_solve<double>(x, y, X, Y, h);
#pragma unroll
    for (int j = 0; j < 9; ++j) H[j] = h[j];
    *status = ok ? 0 : B200GEO_ERR_SINGULAR;
}

}  // namespace b200geo
