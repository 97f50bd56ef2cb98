#include <cstdio>
#include <cstdlib>
#include <algorithm>
// A: original
__device__ __forceinline__ int arcA(const int (&d)[16]) {
    int lo2[16], hi2[16], lo4[16], hi4[16];
#pragma unroll
    for (int k = 0; k < 16; k++) { lo2[k] = min(d[k], d[(k + 1) & 15]); hi2[k] = max(d[k], d[(k + 1) & 15]); }
#pragma unroll
    for (int k = 0; k < 16; k++) { lo4[k] = min(lo2[k], lo2[(k + 2) & 15]); hi4[k] = max(hi2[k], hi2[(k + 2) & 15]); }
    int best = -256;
#pragma unroll
    for (int k = 0; k < 16; k++) {
        int mn = min(min(lo4[k], lo4[(k + 4) & 15]), d[(k + 8) & 15]);
        int mx = max(max(hi4[k], hi4[(k + 4) & 15]), d[(k + 8) & 15]);
        best = max(best, max(mn, -mx));
    }
    return best;
}
// B: only min networks, on d and on -d
__device__ __forceinline__ int min9max(const int (&d)[16]) {
    int lo2[16], lo4[16];
#pragma unroll
    for (int k = 0; k < 16; k++) lo2[k] = min(d[k], d[(k + 1) & 15]);
#pragma unroll
    for (int k = 0; k < 16; k++) lo4[k] = min(lo2[k], lo2[(k + 2) & 15]);
    int best = -256;
#pragma unroll
    for (int k = 0; k < 16; k++) best = max(best, min(min(lo4[k], lo4[(k + 4) & 15]), d[(k + 8) & 15]));
    return best;
}
__device__ __forceinline__ int arcB(const int (&d)[16]) {
    int nd[16];
#pragma unroll
    for (int k = 0; k < 16; k++) nd[k] = -d[k];
    return max(min9max(d), min9max(nd));
}
// C: explicit comparisons with selects
__device__ __forceinline__ int mn(int a, int b) { return a < b ? a : b; }
__device__ __forceinline__ int mx(int a, int b) { return a > b ? a : b; }
__device__ __forceinline__ int arcC(const int (&d)[16]) {
    int best = -256;
#pragma unroll
    for (int k = 0; k < 16; k++) {
        int lo = d[k], hi = d[k];
#pragma unroll
        for (int j = 1; j < 9; j++) { lo = mn(lo, d[(k + j) & 15]); hi = mx(hi, d[(k + j) & 15]); }
        best = mx(best, mx(lo, -hi));
    }
    return best;
}
// D: packed 16-bit SIMD (two polarities in one register): lanes (d, -d), min network, then max of halves
__device__ __forceinline__ int arcD(const int (&d)[16]) {
    unsigned p[16], l2[16], l4[16];
#pragma unroll
    for (int k = 0; k < 16; k++) p[k] = ((unsigned)(d[k] & 0xFFFF)) | ((unsigned)((-d[k]) & 0xFFFF) << 16);
#pragma unroll
    for (int k = 0; k < 16; k++) l2[k] = __vmins2(p[k], p[(k + 1) & 15]);
#pragma unroll
    for (int k = 0; k < 16; k++) l4[k] = __vmins2(l2[k], l2[(k + 2) & 15]);
    unsigned best = 0x80008000u;   // (-32768, -32768)
#pragma unroll
    for (int k = 0; k < 16; k++) best = __vmaxs2(best, __vmins2(__vmins2(l4[k], l4[(k + 4) & 15]), p[(k + 8) & 15]));
    int a = (short)(best & 0xFFFF), b = (short)(best >> 16);
    return a > b ? a : b;
}
template <int V> __global__ void k(const int* din, int n, int* out) {
    int i = blockIdx.x * blockDim.x + threadIdx.x; if (i >= n) return;
    int d[16];
    for (int q = 0; q < 16; q++) d[q] = din[i * 16 + q];
    out[i] = V == 0 ? arcA(d) : V == 1 ? arcB(d) : V == 2 ? arcC(d) : arcD(d);
}
int main() {
    const int n = 4096; static int h[n * 16]; static int ref[n], got[n];
    srand(1);
    for (int i = 0; i < n; i++) {
        int base = rand() % 3;
        for (int q = 0; q < 16; q++) h[i * 16 + q] = base == 0 ? rand() % 511 - 255 : base == 1 ? rand() % 60 : -(rand() % 40) + (q > 8 ? 90 : 0);
        int best = -256;
        for (int k = 0; k < 16; k++) { int lo = 999, hi = -999; for (int j = 0; j < 9; j++) { int v = h[i * 16 + ((k + j) & 15)]; lo = std::min(lo, v); hi = std::max(hi, v); } best = std::max(best, std::max(lo, -hi)); }
        ref[i] = best;
    }
    int *d, *o; cudaMalloc(&d, sizeof(h)); cudaMalloc(&o, sizeof(got)); cudaMemcpy(d, h, sizeof(h), cudaMemcpyHostToDevice);
    for (int v = 0; v < 4; v++) {
        if (v == 0) k<0><<<n / 128, 128>>>(d, n, o); if (v == 1) k<1><<<n / 128, 128>>>(d, n, o);
        if (v == 2) k<2><<<n / 128, 128>>>(d, n, o); if (v == 3) k<3><<<n / 128, 128>>>(d, n, o);
        cudaMemcpy(got, o, sizeof(got), cudaMemcpyDeviceToHost);
        int bad = 0; for (int i = 0; i < n; i++) bad += got[i] != ref[i];
        printf("variant %c: %d / %d mismatches (%s)\n", 'A' + v, bad, n, cudaGetErrorString(cudaGetLastError()));
    }
}
