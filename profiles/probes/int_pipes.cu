// Register-only throughput probes for the integer instructions the image kernels lean on (B200, sm_100a):
// IMAD.HI (__umulhi) vs IMAD + SHF, IDP (dp4a / dp2a), VABSDIFF4, PRMT.  Prints giga-instructions (thread level) per second.
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o int_pipes int_pipes.cu && ./int_pipes
#include <cstdio>
#include <cuda_runtime.h>

template <int OP>
__global__ void probe(unsigned seed, int iters, unsigned* sink)
{
    unsigned a = seed + threadIdx.x, b = seed * 3 + blockIdx.x, c = seed ^ 0x9e3779b9u, d = a ^ b;
    unsigned x0 = a, x1 = b, x2 = c, x3 = d, x4 = a + 1, x5 = b + 2, x6 = c + 3, x7 = d + 4;
#pragma unroll 1
    for (int i = 0; i < iters; i++) {
#pragma unroll
        for (int k = 0; k < 8; k++) {
            if (OP == 0) { x0 = __umulhi(x0, a); x1 = __umulhi(x1, b); x2 = __umulhi(x2, c); x3 = __umulhi(x3, d); x4 = __umulhi(x4, a); x5 = __umulhi(x5, b); x6 = __umulhi(x6, c); x7 = __umulhi(x7, d); }
            if (OP == 1) { x0 = x0 * a + b; x1 = x1 * b + c; x2 = x2 * c + d; x3 = x3 * d + a; x4 = x4 * a + c; x5 = x5 * b + d; x6 = x6 * c + a; x7 = x7 * d + b; }
            if (OP == 2) { x0 = __dp4a(x0, a, b); x1 = __dp4a(x1, b, c); x2 = __dp4a(x2, c, d); x3 = __dp4a(x3, d, a); x4 = __dp4a(x4, a, c); x5 = __dp4a(x5, b, d); x6 = __dp4a(x6, c, a); x7 = __dp4a(x7, d, b); }
            if (OP == 3) { x0 = __vabsdiffu4(x0, a); x1 = __vabsdiffu4(x1, b); x2 = __vabsdiffu4(x2, c); x3 = __vabsdiffu4(x3, d); x4 = __vabsdiffu4(x4, a); x5 = __vabsdiffu4(x5, b); x6 = __vabsdiffu4(x6, c); x7 = __vabsdiffu4(x7, d); }
            if (OP == 4) { x0 = __byte_perm(x0, a, b); x1 = __byte_perm(x1, b, c); x2 = __byte_perm(x2, c, d); x3 = __byte_perm(x3, d, a); x4 = __byte_perm(x4, a, c); x5 = __byte_perm(x5, b, d); x6 = __byte_perm(x6, c, a); x7 = __byte_perm(x7, d, b); }
            if (OP == 5) { x0 = __vminu2(x0, a) + 1; x1 = __vminu2(x1, b) + 1; x2 = __vminu2(x2, c) + 1; x3 = __vminu2(x3, d) + 1; x4 = __vminu2(x4, a) + 1; x5 = __vminu2(x5, b) + 1; x6 = __vminu2(x6, c) + 1; x7 = __vminu2(x7, d) + 1; }
            if (OP == 6) { x0 = __funnelshift_r(x0, a, 8) ^ b; x1 = __funnelshift_r(x1, b, 8) ^ c; x2 = __funnelshift_r(x2, c, 8) ^ d; x3 = __funnelshift_r(x3, d, 8) ^ a; x4 = __funnelshift_r(x4, a, 8) ^ c; x5 = __funnelshift_r(x5, b, 8) ^ d; x6 = __funnelshift_r(x6, c, 8) ^ a; x7 = __funnelshift_r(x7, d, 8) ^ b; }
        }
    }
    if ((x0 ^ x1 ^ x2 ^ x3 ^ x4 ^ x5 ^ x6 ^ x7) == 0x12345u) sink[0] = x0;
}

template <int OP>
static double run(const char* name, double per_iter)
{
    unsigned* sink; cudaMalloc(&sink, 4);
    cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
    const int blocks = 148 * 8, threads = 256, iters = 4000;
    probe<OP><<<blocks, threads>>>(1, 100, sink);
    cudaEventRecord(e0);
    probe<OP><<<blocks, threads>>>(7, iters, sink);
    cudaEventRecord(e1); cudaEventSynchronize(e1);
    float ms; cudaEventElapsedTime(&ms, e0, e1);
    const double g = (double)blocks * threads * iters * per_iter / (ms * 1e-3) / 1e9;
    printf("%-28s %8.1f G thread-instr/s  (%.1f per clk per SM at 1.965 GHz)\n", name, g, g / 148 / 1.965);
    cudaFree(sink);
    return g;
}

int main()
{
    run<0>("IMAD.HI (__umulhi)", 64);
    run<1>("IMAD (mul-add)", 64);
    run<2>("IDP.4A (__dp4a)", 64);
    run<3>("VABSDIFF4.U8", 64);
    run<4>("PRMT", 64);
    run<5>("VIMNMX.U16x2 + IADD", 128);
    run<6>("SHF + LOP3", 128);
    return 0;
}
