// tma_box_probe.cu - which 3-D u8 TMA box loads does a B200 accept?  Each variant runs in a forked child so that a faulting
// variant ("illegal instruction") does not poison the context of the others.
//   nvcc -gencode arch=compute_100a,code=sm_100a -o tma_box_probe tma_box_probe.cu -lcuda
#include <cuda.h>
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#include <sys/wait.h>
#include <unistd.h>
#include <vector>

__device__ __forceinline__ void mbar_init(unsigned long long* bar, unsigned count)
{
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"((unsigned)__cvta_generic_to_shared(bar)), "r"(count));
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
}
__device__ __forceinline__ void mbar_expect_tx(unsigned long long* bar, unsigned bytes)
{
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"((unsigned)__cvta_generic_to_shared(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_wait(unsigned long long* bar, unsigned parity)
{
    asm volatile("{\n\t.reg .pred p;\n\tWAIT_%=:\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n\t@p bra DONE_%=;\n\tbra WAIT_%=;\n\tDONE_%=:\n\t}"
                 ::"r"((unsigned)__cvta_generic_to_shared(bar)), "r"(parity) : "memory");
}
__device__ __forceinline__ void tma_load_3d(void* smem_dst, const CUtensorMap* map, int x, int y, int z, unsigned long long* bar)
{
    asm volatile("cp.async.bulk.tensor.3d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%2, %3, %4}], [%5];"
                 ::"r"((unsigned)__cvta_generic_to_shared(smem_dst)), "l"(map), "r"(x), "r"(y), "r"(z),
                   "r"((unsigned)__cvta_generic_to_shared(bar)) : "memory");
}

struct Maps { CUtensorMap m[4]; };

__global__ void k_probe(const __grid_constant__ Maps maps, int which, int x, int y, int z, int bw, int bh, int dst_off, uint8_t* out)
{
    extern __shared__ __align__(128) uint8_t smem[];
    __shared__ __align__(8) unsigned long long bar;
    if (threadIdx.x == 0) mbar_init(&bar, 1);
    __syncthreads();
    if (threadIdx.x == 0) {
        mbar_expect_tx(&bar, bw * bh);
        tma_load_3d(smem + dst_off, &maps.m[which], x, y, z, &bar);
    }
    mbar_wait(&bar, 0);
    for (int i = threadIdx.x; i < bw * bh; i += blockDim.x) out[i] = smem[dst_off + i];
}

typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*,
                                  const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

static int run_variant(const char* name, int w, int h, int pitch, int frames, int bw, int bh, int which, int x, int y, int z, int dst_off)
{
    void* p = nullptr; cudaDriverEntryPointQueryResult q;
    cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &q);
    EncodeTiledFn enc = (EncodeTiledFn)p;
    uint8_t* d; size_t bytes = (size_t)pitch * h * frames;
    cudaMalloc(&d, bytes);
    std::vector<uint8_t> hst(bytes);
    for (size_t i = 0; i < bytes; i++) hst[i] = (uint8_t)((i * 2654435761u) >> 13);
    cudaMemcpy(d, hst.data(), bytes, cudaMemcpyHostToDevice);
    alignas(64) Maps maps; memset(&maps, 0, sizeof(maps));
    for (int k = 0; k < 4; k++) {
        cuuint64_t dims[3] = {(cuuint64_t)w, (cuuint64_t)h, (cuuint64_t)frames};
        cuuint64_t strides[2] = {(cuuint64_t)pitch, (cuuint64_t)pitch * h};
        cuuint32_t box[3] = {(cuuint32_t)bw, (cuuint32_t)bh, 1}, es[3] = {1, 1, 1};
        CUresult r = enc(&maps.m[k], CU_TENSOR_MAP_DATA_TYPE_UINT8, 3, d, dims, strides, box, es, CU_TENSOR_MAP_INTERLEAVE_NONE,
                         CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_L2_128B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
        if (r != CUDA_SUCCESS) { printf("%-44s encode failed (%d)\n", name, (int)r); return 1; }
    }
    uint8_t* out; cudaMalloc(&out, bw * bh);
    cudaFuncSetAttribute(k_probe, cudaFuncAttributeMaxDynamicSharedMemorySize, 96 * 1024);
    k_probe<<<1, 64, 64 * 1024>>>(maps, which, x, y, z, bw, bh, dst_off, out);
    cudaError_t e = cudaDeviceSynchronize();
    if (e != cudaSuccess) { printf("%-44s KERNEL FAULT: %s\n", name, cudaGetErrorString(e)); return 1; }
    std::vector<uint8_t> got(bw * bh);
    cudaMemcpy(got.data(), out, bw * bh, cudaMemcpyDeviceToHost);
    int bad = 0;
    for (int r = 0; r < bh; r++)
        for (int c = 0; c < bw; c++) {
            const int gx = x + c, gy = y + r;
            const uint8_t want = (gx >= 0 && gx < w && gy >= 0 && gy < h) ? hst[(size_t)z * pitch * h + (size_t)gy * pitch + gx] : 0;
            bad += got[r * bw + c] != want;
        }
    printf("%-44s %s (%d wrong bytes)\n", name, bad ? "WRONG DATA" : "ok", bad);
    return bad != 0;
}

int main()
{
    struct V { const char* name; int w, h, pitch, frames, bw, bh, which, x, y, z, dst; } v[] = {
        {"box 48x37 x aligned, map 0", 752, 480, 768, 2, 48, 37, 0, 96, 50, 1, 0},
        {"box 48x37 x = 101 (unaligned), map 0", 752, 480, 768, 2, 48, 37, 0, 101, 50, 1, 0},
        {"box 48x37 x = 101, map 1 (dynamic index)", 752, 480, 768, 2, 48, 37, 1, 101, 50, 1, 0},
        {"box 48x37 x = 101, map 3", 752, 480, 768, 2, 48, 37, 3, 101, 50, 0, 0},
        {"box 48x37 w = 627 pitch 640", 627, 400, 640, 1, 48, 37, 0, 217, 33, 0, 0},
        {"box 48x37 near right edge (OOB fill)", 627, 400, 640, 1, 48, 37, 0, 600, 33, 0, 0},
        {"box 48x37 dst offset 1792", 752, 480, 768, 1, 48, 37, 0, 101, 50, 0, 1792},
        {"box 48x37 dst offset 3584", 752, 480, 768, 1, 48, 37, 0, 101, 50, 0, 3584},
        {"box 64x37 x = 101", 752, 480, 768, 1, 64, 37, 0, 101, 50, 0, 0},
        {"box 48x32 x = 101", 752, 480, 768, 1, 48, 32, 0, 101, 50, 0, 0},
        {"box 32x37 x = 101", 752, 480, 768, 1, 32, 37, 0, 101, 50, 0, 0},
        {"box 16x37 x = 101", 752, 480, 768, 1, 16, 37, 0, 101, 50, 0, 0},
    };
    for (auto& t : v) {
        fflush(stdout);
        pid_t pid = fork();
        if (pid == 0) { int rc = run_variant(t.name, t.w, t.h, t.pitch, t.frames, t.bw, t.bh, t.which, t.x, t.y, t.z, t.dst); fflush(stdout); _exit(rc); }
        int st = 0; waitpid(pid, &st, 0);
    }
    return 0;
}
