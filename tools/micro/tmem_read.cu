// Micro-benchmark: how fast can one SM read its tensor memory (tcgen05.ld) into registers?
// Decides the floor of the top-k GEMM epilogue (a 128x256 fp32 accumulator = 128 KB per tile).
// build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o tmem_read tmem_read.cu ; run: ./tmem_read
#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>

__device__ __forceinline__ uint32_t smem_u32(const void *p) { return (uint32_t)__cvta_generic_to_shared(p); }

#define LD32(taddr, r)                                                                                              \
    asm volatile("tcgen05.ld.sync.aligned.32x32b.x32.b32 "                                                          \
                 "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, "                           \
                 "%16, %17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31}, [%32];"           \
                 : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]),   \
                   "=r"(r[8]), "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]),          \
                   "=r"(r[15]), "=r"(r[16]), "=r"(r[17]), "=r"(r[18]), "=r"(r[19]), "=r"(r[20]), "=r"(r[21]),        \
                   "=r"(r[22]), "=r"(r[23]), "=r"(r[24]), "=r"(r[25]), "=r"(r[26]), "=r"(r[27]), "=r"(r[28]),        \
                   "=r"(r[29]), "=r"(r[30]), "=r"(r[31])                                                             \
                 : "r"(taddr)                                                                                        \
                 : "memory")
#define LD16(taddr, r)                                                                                              \
    asm volatile("tcgen05.ld.sync.aligned.32x32b.x16.b32 "                                                          \
                 "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15}, [%16];"                    \
                 : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]),   \
                   "=r"(r[8]), "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]),          \
                   "=r"(r[15])                                                                                       \
                 : "r"(taddr)                                                                                        \
                 : "memory")
// 16 lanes x 256 bits, x8 repeats: 32 registers per thread, covers 16 lanes x 64 columns
#define LD16x256(taddr, r)                                                                                          \
    asm volatile("tcgen05.ld.sync.aligned.16x256b.x8.b32 "                                                          \
                 "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, "                           \
                 "%16, %17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31}, [%32];"           \
                 : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]),   \
                   "=r"(r[8]), "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]),          \
                   "=r"(r[15]), "=r"(r[16]), "=r"(r[17]), "=r"(r[18]), "=r"(r[19]), "=r"(r[20]), "=r"(r[21]),        \
                   "=r"(r[22]), "=r"(r[23]), "=r"(r[24]), "=r"(r[25]), "=r"(r[26]), "=r"(r[27]), "=r"(r[28]),        \
                   "=r"(r[29]), "=r"(r[30]), "=r"(r[31])                                                             \
                 : "r"(taddr)                                                                                        \
                 : "memory")
#define LDWAIT() asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory")

// variant 0: 32x32b.x32, wait after every load; 1: 32x32b.x32, wait after every 2nd load (two in flight);
// 2: 32x32b.x16; 3: 16x256b.x8
template <int VARIANT>
__global__ void __launch_bounds__(512, 1) k_tmem_read(int iters, long long *cycles, unsigned *sink) {
    __shared__ uint32_t s_tmem;
    const int warp = threadIdx.x >> 5;
    if (warp == 0) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(&s_tmem)), "r"(512)
                     : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
    }
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    const uint32_t base = s_tmem + ((uint32_t)((warp & 3) * 32) << 16);
    uint32_t a[32], b[32];
    unsigned acc = 0;
    __syncthreads();
    const long long t0 = clock64();
    for (int i = 0; i < iters; i++) {
        const uint32_t col = (uint32_t)((i * 64 + (warp >> 2) * 32) & 448);
        if (VARIANT == 0) {
            LD32(base + col, a);
            LDWAIT();
            acc ^= a[0] ^ a[31];
            LD32(base + col + 32, b);
            LDWAIT();
            acc ^= b[0] ^ b[31];
        } else if (VARIANT == 1) {
            LD32(base + col, a);
            LD32(base + col + 32, b);
            LDWAIT();
            acc ^= a[0] ^ a[31] ^ b[0] ^ b[31];
        } else if (VARIANT == 2) {
            LD16(base + col, a);
            LD16(base + col + 16, (a + 16));
            LD16(base + col + 32, b);
            LD16(base + col + 48, (b + 16));
            LDWAIT();
            acc ^= a[0] ^ a[31] ^ b[0] ^ b[31];
        } else {
            LD16x256(base + col, a);
            LD16x256(base + col + ((uint32_t)16 << 16), b);
            LDWAIT();
            acc ^= a[0] ^ a[31] ^ b[0] ^ b[31];
        }
    }
    __syncthreads();
    const long long t1 = clock64();
    if (threadIdx.x == 0) cycles[blockIdx.x] = t1 - t0;
    if (acc == 0x12345u) sink[0] = acc;
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    if (warp == 0) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(s_tmem), "r"(512) : "memory");
}

template <int V>
void run(const char *name, int warps, int grid) {
    long long *cyc;
    unsigned *sink;
    cudaMalloc(&cyc, sizeof(long long) * grid);
    cudaMalloc(&sink, 4);
    const int iters = 4000;
    k_tmem_read<V><<<grid, warps * 32>>>(10, cyc, sink);
    k_tmem_read<V><<<grid, warps * 32>>>(iters, cyc, sink);
    cudaError_t e = cudaDeviceSynchronize();
    long long h = 0;
    cudaMemcpy(&h, cyc, sizeof(h), cudaMemcpyDeviceToHost);
    // every warp reads 2 x 4 KB per iteration (32 lanes x 64 columns x 4 B)
    const double bytes = (double)iters * warps * 8192.0;
    printf("%-28s warps=%2d grid=%3d  %8lld cycles  %7.1f B/clk/SM  (%s)\n", name, warps, grid, h, bytes / (double)h,
           cudaGetErrorString(e));
    cudaFree(cyc);
    cudaFree(sink);
}

int main() {
    for (int grid : {1, 148}) {
        for (int w : {1, 4, 8, 16}) {
            run<0>("32x32b.x32 wait each", w, grid);
            run<1>("32x32b.x32 two in flight", w, grid);
            run<2>("32x32b.x16 four in flight", w, grid);
            run<3>("16x256b.x8 two in flight", w, grid);
        }
    }
    return 0;
}
