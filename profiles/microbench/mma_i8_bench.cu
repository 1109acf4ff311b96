// Is the tensor pipe worth it for the horizontal FIR?  
// Measures legacy warp-level integer MMA on sm_100a:
//   mma.sync.aligned.m16n8k32.row.col.s32.u8.s8.s32   (4096 MAC per warp instruction)
// alone, and fed by ldmatrix.x4 from shared memory the way a banded H-stage would feed it
// (A = 16 source rows x 32 bytes, B = 32 x 8 banded coefficient block kept in registers).
// nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o mma_i8_bench mma_i8_bench.cu
#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>

#define ITERS 2048
#define NACC 8

__device__ __forceinline__ void mma_u8s8(int (&c)[4], const uint32_t (&a)[4], const uint32_t (&b)[2])
{
    asm volatile("mma.sync.aligned.m16n8k32.row.col.s32.u8.s8.s32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};"
                 : "+r"(c[0]), "+r"(c[1]), "+r"(c[2]), "+r"(c[3])
                 : "r"(a[0]), "r"(a[1]), "r"(a[2]), "r"(a[3]), "r"(b[0]), "r"(b[1]));
}

__global__ void __launch_bounds__(1024) k_mma(int *out, long long *cycles, int warps_active)
{
    int c[NACC][4];
    uint32_t a[4], b[2];
    for (int i = 0; i < NACC; i++) for (int j = 0; j < 4; j++) c[i][j] = threadIdx.x + i + j;
    for (int j = 0; j < 4; j++) a[j] = threadIdx.x * 0x01010101u + j;
    b[0] = threadIdx.x; b[1] = ~threadIdx.x;
    __syncthreads();
    long long t0 = clock64();
    if ((threadIdx.x >> 5) < warps_active)
        for (int it = 0; it < ITERS; it++) {
#pragma unroll
            for (int i = 0; i < NACC; i++) mma_u8s8(c[i], a, b);
        }
    long long t1 = clock64();
    int s = 0;
    for (int i = 0; i < NACC; i++) for (int j = 0; j < 4; j++) s ^= c[i][j];
    out[blockIdx.x * blockDim.x + threadIdx.x] = s;
    if (threadIdx.x == 0) cycles[blockIdx.x] = t1 - t0;
}

// ldmatrix-fed: every iteration loads a fresh A fragment (16 rows x 32 B) from shared memory
__global__ void __launch_bounds__(1024) k_mma_ldm(int *out, long long *cycles)
{
    __shared__ __align__(128) unsigned char rows[32][8 * 144];      // per warp: 8 rows, 144-byte pitch (conflict-free)
    int c[NACC][4];
    uint32_t b[2];
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    for (int i = 0; i < 8 * 144; i += 32) if (i + lane < 8 * 144) rows[warp][i + lane] = (unsigned char)(i + lane);
    for (int i = 0; i < NACC; i++) for (int j = 0; j < 4; j++) c[i][j] = threadIdx.x + i + j;
    b[0] = threadIdx.x; b[1] = ~threadIdx.x;
    // ldmatrix.x4 row addresses: lanes 0-15 rows 0-15 at byte 0, lanes 16-31 rows 0-15 at byte 16
    const uint32_t base = (uint32_t)__cvta_generic_to_shared(&rows[warp][(lane & 7) * 144 + (lane >> 3) * 16]);
    __syncthreads();
    long long t0 = clock64();
    for (int it = 0; it < ITERS; it++) {
#pragma unroll
        for (int i = 0; i < NACC; i++) {
            uint32_t a[4];
            asm volatile("ldmatrix.sync.aligned.m8n8.x4.shared.b16 {%0,%1,%2,%3}, [%4];"
                         : "=r"(a[0]), "=r"(a[1]), "=r"(a[2]), "=r"(a[3]) : "r"(base + (i & 3) * 32));
            mma_u8s8(c[i], a, b);
        }
    }
    long long t1 = clock64();
    int s = 0;
    for (int i = 0; i < NACC; i++) for (int j = 0; j < 4; j++) s ^= c[i][j];
    out[blockIdx.x * blockDim.x + threadIdx.x] = s;
    if (threadIdx.x == 0) cycles[blockIdx.x] = t1 - t0;
}

static double avg_cycles(long long *cyc)
{
    long long h[148]; cudaMemcpy(h, cyc, sizeof(h), cudaMemcpyDeviceToHost);
    double avg = 0; for (int i = 0; i < 148; i++) avg += h[i];
    return avg / 148;
}

int main()
{
    int *out; long long *cyc;
    cudaMalloc(&out, 148 * 1024 * sizeof(int));
    cudaMalloc(&cyc, 148 * sizeof(long long));
    for (int w : {4, 8, 16, 32}) {
        k_mma<<<148, 1024>>>(out, cyc, w); cudaDeviceSynchronize();
        k_mma<<<148, 1024>>>(out, cyc, w); cudaDeviceSynchronize();
        double a = avg_cycles(cyc);
        double per_clk = (double)w * ITERS * NACC / a;
        printf("mma.m16n8k32.u8.s8  %2d warps/SM: %6.3f warp-MMA/clk/SM = %7.0f MAC/clk/SM  (IDP.4A pipe: 2 warp-instr/clk/SM = 256 MAC/clk/SM)\n",
               w, per_clk, per_clk * 4096);
    }
    k_mma_ldm<<<148, 1024>>>(out, cyc); cudaDeviceSynchronize();
    k_mma_ldm<<<148, 1024>>>(out, cyc); cudaDeviceSynchronize();
    double a = avg_cycles(cyc);
    printf("ldmatrix.x4 + mma    32 warps/SM: %6.3f pairs/clk/SM\n", 32.0 * ITERS * NACC / a);
    cudaError_t e = cudaGetLastError();
    printf("status: %s\n", cudaGetErrorString(e));
    return 0;
}
