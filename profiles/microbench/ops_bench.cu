// Integer-pipe throughput microbenchmark for the ops the fused yuv->rgb kernel is built from.
// nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o ops_bench ops_bench.cu ; ./ops_bench
// Reports warp-instructions per clock per SM (4.0 = one per SMSP per clock = issue limit).
#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>

#define ITERS 4096
#define NACC 8

template <int OP>
__device__ __forceinline__ void op(int &a, int b, int c)
{
    if (OP == 0) asm volatile("mad.lo.s32 %0, %0, %1, %2;" : "+r"(a) : "r"(b), "r"(c));
    if (OP == 1) asm volatile("add.s32 %0, %0, %1;" : "+r"(a) : "r"(b));
    if (OP == 2) asm volatile("prmt.b32 %0, %0, %1, 0x7632;" : "+r"(a) : "r"(b));
    if (OP == 3) asm volatile("shr.s32 %0, %0, 3;" : "+r"(a));
    if (OP == 4) asm volatile("min.relu.s32 %0, %0, %1;" : "+r"(a) : "r"(b));
    if (OP == 5) asm volatile("min.relu.s16x2 %0, %0, %1;" : "+r"(a) : "r"(b));
    if (OP == 6) asm volatile("dp4a.u32.s32 %0, %1, %2, %0;" : "+r"(a) : "r"(b), "r"(c));
    if (OP == 7) { long long t; asm volatile("mad.wide.s32 %0, %1, %2, %3;" : "=l"(t) : "r"(a), "r"(b), "l"((long long)c << 20)); a = (int)(t >> 32); }
    if (OP == 8) asm volatile("cvt.pack.sat.u8.s32.b32 %0, %1, %2, %0;" : "+r"(a) : "r"(b), "r"(c));
    if (OP == 9) asm volatile("mad.hi.s32 %0, %0, %1, %2;" : "+r"(a) : "r"(b), "r"(c));
    if (OP == 10) a = __viaddmin_s16x2_relu(a, b, c);
    if (OP == 11) { float f = __int_as_float(a); asm volatile("fma.rn.f32 %0, %0, %1, %2;" : "+f"(f) : "f"(__int_as_float(b)), "f"(__int_as_float(c))); a = __float_as_int(f); }
    if (OP == 12) { float f; asm volatile("cvt.rn.f32.s32 %0, %1;" : "=f"(f) : "r"(a)); a = __float_as_int(f) ^ b; }
    if (OP == 13) asm volatile("lop3.b32 %0, %0, %1, %2, 0x96;" : "+r"(a) : "r"(b), "r"(c));
    if (OP == 14) asm volatile("shf.r.wrap.b32 %0, %0, %1, 8;" : "+r"(a) : "r"(b));
    if (OP == 15) asm volatile("bfe.u32 %0, %0, 8, 8;" : "+r"(a));
    if (OP == 16) asm volatile("max.s32 %0, %0, %1;" : "+r"(a) : "r"(b));
    if (OP == 17) asm volatile("vadd.s32.s32.s32.sat %0, %0, %1;" : "+r"(a) : "r"(b));
    if (OP == 18) { asm volatile("mad.lo.s32 %0, %0, %1, %2;" : "+r"(a) : "r"(b), "r"(c)); }
}

template <int OP>
__global__ void __launch_bounds__(1024) k(int *out, int b, int c, long long *cycles)
{
    int acc[NACC];
    for (int i = 0; i < NACC; i++) acc[i] = threadIdx.x + i;
    __syncthreads();
    long long t0 = clock64();
    for (int it = 0; it < ITERS; it++) {
#pragma unroll
        for (int i = 0; i < NACC; i++) op<OP>(acc[i], b, c);
    }
    long long t1 = clock64();
    int s = 0;
    for (int i = 0; i < NACC; i++) s ^= acc[i];
    out[blockIdx.x * blockDim.x + threadIdx.x] = s;
    if (threadIdx.x == 0) cycles[blockIdx.x] = t1 - t0;
}

// mixed: IMAD + PRMT interleaved (do the two pipes dual-issue to 4/clk?)
__global__ void __launch_bounds__(1024) kmix(int *out, int b, int c, long long *cycles)
{
    int acc[NACC];
    for (int i = 0; i < NACC; i++) acc[i] = threadIdx.x + i;
    __syncthreads();
    long long t0 = clock64();
    for (int it = 0; it < ITERS; it++) {
#pragma unroll
        for (int i = 0; i < NACC; i += 2) { op<0>(acc[i], b, c); op<2>(acc[i + 1], b, c); }
    }
    long long t1 = clock64();
    int s = 0;
    for (int i = 0; i < NACC; i++) s ^= acc[i];
    out[blockIdx.x * blockDim.x + threadIdx.x] = s;
    if (threadIdx.x == 0) cycles[blockIdx.x] = t1 - t0;
}

template <int OPA, int OPB>
__global__ void __launch_bounds__(1024) kpair(int *out, int b, int c, long long *cycles)
{
    int acc[NACC];
    for (int i = 0; i < NACC; i++) acc[i] = threadIdx.x + i;
    __syncthreads();
    long long t0 = clock64();
    for (int it = 0; it < ITERS; it++) {
#pragma unroll
        for (int i = 0; i < NACC; i += 2) { op<OPA>(acc[i], b, c); op<OPB>(acc[i + 1], b, c); }
    }
    long long t1 = clock64();
    int s = 0;
    for (int i = 0; i < NACC; i++) s ^= acc[i];
    out[blockIdx.x * blockDim.x + threadIdx.x] = s;
    if (threadIdx.x == 0) cycles[blockIdx.x] = t1 - t0;
}

template <int OPA, int OPB>
void runpair(const char *name, int *out, long long *cyc)
{
    kpair<OPA, OPB><<<148, 1024>>>(out, 3, 5, cyc); cudaDeviceSynchronize();
    kpair<OPA, OPB><<<148, 1024>>>(out, 3, 5, cyc); cudaDeviceSynchronize();
    long long h[148]; cudaMemcpy(h, cyc, sizeof(h), cudaMemcpyDeviceToHost);
    double avg = 0; for (int i = 0; i < 148; i++) avg += h[i]; avg /= 148;
    printf("%-28s %7.3f warp-instr/clk/SM\n", name, 32.0 * ITERS * NACC / avg);
}

template <int OP>
void run(const char *name, int *out, long long *cyc)
{
    int sms = 148;
    k<OP><<<sms, 1024>>>(out, 3, 5, cyc);
    cudaDeviceSynchronize();
    k<OP><<<sms, 1024>>>(out, 3, 5, cyc);
    cudaDeviceSynchronize();
    long long h[148];
    cudaMemcpy(h, cyc, sizeof(h), cudaMemcpyDeviceToHost);
    double avg = 0;
    for (int i = 0; i < sms; i++) avg += h[i];
    avg /= sms;
    double winstr = 32.0 * ITERS * NACC;   // 32 warps/SM
    printf("%-28s %7.3f warp-instr/clk/SM  (%.0f cycles)\n", name, winstr / avg, avg);
}

int main()
{
    int *out; long long *cyc;
    cudaMalloc(&out, 148 * 1024 * sizeof(int));
    cudaMalloc(&cyc, 148 * sizeof(long long));
    run<0>("IMAD (mad.lo)", out, cyc);
    run<1>("IADD", out, cyc);
    run<2>("PRMT", out, cyc);
    run<3>("SHR imm", out, cyc);
    run<4>("min.relu.s32", out, cyc);
    run<5>("min.relu.s16x2", out, cyc);
    run<6>("dp4a.u32.s32", out, cyc);
    run<7>("mad.wide.s32 (+hi extract)", out, cyc);
    run<8>("cvt.pack.sat.u8.s32 (I2IP)", out, cyc);
    run<9>("mad.hi.s32", out, cyc);
    run<10>("viaddmin_s16x2_relu", out, cyc);
    run<11>("FFMA", out, cyc);
    run<12>("I2F + xor", out, cyc);
    run<13>("LOP3", out, cyc);
    run<14>("SHF funnel", out, cyc);
    run<15>("BFE", out, cyc);
    run<16>("max.s32", out, cyc);
    run<17>("vadd.sat", out, cyc);
    {
        kmix<<<148, 1024>>>(out, 3, 5, cyc); cudaDeviceSynchronize();
        kmix<<<148, 1024>>>(out, 3, 5, cyc); cudaDeviceSynchronize();
        long long h[148]; cudaMemcpy(h, cyc, sizeof(h), cudaMemcpyDeviceToHost);
        double avg = 0; for (int i = 0; i < 148; i++) avg += h[i]; avg /= 148;
        printf("%-28s %7.3f warp-instr/clk/SM\n", "IMAD+PRMT interleaved", 32.0 * ITERS * NACC / avg);
    }
    runpair<0, 6>("IMAD + IDP4A", out, cyc);
    runpair<2, 6>("PRMT + IDP4A", out, cyc);
    runpair<0, 4>("IMAD + VIMNMX", out, cyc);
    runpair<2, 4>("PRMT + VIMNMX", out, cyc);
    runpair<0, 5>("IMAD + VIMNMX.S16x2", out, cyc);
    runpair<2, 5>("PRMT + VIMNMX.S16x2", out, cyc);
    runpair<2, 14>("PRMT + SHF", out, cyc);
    runpair<0, 14>("IMAD + SHF", out, cyc);
    runpair<0, 8>("IMAD + I2IP", out, cyc);
    runpair<2, 8>("PRMT + I2IP", out, cyc);
    runpair<0, 13>("IMAD + LOP3", out, cyc);
    runpair<2, 13>("PRMT + LOP3", out, cyc);
    runpair<0, 11>("IMAD + FFMA", out, cyc);
    runpair<2, 11>("PRMT + FFMA", out, cyc);
    runpair<6, 11>("IDP4A + FFMA", out, cyc);
    runpair<0, 10>("IMAD + VIADDMNMX", out, cyc);
    runpair<2, 10>("PRMT + VIADDMNMX", out, cyc);
    runpair<0, 9>("IMAD + IMAD.HI", out, cyc);
    return 0;
}
