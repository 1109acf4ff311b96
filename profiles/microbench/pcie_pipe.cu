// How fast can ONE synchronous host->device->host frame round trip be on this box?
// 4K yuv420p in (12.4 MB) / rgb24 out (24.9 MB), pinned host memory, banded over 3 streams.
#include <cstdio>
#include <chrono>
#include <cuda_runtime.h>
__global__ void touch(unsigned char *o, const unsigned char *i, size_t n) {
    size_t k = blockIdx.x * (size_t)blockDim.x + threadIdx.x;
    if (k < n) o[k] = i[k / 2];
}
int main() {
    const size_t IN = 3840ull * 2160 * 3 / 2, OUT = 3840ull * 2160 * 3;
    unsigned char *hi, *ho, *di, *dout;
    cudaHostAlloc(&hi, IN, 0); cudaHostAlloc(&ho, OUT, 0); cudaMalloc(&di, IN); cudaMalloc(&dout, OUT);
    cudaStream_t s0, s1, s2; cudaStreamCreateWithFlags(&s0, cudaStreamNonBlocking);
    cudaStreamCreateWithFlags(&s1, cudaStreamNonBlocking); cudaStreamCreateWithFlags(&s2, cudaStreamNonBlocking);
    cudaEvent_t e_in[32], e_k[32];
    for (int i = 0; i < 32; i++) { cudaEventCreateWithFlags(&e_in[i], cudaEventDisableTiming); cudaEventCreateWithFlags(&e_k[i], cudaEventDisableTiming); }
    for (int bands : {1, 2, 4, 8, 16}) {
        for (int rep = 0; rep < 3; rep++) {
            auto t0 = std::chrono::steady_clock::now();
            const int N = 20;
            for (int it = 0; it < N; it++) {
                for (int b = 0; b < bands; b++) {
                    size_t i0 = IN * b / bands, i1 = IN * (b + 1) / bands, o0 = OUT * b / bands, o1 = OUT * (b + 1) / bands;
                    cudaMemcpyAsync(di + i0, hi + i0, i1 - i0, cudaMemcpyHostToDevice, s0);
                    cudaEventRecord(e_in[b], s0); cudaStreamWaitEvent(s1, e_in[b], 0);
                    touch<<<(unsigned)((o1 - o0 + 255) / 256), 256, 0, s1>>>(dout + o0, di + i0, o1 - o0);
                    cudaEventRecord(e_k[b], s1); cudaStreamWaitEvent(s2, e_k[b], 0);
                    cudaMemcpyAsync(ho + o0, dout + o0, o1 - o0, cudaMemcpyDeviceToHost, s2);
                }
                cudaStreamSynchronize(s2);
            }
            double ms = std::chrono::duration<double, std::milli>(std::chrono::steady_clock::now() - t0).count() / N;
            if (rep == 2) printf("bands %2d: %.3f ms/frame  (%.1f Gpix/s)\n", bands, ms, 3840.0 * 2160 / ms / 1e6);
        }
    }
    return 0;
}
