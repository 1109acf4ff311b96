// Aggregate host<->device bandwidth of the box: how much page-locked frame traffic can N GPUs move at once?
// One process, one thread per device; every thread streams 4K yuv420p frames in (12.4 MB) and rgb24 frames out
// (24.9 MB) on two streams, full duplex, for about a second.  Run for N = 1, 2, 4, 8 concurrently.
//   nvcc -O2 -o pcie_multi pcie_multi.cu -lpthread ; ./pcie_multi
// The e2e (host-frame) throughput of the scaler cannot exceed  min(H2D / 12.4 MB, D2H / 24.9 MB)  frames/s.
#include <chrono>
#include <cstdio>
#include <cstdlib>
#include <thread>
#include <vector>
#include <atomic>
#include <cuda_runtime.h>

static const size_t IN = 3840ull * 2160 * 3 / 2, OUT = 3840ull * 2160 * 3;
static std::atomic<int> ready{0};
static std::atomic<bool> go{false};

struct Result { double h2d_gbs, d2h_gbs; };

static void worker(int dev, int ndev, double seconds, int mode, Result *res)
{
    cudaSetDevice(dev);
    unsigned char *hi, *ho, *di, *dout;
    cudaHostAlloc(&hi, IN, cudaHostAllocDefault); cudaHostAlloc(&ho, OUT, cudaHostAllocDefault);
    cudaMalloc(&di, IN); cudaMalloc(&dout, OUT);
    cudaStream_t s0, s1;
    cudaStreamCreateWithFlags(&s0, cudaStreamNonBlocking); cudaStreamCreateWithFlags(&s1, cudaStreamNonBlocking);
    cudaMemcpyAsync(di, hi, IN, cudaMemcpyHostToDevice, s0); cudaMemcpyAsync(ho, dout, OUT, cudaMemcpyDeviceToHost, s1);
    cudaDeviceSynchronize();
    ready++;
    while (!go.load()) std::this_thread::yield();
    auto t0 = std::chrono::steady_clock::now();
    long n_in = 0, n_out = 0;
    for (;;) {
        for (int k = 0; k < 8; k++) {
            if (mode != 2) { cudaMemcpyAsync(di, hi, IN, cudaMemcpyHostToDevice, s0); n_in++; }
            if (mode != 1) { cudaMemcpyAsync(ho, dout, OUT, cudaMemcpyDeviceToHost, s1); n_out++; }
        }
        cudaStreamSynchronize(s0); cudaStreamSynchronize(s1);
        double dt = std::chrono::duration<double>(std::chrono::steady_clock::now() - t0).count();
        if (dt > seconds) { res->h2d_gbs = n_in * (double)IN / dt / 1e9; res->d2h_gbs = n_out * (double)OUT / dt / 1e9; break; }
    }
    cudaFreeHost(hi); cudaFreeHost(ho); cudaFree(di); cudaFree(dout);
}

int main(int argc, char **argv)
{
    int nd = 0;
    cudaGetDeviceCount(&nd);
    const double seconds = argc > 1 ? atof(argv[1]) : 1.0;
    printf("devices: %d, host threads: %u\n", nd, std::thread::hardware_concurrency());
    const char *names[3] = { "duplex", "H2D only", "D2H only" };
    for (int mode = 0; mode < 3; mode++)
        for (int n = 1; n <= nd; n *= 2) {
            std::vector<Result> r(n);
            std::vector<std::thread> th;
            ready = 0; go = false;
            for (int d = 0; d < n; d++) th.emplace_back(worker, d, n, seconds, mode, &r[d]);
            while (ready.load() < n) std::this_thread::yield();
            go = true;
            for (auto &t : th) t.join();
            double hi = 0, ho = 0, lo_h = 1e9, lo_d = 1e9;
            for (auto &x : r) { hi += x.h2d_gbs; ho += x.d2h_gbs; if (x.h2d_gbs < lo_h) lo_h = x.h2d_gbs; if (x.d2h_gbs < lo_d) lo_d = x.d2h_gbs; }
            double cap = 1e18;
            if (mode != 2 && hi > 0) cap = hi * 1e9 / IN;
            if (mode != 1 && ho > 0 && ho * 1e9 / OUT < cap) cap = ho * 1e9 / OUT;
            printf("%-9s %d GPU%s: H2D %7.1f GB/s (slowest %5.1f)  D2H %7.1f GB/s (slowest %5.1f)  -> ceiling %6.1f Gpixel/s of 4K yuv420p->rgb24\n",
                   names[mode], n, n > 1 ? "s" : " ", hi, mode != 2 ? lo_h : 0.0, ho, mode != 1 ? lo_d : 0.0,
                   mode == 0 ? cap * 3840 * 2160 / 1e9 : 0.0);
        }
    return 0;
}
