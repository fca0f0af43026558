// PCIe probe for the e2e leg (not part of the product): 1D vs pitched H2D copies at the bench's geometries.
#include <cuda_runtime.h>
#include <chrono>
#include <cstdio>
#include <cstdlib>
static double now() { return std::chrono::duration<double>(std::chrono::steady_clock::now().time_since_epoch()).count(); }
int main() {
    const size_t rows = 1024;
    for (int u8 = 0; u8 < 2; u8++) {
        const size_t width = 196608 * (u8 ? 2 : 8), dpitch = 262144 * (u8 ? 2 : 8);
        unsigned char *h, *d, *hb, *db;
        cudaMallocHost(&h, rows * width);
        cudaMalloc(&d, rows * dpitch);
        cudaMallocHost(&hb, 1024 * 230400);
        cudaMalloc(&db, 1024 * 230400);
        cudaStream_t s1, s2;
        cudaStreamCreate(&s1); cudaStreamCreate(&s2);
        for (int variant = 0; variant < 5; variant++) {
            double best = 1e9;
            for (int rep = 0; rep < 6; rep++) {
                cudaDeviceSynchronize();
                double t0 = now();
                if (variant == 0) cudaMemcpyAsync(d, h, rows * width, cudaMemcpyHostToDevice, s1);
                if (variant == 1) cudaMemcpy2DAsync(d, dpitch, h, width, width, rows, cudaMemcpyHostToDevice, s1);
                if (variant == 2) for (int w = 0; w < 4; w++) cudaMemcpy2DAsync(d + w * 256 * dpitch, dpitch, h + w * 256 * width, width, width, 256, cudaMemcpyHostToDevice, s1);
                if (variant == 3) for (size_t r = 0; r < rows; r++) cudaMemcpyAsync(d + r * dpitch, h + r * width, width, cudaMemcpyHostToDevice, s1);
                if (variant == 4) {
                    for (int w = 0; w < 4; w++) {
                        cudaMemcpy2DAsync(d + w * 256 * dpitch, dpitch, h + w * 256 * width, width, width, 256, cudaMemcpyHostToDevice, s1);
                        cudaMemcpyAsync(hb + w * 256 * 230400, db + w * 256 * 230400, 256 * 230400, cudaMemcpyDeviceToHost, s2);
                    }
                }
                cudaDeviceSynchronize();
                double dt = now() - t0;
                if (dt < best) best = dt;
            }
            const char* names[] = {"1D whole", "2D whole", "2D x4 ways", "1D per row", "2D x4 ways + D2H 236MB concurrent"};
            printf("%s %-36s %.2f ms  %.1f GB/s (H2D bytes)\n", u8 ? "u8 " : "c64", names[variant], best * 1e3, rows * width / best / 1e9);
        }
        cudaFreeHost(h); cudaFree(d); cudaFreeHost(hb); cudaFree(db);
    }
    return 0;
}
