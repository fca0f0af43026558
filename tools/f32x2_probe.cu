// Micro-probe: issue throughput of FADD/FFMA/FMUL vs their packed f32x2 forms (FADD2/FFMA2/FMUL2) on sm_100a.
// Build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o f32x2_probe f32x2_probe.cu ; run on the GPU box.
#include <cstdio>
#include <cuda_runtime.h>
typedef unsigned long long u64;
#define CHAINS 8
template <int MODE>
__global__ void __launch_bounds__(256) probe(float* out, int iters, float seed) {
    float a[2 * CHAINS];
    for (int i = 0; i < 2 * CHAINS; i++) a[i] = seed + threadIdx.x * 1e-3f + i;
    const float b = seed * 0.5f, c = seed * 0.25f;
    u64 bb, cc;
    asm("mov.b64 %0, {%1, %2};" : "=l"(bb) : "f"(b), "f"(b));
    asm("mov.b64 %0, {%1, %2};" : "=l"(cc) : "f"(c), "f"(c));
    if (MODE == 0 || MODE == 2 || MODE == 4) {
        for (int it = 0; it < iters; it++) {
#pragma unroll
            for (int u = 0; u < 8; u++) {
#pragma unroll
                for (int i = 0; i < 2 * CHAINS; i++) {
                    if (MODE == 0) asm volatile("add.rn.f32 %0, %0, %1;" : "+f"(a[i]) : "f"(b));
                    if (MODE == 2) asm volatile("fma.rn.f32 %0, %0, %1, %2;" : "+f"(a[i]) : "f"(b), "f"(c));
                    if (MODE == 4) asm volatile("mul.rn.f32 %0, %0, %1;" : "+f"(a[i]) : "f"(b));
                }
            }
        }
    } else {
        u64 p[CHAINS];
        for (int i = 0; i < CHAINS; i++) asm("mov.b64 %0, {%1, %2};" : "=l"(p[i]) : "f"(a[2 * i]), "f"(a[2 * i + 1]));
        for (int it = 0; it < iters; it++) {
#pragma unroll
            for (int u = 0; u < 8; u++) {
#pragma unroll
                for (int i = 0; i < CHAINS; i++) {
                    if (MODE == 1) asm volatile("add.rn.f32x2 %0, %0, %1;" : "+l"(p[i]) : "l"(bb));
                    if (MODE == 3) asm volatile("fma.rn.f32x2 %0, %0, %1, %2;" : "+l"(p[i]) : "l"(bb), "l"(cc));
                    if (MODE == 5) asm volatile("mul.rn.f32x2 %0, %0, %1;" : "+l"(p[i]) : "l"(bb));
                }
            }
        }
        for (int i = 0; i < CHAINS; i++) asm("mov.b64 {%0, %1}, %2;" : "=f"(a[2 * i]), "=f"(a[2 * i + 1]) : "l"(p[i]));
    }
    float s = 0;
    for (int i = 0; i < 2 * CHAINS; i++) s += a[i];
    out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}
template <int MODE>
void run(const char* name, float* d) {
    const int iters = 4096, blocks = 148 * 8;
    cudaEvent_t e0, e1;
    cudaEventCreate(&e0); cudaEventCreate(&e1);
    probe<MODE><<<blocks, 256>>>(d, 16, 1.0f);
    cudaEventRecord(e0);
    probe<MODE><<<blocks, 256>>>(d, iters, 1.0f);
    cudaEventRecord(e1);
    cudaEventSynchronize(e1);
    float ms; cudaEventElapsedTime(&ms, e0, e1);
    const double flops_elem = double(blocks) * 256 * iters * 8 * 2 * CHAINS;  // scalar-element operations
    printf("%-8s %8.3f ms  %8.2f T elem-op/s  (%.1f elem-op/clk/SM at 1.965 GHz)\n", name, ms, flops_elem / ms * 1e-9, flops_elem / (ms * 1e-3) / 148 / 1.965e9);
}
int main() {
    float* d; cudaMalloc(&d, 148 * 8 * 256 * 4);
    run<0>("FADD", d); run<1>("FADD2", d); run<2>("FFMA", d); run<3>("FFMA2", d); run<4>("FMUL", d); run<5>("FMUL2", d);
    return 0;
}
