// How many warps / how much ILP does mma.sync.m8n8k4.f64 need to reach its peak on sm_100a?
// nvcc -arch=sm_100a -O3 -o /tmp/dmma_probe scripts/probes/dmma_probe.cu && /tmp/dmma_probe
#include <cstdio>
#include <cuda_runtime.h>

template <int ILP>
__global__ void __launch_bounds__(1024) chain(double* out, int iters, double a0, double b0) {
    double c[ILP][2];
    double a[4], b[4];
#pragma unroll
    for (int i = 0; i < 4; ++i) { a[i] = a0 + i * 1e-9 * threadIdx.x; b[i] = b0 - i * 1e-9; }
#pragma unroll
    for (int i = 0; i < ILP; ++i) { c[i][0] = threadIdx.x * 1e-3 + i; c[i][1] = c[i][0] + 0.5; }
    for (int it = 0; it < iters; ++it) {
#pragma unroll
        for (int i = 0; i < ILP; ++i)
            asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};"
                         : "+d"(c[i][0]), "+d"(c[i][1]) : "d"(a[i & 3]), "d"(b[(i >> 2) & 3]));
    }
    double s = 0;
#pragma unroll
    for (int i = 0; i < ILP; ++i) s += c[i][0] + c[i][1];
    if (s == 123.456) out[0] = s;
}

template <int ILP>
void run(int warps, int ctas_per_sm, double* d) {
    int sms = 148;
    int iters = 20000 / ILP * 8;
    cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
    chain<ILP><<<sms * ctas_per_sm, warps * 32>>>(d, iters / 10, 1e-3, 1e-3);
    cudaEventRecord(e0);
    chain<ILP><<<sms * ctas_per_sm, warps * 32>>>(d, iters, 1e-3, 1e-3);
    cudaEventRecord(e1); cudaEventSynchronize(e1);
    float ms; cudaEventElapsedTime(&ms, e0, e1);
    double fl = 512.0 * ILP * iters * (double)warps * sms * ctas_per_sm;
    printf("ILP %2d warps/CTA %2d CTAs/SM %d -> %6.2f TF  (%.1f clk per DMMA per SM at 1.965 GHz)\n", ILP, warps, ctas_per_sm,
           fl / ms / 1e9, ms * 1e-3 * 1.965e9 / ((double)ILP * iters * warps * ctas_per_sm));
}

int main() {
    double* d; cudaMalloc(&d, 8);
    for (int w : {4, 8, 16, 32}) { run<8>(w, 1, d); run<16>(w, 1, d); run<32>(w, 1, d); }
    run<8>(8, 2, d); run<8>(8, 4, d); run<8>(8, 8, d); run<32>(8, 2, d);
    return 0;
}
