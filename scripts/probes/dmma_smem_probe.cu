// Is the DMMA loop of the big trailing update limited by shared-memory operand delivery?
// Same fragment loads and mma.sync sequence as gemm_big_kernel, operands resident in shared memory, no global traffic.
// nvcc -arch=sm_100a -O3 -o /tmp/p scripts/probes/dmma_smem_probe.cu && /tmp/p
#include <cstdio>
#include <cuda_runtime.h>
constexpr int BM = 128, BN = 128, BK = 32, LDA = BM + 4, LDB = BK + 4;
__device__ __forceinline__ void dmma(double& c0, double& c1, double a, double b) {
    asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};" : "+d"(c0), "+d"(c1) : "d"(a), "d"(b));
}
template <int TI, int WM_COUNT, bool SYNC>
__global__ void __launch_bounds__(WM_COUNT * 4 * 32, 1) loop(double* out, int chunks) {
    extern __shared__ double sm[];
    double* As = sm; double* Bs = sm + BK * LDA;
    for (int e = threadIdx.x; e < BK * LDA + BN * LDB; e += blockDim.x) sm[e] = 1e-3 * (e % 7);
    __syncthreads();
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const int wm = (warp % WM_COUNT) * (8 * TI), wn = (warp / WM_COUNT) * 32;
    const int r = lane >> 2, q = lane & 3;
    double acc[TI][4][2];
    for (int i = 0; i < TI; ++i) for (int j = 0; j < 4; ++j) acc[i][j][0] = acc[i][j][1] = 0.0;
    for (int c = 0; c < chunks; ++c) {
        if (SYNC) __syncthreads();
#pragma unroll
        for (int ks = 0; ks < BK; ks += 4) {
            double a[TI], b[4];
#pragma unroll
            for (int i = 0; i < TI; ++i) a[i] = As[(ks + q) * LDA + wm + 8 * i + r];
#pragma unroll
            for (int j = 0; j < 4; ++j) b[j] = Bs[(wn + 8 * j + r) * LDB + ks + q];
#pragma unroll
            for (int i = 0; i < TI; ++i)
#pragma unroll
                for (int j = 0; j < 4; ++j) dmma(acc[i][j][0], acc[i][j][1], a[i], b[j]);
        }
    }
    double s = 0;
    for (int i = 0; i < TI; ++i) for (int j = 0; j < 4; ++j) s += acc[i][j][0] + acc[i][j][1];
    if (s == 123.456) out[0] = s;
}
template <int TI, int WM_COUNT, bool SYNC>
void run(const char* name, double* d) {
    const int threads = WM_COUNT * 4 * 32, chunks = 4000;
    const size_t smem = (BK * LDA + BN * LDB) * sizeof(double);
    cudaFuncSetAttribute(loop<TI, WM_COUNT, SYNC>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
    loop<TI, WM_COUNT, SYNC><<<148, threads, smem>>>(d, 100);
    cudaEventRecord(e0);
    loop<TI, WM_COUNT, SYNC><<<148, threads, smem>>>(d, chunks);
    cudaEventRecord(e1); cudaEventSynchronize(e1);
    float ms; cudaEventElapsedTime(&ms, e0, e1);
    printf("%-40s %6.2f TF\n", name, 2.0 * BM * BN * BK * chunks * 148 / ms / 1e9);
}
int main() {
    double* d; cudaMalloc(&d, 8);
    run<8, 2, false>("8 warps 64x32, no barrier", d);
    run<8, 2, true>("8 warps 64x32, barrier per chunk", d);
    run<4, 4, false>("16 warps 32x32, no barrier", d);
    run<4, 4, true>("16 warps 32x32, barrier per chunk", d);
    printf("%s\n", cudaGetErrorString(cudaGetLastError()));
    return 0;
}
