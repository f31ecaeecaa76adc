// fp64_latency.cu — dependent-issue latency and per-scheduler throughput of the FP64 pipe on sm_100a.
// Answers: how many independent FP64 chains does one SM sub-partition need in flight to keep its FP64 pipe busy?
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -fmad=false -o fp64_latency fp64_latency.cu && ./fp64_latency
#include <cstdio>
#include <cuda_runtime.h>

template <int ILP, int OP>
__global__ void chain(double* out, long long* cycles, int iters, double a, double b) {
    double x[ILP];
#pragma unroll
    for (int k = 0; k < ILP; ++k) x[k] = (double)threadIdx.x + k;
    __syncthreads();
    long long t0 = clock64();
#pragma unroll 1
    for (int i = 0; i < iters; ++i) {
#pragma unroll
        for (int r = 0; r < 8; ++r) {
#pragma unroll
            for (int k = 0; k < ILP; ++k) {
                if (OP == 0) x[k] = fma(x[k], a, b);
                else if (OP == 1) x[k] = x[k] + b;
                else if (OP == 2) x[k] = x[k] * a;
                else { double s; asm volatile("rcp.approx.ftz.f64 %0, %1;" : "=d"(s) : "d"(x[k])); x[k] = s; }
            }
        }
    }
    long long t1 = clock64();
    double s = 0;
#pragma unroll
    for (int k = 0; k < ILP; ++k) s += x[k];
    if (s == -1.2345) out[0] = s;
    if (threadIdx.x == 0) cycles[blockIdx.x] = t1 - t0;
}

template <int ILP, int OP>
void run(const char* name, int warps) {
    double* d; long long* c;
    cudaMalloc(&d, 8); cudaMalloc(&c, 8);
    const int iters = 4096;
    chain<ILP, OP><<<1, 32 * warps>>>(d, c, iters, 0.999, 0.001);
    chain<ILP, OP><<<1, 32 * warps>>>(d, c, iters, 0.999, 0.001);
    long long h; cudaMemcpy(&h, c, 8, cudaMemcpyDeviceToHost);
    const double ops = (double)iters * 8 * ILP;  // per warp
    // warps are spread over the 4 sub-partitions: warps per scheduler = ceil(warps / 4)
    printf("%-6s warps/CTA %2d (per scheduler %d) ILP %d : %.2f cycles per op per warp, %.3f warp-ops/cycle/scheduler\n", name, warps, (warps + 3) / 4, ILP,
           (double)h / ops, ops * ((warps + 3) / 4) / (double)h);
    cudaFree(d); cudaFree(c);
}

int main() {
    run<1, 0>("DFMA", 1); run<2, 0>("DFMA", 1); run<4, 0>("DFMA", 1); run<8, 0>("DFMA", 1);
    run<1, 0>("DFMA", 4); run<1, 0>("DFMA", 8); run<1, 0>("DFMA", 16); run<2, 0>("DFMA", 16); run<1, 0>("DFMA", 32); run<4, 0>("DFMA", 16);
    run<1, 1>("DADD", 1); run<4, 1>("DADD", 1); run<1, 2>("DMUL", 1); run<4, 2>("DMUL", 1);
    run<1, 3>("RCP64", 1); run<4, 3>("RCP64", 1); run<1, 3>("RCP64", 16);
    return 0;
}
