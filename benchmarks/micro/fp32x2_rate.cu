// fp32x2_rate.cu — issue rate and dependent latency of the packed single-precision instructions of sm_100a
// (PTX fma.rn.f32x2 / add.rn.f32x2 / mul.rn.f32x2 -> SASS FFMA2 / FADD2 / FMUL2) beside scalar FFMA, and of FSETP.
// Question: does a packed instruction cost one issue slot for two operations, and how long does the pipe stay busy?
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -fmad=false -o fp32x2_rate fp32x2_rate.cu && ./fp32x2_rate
#include <cstdio>
#include <cuda_runtime.h>

template <int ILP, int OP>
__global__ void chain(float* out, long long* cycles, int iters, float a, float b) {
    float x[ILP];
    unsigned long long p[ILP];
    const float2 ab = make_float2(a, a), bb = make_float2(b, b);
    const unsigned long long a2 = *reinterpret_cast<const unsigned long long*>(&ab), b2 = *reinterpret_cast<const unsigned long long*>(&bb);
#pragma unroll
    for (int k = 0; k < ILP; ++k) {
        x[k] = (float)threadIdx.x + k;
        const float2 v = make_float2(x[k], x[k] + 0.5f);
        p[k] = *reinterpret_cast<const unsigned long long*>(&v);
    }
    int hits = 0;
    __syncthreads();
    long long t0 = clock64();
#pragma unroll 1
    for (int i = 0; i < iters; ++i) {
#pragma unroll
        for (int r = 0; r < 8; ++r) {
#pragma unroll
            for (int k = 0; k < ILP; ++k) {
                if (OP == 0) x[k] = fmaf(x[k], a, b);
                else if (OP == 1) asm volatile("fma.rn.f32x2 %0, %0, %1, %2;" : "+l"(p[k]) : "l"(a2), "l"(b2));
                else if (OP == 2) asm volatile("add.rn.f32x2 %0, %0, %1;" : "+l"(p[k]) : "l"(b2));
                else if (OP == 3) asm volatile("mul.rn.f32x2 %0, %0, %1;" : "+l"(p[k]) : "l"(a2));
                else if (OP == 4) {  // FFMA + FSETP pairs: does the compare share the FMA pipe?
                    x[k] = fmaf(x[k], a, b);
                    asm volatile("{ .reg .pred q; setp.gt.f32 q, %1, %2; @q add.s32 %0, %0, 1; }" : "+r"(hits) : "f"(x[k]), "f"(b));
                }
            }
        }
    }
    long long t1 = clock64();
    float s = (float)hits;
#pragma unroll
    for (int k = 0; k < ILP; ++k) s += x[k] + (float)(p[k] & 0xffff);
    if (s == -1.2345f) out[0] = s;
    if (threadIdx.x == 0) cycles[blockIdx.x] = t1 - t0;
}

template <int ILP, int OP>
void run(const char* name, int warps) {
    float* d; long long* c;
    cudaMalloc(&d, 8); cudaMalloc(&c, 8);
    const int iters = 4096;
    chain<ILP, OP><<<1, 32 * warps>>>(d, c, iters, 0.999f, 0.001f);
    chain<ILP, OP><<<1, 32 * warps>>>(d, c, iters, 0.999f, 0.001f);
    long long h; cudaMemcpy(&h, c, 8, cudaMemcpyDeviceToHost);
    const double ops = (double)iters * 8 * ILP;  // instructions per warp (OP 4: FFMA + FSETP + predicated IADD count as one)
    printf("%-10s warps/CTA %2d (per scheduler %d) ILP %d : %.2f cycles per instr per warp, %.3f warp-instr/cycle/scheduler\n", name, warps, (warps + 3) / 4, ILP,
           (double)h / ops, ops * ((warps + 3) / 4) / (double)h);
    cudaFree(d); cudaFree(c);
}

int main() {
    run<1, 0>("FFMA", 1); run<8, 0>("FFMA", 1); run<4, 0>("FFMA", 16); run<8, 0>("FFMA", 16);
    run<1, 1>("FFMA2", 1); run<8, 1>("FFMA2", 1); run<4, 1>("FFMA2", 16); run<8, 1>("FFMA2", 16);
    run<1, 2>("FADD2", 1); run<8, 2>("FADD2", 16);
    run<1, 3>("FMUL2", 1); run<8, 3>("FMUL2", 16);
    run<8, 4>("FFMA+FSETP", 16);
    return 0;
}
