// Microbenchmark: issue rate of packed FP32 (FFMA2 / FMUL2 / FADD2) vs scalar FFMA on sm_100a, alone and mixed with
// ALU-pipe work (FMNMX). Build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o ffma2 ffma2.cu ; prints warp
// instructions per cycle per SM sub-partition.
#include <cstdio>
#include <cuda_runtime.h>

typedef unsigned long long u64;
__device__ __forceinline__ u64 fma2(u64 a, u64 b, u64 c) { u64 r; asm volatile("fma.rn.f32x2 %0, %1, %2, %3;" : "=l"(r) : "l"(a), "l"(b), "l"(c)); return r; }
__device__ __forceinline__ u64 mul2(u64 a, u64 b) { u64 r; asm volatile("mul.rn.f32x2 %0, %1, %2;" : "=l"(r) : "l"(a), "l"(b)); return r; }
__device__ __forceinline__ u64 add2(u64 a, u64 b) { u64 r; asm volatile("add.rn.f32x2 %0, %1, %2;" : "=l"(r) : "l"(a), "l"(b)); return r; }
__device__ __forceinline__ float fma1(float a, float b, float c) { float r; asm volatile("fma.rn.f32 %0, %1, %2, %3;" : "=f"(r) : "f"(a), "f"(b), "f"(c)); return r; }
__device__ __forceinline__ float mnmx(float a, float b) { float r; asm volatile("max.f32 %0, %1, %2;" : "=f"(r) : "f"(a), "f"(b)); return r; }

template <int MODE>
__global__ void bench(float* out, long long* cycles, int iters, float seed) {
    float s[8];
    u64 p[8];
    float m[8];
    for (int i = 0; i < 8; i++) {
        s[i] = seed + i + threadIdx.x;
        float2 v = make_float2(seed + i, seed - i + threadIdx.x);
        p[i] = *reinterpret_cast<u64*>(&v);
        m[i] = seed * i;
    }
    const float b = 1.0000001f, c = 1e-9f;
    float2 bb = make_float2(b, b), cc = make_float2(c, c);
    const u64 b2 = *reinterpret_cast<u64*>(&bb), c2 = *reinterpret_cast<u64*>(&cc);
    __syncthreads();
    const long long t0 = clock64();
    for (int it = 0; it < iters; it++) {
#pragma unroll
        for (int i = 0; i < 8; i++) {
            if (MODE == 0) s[i] = fma1(s[i], b, c);                       // 8 FFMA
            if (MODE == 1) p[i] = fma2(p[i], b2, c2);                     // 8 FFMA2
            if (MODE == 2) { p[i] = fma2(p[i], b2, c2); m[i] = mnmx(m[i], s[i]); }   // 8 FFMA2 + 8 FMNMX
            if (MODE == 3) { s[i] = fma1(s[i], b, c); m[i] = mnmx(m[i], s[i]); }      // 8 FFMA + 8 FMNMX
            if (MODE == 4) p[i] = mul2(p[i], b2);                         // 8 FMUL2
            if (MODE == 5) p[i] = add2(p[i], c2);                         // 8 FADD2
            if (MODE == 6) { p[i] = fma2(p[i], b2, c2); s[i] = fma1(s[i], b, c); }    // 8 FFMA2 + 8 FFMA
        }
    }
    const long long t1 = clock64();
    float acc = 0;
    for (int i = 0; i < 8; i++) { float2 v = *reinterpret_cast<float2*>(&p[i]); acc += s[i] + v.x + v.y + m[i]; }
    out[blockIdx.x * blockDim.x + threadIdx.x] = acc;
    if (threadIdx.x == 0) cycles[blockIdx.x] = t1 - t0;
}

template <int MODE>
void run(const char* name, int instr_per_iter) {
    const int blocks = 148 * 2, threads = 512, iters = 20000;   // 32 warps per SM = 8 per sub-partition
    float* out; long long* cyc;
    cudaMalloc(&out, blocks * threads * 4); cudaMalloc(&cyc, blocks * 8);
    bench<MODE><<<blocks, threads>>>(out, cyc, 100, 1.0f);
    bench<MODE><<<blocks, threads>>>(out, cyc, iters, 1.0f);
    cudaDeviceSynchronize();
    long long h[148 * 2]; cudaMemcpy(h, cyc, sizeof h, cudaMemcpyDeviceToHost);
    double avg = 0; for (int i = 0; i < blocks; i++) avg += h[i]; avg /= blocks;
    // per SM sub-partition: 8 warps x iters x instr_per_iter warp instructions in `avg` cycles
    printf("%-28s %.3f warp-instr / cycle / sub-partition  (%.0f cycles)\n", name, 8.0 * iters * instr_per_iter / avg, avg);
    cudaFree(out); cudaFree(cyc);
}

int main() {
    run<0>("FFMA", 8);
    run<1>("FFMA2", 8);
    run<4>("FMUL2", 8);
    run<5>("FADD2", 8);
    run<2>("FFMA2 + FMNMX", 16);
    run<3>("FFMA + FMNMX", 16);
    run<6>("FFMA2 + FFMA", 16);
    return 0;
}
