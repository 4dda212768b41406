// Microbenchmark: how fast can a kernel store a 1920x1080 frame of 16-byte records into page-locked host memory, as a
// function of the contiguous bytes one warp store instruction covers? (rt_primary with a pinned destination stores one
// 8x4-pixel tile per warp: four 128-byte row pieces per instruction.)
// Build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o host_store host_store.cu ; prints GB/s per pattern.
#include <cstdio>
#include <cstdlib>
#include <cuda_runtime.h>

#define CK(x) do { cudaError_t e = (x); if (e != cudaSuccess) { printf("%s: %s\n", #x, cudaGetErrorString(e)); exit(1); } } while (0)

// SEG = pixels per contiguous run of one warp store: 8 (8x4 tile), 16 (16x2), 32 (32x1); 0 = linear (warp i -> bytes [512 i, +512))
template <int SEG>
__global__ void store_kernel(float4* out, int w, int h, unsigned long long* counter) {
    const int lane = threadIdx.x & 31;
    const int rows = SEG ? 32 / SEG : 1;
    const int tiles_x = SEG ? (w + SEG - 1) / SEG : 0, tiles_y = SEG ? (h + rows - 1) / rows : 0;
    const long long items = SEG ? (long long)tiles_x * tiles_y : ((long long)w * h + 31) / 32;
    for (;;) {
        unsigned long long it = 0;
        if (lane == 0) it = atomicAdd(counter, 1ull);
        it = __shfl_sync(0xffffffffu, it, 0);
        if ((long long)it >= items) break;
        long long idx;
        if (SEG) {
            const int ty = (int)(it / tiles_x), tx = (int)(it - (long long)ty * tiles_x);
            const int x = tx * SEG + (lane % SEG), y = ty * rows + lane / SEG;
            if (x >= w || y >= h) continue;
            idx = (long long)y * w + x;
        } else {
            idx = (long long)it * 32 + lane;
            if (idx >= (long long)w * h) continue;
        }
        out[idx] = make_float4(__int_as_float(-1), 4294967296.0f, (float)idx, 0.0f);
    }
}

template <int SEG>
static float run(float4* dst, int w, int h, unsigned long long* counter, int iters) {
    cudaEvent_t a, b;
    CK(cudaEventCreate(&a));
    CK(cudaEventCreate(&b));
    float best = 1e30f, sum = 0;
    for (int i = 0; i < iters + 3; i++) {
        CK(cudaMemsetAsync(counter, 0, 8));
        CK(cudaEventRecord(a));
        store_kernel<SEG><<<148 * 8, 128>>>(dst, w, h, counter);
        CK(cudaEventRecord(b));
        CK(cudaEventSynchronize(b));
        float ms;
        CK(cudaEventElapsedTime(&ms, a, b));
        if (i >= 3) { sum += ms; if (ms < best) best = ms; }
    }
    return sum / iters;
}

int main() {
    const int w = 1920, h = 1080;
    const size_t bytes = (size_t)w * h * 16;
    float4 *host, *dev_alias, *dev;
    unsigned long long* counter;
    CK(cudaHostAlloc(&host, bytes, cudaHostAllocMapped));
    CK(cudaHostGetDevicePointer(&dev_alias, host, 0));
    CK(cudaMalloc(&dev, bytes));
    CK(cudaMalloc(&counter, 8));
    const double gb = bytes / 1e9;
    cudaEvent_t a, b;
    CK(cudaEventCreate(&a));
    CK(cudaEventCreate(&b));
    float sum = 0;
    for (int i = 0; i < 13; i++) {
        CK(cudaEventRecord(a));
        CK(cudaMemcpyAsync(host, dev, bytes, cudaMemcpyDeviceToHost));
        CK(cudaEventRecord(b));
        CK(cudaEventSynchronize(b));
        float ms;
        CK(cudaEventElapsedTime(&ms, a, b));
        if (i >= 3) sum += ms;
    }
    printf("{\"bytes\": %zu, \"cudaMemcpy_D2H_GBps\": %.2f", bytes, gb / (sum / 10 * 1e-3));
    for (int rep = 0; rep < 2; rep++) {
        const float t8 = run<8>(dev_alias, w, h, counter, 10), t16 = run<16>(dev_alias, w, h, counter, 10),
                    t32 = run<32>(dev_alias, w, h, counter, 10), t0 = run<0>(dev_alias, w, h, counter, 10);
        printf(", \"round%d\": {\"tile_8x4_128B_GBps\": %.2f, \"tile_16x2_256B_GBps\": %.2f, \"tile_32x1_512B_GBps\": %.2f, \"linear_512B_GBps\": %.2f}", rep,
               gb / (t8 * 1e-3), gb / (t16 * 1e-3), gb / (t32 * 1e-3), gb / (t0 * 1e-3));
    }
    const float tl = run<8>(dev, w, h, counter, 10);
    printf(", \"tile_8x4_into_device_memory_ms\": %.4f}\n", tl);
    return 0;
}
