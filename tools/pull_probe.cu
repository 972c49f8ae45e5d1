// Development aid: how fast do the sampled rows of a pinned depth batch reach the device -- strided copy-engine copy against a kernel
// that reads the mapped host rows itself?   nvcc -arch=sm_100a -O3 -o tools/pull_probe tools/pull_probe.cu && ./tools/pull_probe
#include <cstdio>
#include <cuda_runtime.h>
#define CK(x) do { cudaError_t e = (x); if (e != cudaSuccess) { printf("%s: %s\n", #x, cudaGetErrorString(e)); return 1; } } while (0)

__global__ void k_pull(const uint4 *__restrict__ src, uint4 *__restrict__ dst, size_t row_u4, size_t src_pitch_u4, size_t dst_pitch_u4, size_t n_rows) {
    const size_t per = row_u4;                       // uint4 per row
    const size_t total = per * n_rows;
    for (size_t i = blockIdx.x * size_t(blockDim.x) + threadIdx.x; i < total; i += size_t(gridDim.x) * blockDim.x) {
        const size_t r = i / per, c = i - r * per;
        dst[r * dst_pitch_u4 + c] = __ldcs(src + r * src_pitch_u4 + c);
    }
}

int main() {
    const int frames = 250, rows = 480, cols = 640, dis = 3, h = 160;
    const size_t pitch = size_t(cols) * 4, fsz = pitch * rows;
    float *host = nullptr, *dev = nullptr;
    CK(cudaHostAlloc(&host, fsz * frames, cudaHostAllocMapped));
    for (size_t i = 0; i < fsz * frames / 4; i += 1024) host[i] = float(i);
    CK(cudaMalloc(&dev, fsz * frames));
    cudaStream_t st; CK(cudaStreamCreate(&st));
    cudaEvent_t e0, e1; CK(cudaEventCreate(&e0)); CK(cudaEventCreate(&e1));
    const size_t n_rows = size_t(h) * frames;        // rows % dis == 0: the sampled rows of consecutive frames are equally spaced
    const double mb = double(n_rows) * pitch / 1e6;
    float ms;
    for (int rep = 0; rep < 3; ++rep) {
        CK(cudaEventRecord(e0, st));
        CK(cudaMemcpy2DAsync(dev, pitch * dis, host, pitch * dis, pitch, n_rows, cudaMemcpyHostToDevice, st));
        CK(cudaEventRecord(e1, st)); CK(cudaEventSynchronize(e1)); CK(cudaEventElapsedTime(&ms, e0, e1));
        printf("strided copy engine : %.1f MB in %.3f ms = %.1f GB/s\n", mb, ms, mb / ms);
    }
    for (int rep = 0; rep < 3; ++rep) {
        CK(cudaEventRecord(e0, st));
        CK(cudaMemcpyAsync(dev, host, size_t(mb * 1e6), cudaMemcpyHostToDevice, st));
        CK(cudaEventRecord(e1, st)); CK(cudaEventSynchronize(e1)); CK(cudaEventElapsedTime(&ms, e0, e1));
        printf("contiguous copy     : %.1f MB in %.3f ms = %.1f GB/s\n", mb, ms, mb / ms);
    }
    const uint4 *hs; CK(cudaHostGetDevicePointer((void **)&hs, host, 0));
    for (int grid : {74, 148, 296, 592, 1184}) for (int block : {256, 512}) {
        for (int rep = 0; rep < 2; ++rep) {
            CK(cudaEventRecord(e0, st));
            k_pull<<<grid, block, 0, st>>>(hs, reinterpret_cast<uint4 *>(dev), pitch / 16, pitch * dis / 16, pitch * dis / 16, n_rows);
            CK(cudaEventRecord(e1, st)); CK(cudaEventSynchronize(e1)); CK(cudaEventElapsedTime(&ms, e0, e1));
        }
        printf("pull kernel %4d x %3d: %.3f ms = %.1f GB/s\n", grid, block, ms, mb / ms);
    }
    CK(cudaGetLastError());
    return 0;
}
