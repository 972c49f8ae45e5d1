// kernel of tools/tma_probe3.cu, compiled to a cubin and loaded through the driver API (cuModuleLoadData)
#include <cuda.h>
extern "C" __global__ void k3(const __grid_constant__ CUtensorMap map, int c0, int c1, int bytes, float *out) {
    extern __shared__ __align__(1024) unsigned char dyn[];
    float *tile = reinterpret_cast<float *>(dyn);
    const unsigned bar = (unsigned)__cvta_generic_to_shared(dyn + 32768), dst = (unsigned)__cvta_generic_to_shared(dyn);
    if (threadIdx.x == 0) {
        asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(bar));
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    __syncthreads();
    if (threadIdx.x == 0) {
        asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"(bytes) : "memory");
        asm volatile("cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%2, %3}], [%4];"
                     ::"r"(dst), "l"(&map), "r"(c0), "r"(c1), "r"(bar) : "memory");
    }
    unsigned done = 0;
    while (!done) asm volatile("{ .reg .pred p; mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2; selp.u32 %0, 1, 0, p; }" : "=r"(done) : "r"(bar), "r"(0) : "memory");
    for (int i = threadIdx.x; i < 8192; i += blockDim.x) out[i] = tile[i];
}
