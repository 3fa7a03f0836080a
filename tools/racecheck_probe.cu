// Two minimal, textbook-correct programs for `compute-sanitizer --tool racecheck`:
//   probe_mbarrier: warp 0 writes shared memory and arrives on an mbarrier (release), warp 1 waits on the phase
//                   (acquire) and reads -- the hand-over the role-split convolution epilogue uses between its warps;
//   probe_alloc2  : a CTA pair that does nothing but tcgen05.alloc.cta_group::2 / read the slot after a cluster barrier
//                   / dealloc -- the collective TMEM allocation of the CTA-pair convolution.
// If racecheck reports hazards on these, its reports of the same form on the product kernels are tool artefacts.
// build: nvcc -gencode arch=compute_100a,code=sm_100a -o racecheck_probe tools/racecheck_probe.cu
#include <cstdint>
#include <cstdio>
#include <cuda_runtime.h>

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return static_cast<uint32_t>(__cvta_generic_to_shared(p)); }

__global__ void probe_mbarrier(float* out) {
  __shared__ float buf[32];
  __shared__ uint64_t bar;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  if (threadIdx.x == 0) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(smem_u32(&bar)));
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  __syncthreads();
  if (warp == 0) {
    buf[lane] = 1.0f + lane;
    __syncwarp();
    if (lane == 0) asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(&bar)) : "memory");
  } else {
    uint32_t ok = 0;
    while (!ok)
      asm volatile("{\n\t.reg .pred P;\n\tmbarrier.try_wait.parity.shared::cta.b64 P, [%1], 0;\n\tselp.u32 %0, 1, 0, P;\n\t}\n"
                   : "=r"(ok) : "r"(smem_u32(&bar)) : "memory");
    out[lane] = buf[lane];
  }
}

__global__ void __cluster_dims__(2, 1, 1) probe_alloc2(uint32_t* out) {
  // the product kernel's layout: the TMEM slot sits right behind the mbarriers another warp initialises meanwhile
  __shared__ __align__(16) struct { uint64_t bars[4]; uint32_t slot_; uint32_t pad; } blk;
  uint32_t& slot = blk.slot_;
  const int warp = threadIdx.x >> 5;
  if (warp == 1 && (threadIdx.x & 31) == 0) {
    for (int i = 0; i < 4; ++i) asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(smem_u32(&blk.bars[i])));
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  if (warp == 0) {
    asm volatile("tcgen05.alloc.cta_group::2.sync.aligned.shared::cta.b32 [%0], 128;" ::"r"(smem_u32(&slot)) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::2.sync.aligned;" ::: "memory");
  }
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  asm volatile("barrier.cluster.arrive.release.aligned;" ::: "memory");
  asm volatile("barrier.cluster.wait.acquire.aligned;" ::: "memory");
  asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
  const uint32_t base = slot;
  if (threadIdx.x == 0) out[blockIdx.x] = base;
  asm volatile("barrier.cluster.arrive.release.aligned;" ::: "memory");
  asm volatile("barrier.cluster.wait.acquire.aligned;" ::: "memory");
  if (warp == 0) asm volatile("tcgen05.dealloc.cta_group::2.sync.aligned.b32 %0, 128;" ::"r"(base) : "memory");
}

int main() {
  float* o;
  uint32_t* u;
  cudaMalloc(&o, 32 * sizeof(float));
  cudaMalloc(&u, 2 * sizeof(uint32_t));
  probe_mbarrier<<<1, 64>>>(o);
  probe_alloc2<<<2, 128>>>(u);
  cudaError_t e = cudaDeviceSynchronize();
  float h[32];
  cudaMemcpy(h, o, sizeof h, cudaMemcpyDeviceToHost);
  printf("probe: %s, mbarrier hand-over read %g .. %g\n", cudaGetErrorString(e), h[0], h[31]);
  return e != cudaSuccess;
}
