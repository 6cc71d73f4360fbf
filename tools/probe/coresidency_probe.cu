// Do CTAs of a small kernel join SMs that hold one persistent CTA with ~200 KB of shared memory (and vice versa)?
// nvcc -O3 -gencode arch=compute_100a,code=sm_100a -o coresidency_probe coresidency_probe.cu && ./coresidency_probe
#include <cstdio>
#include <cuda_runtime.h>

__global__ void big_kernel(long long cycles, int* sink) {  // 1 CTA per SM, big dynamic shared memory
  extern __shared__ int sm[];
  const long long t0 = clock64();
  int acc = 0;
  while (clock64() - t0 < cycles) acc += sm[(threadIdx.x * 33 + acc) & 1023];
  if (acc == 0x7fffffff) sink[0] = acc;
}
template <int REGS_HINT>
__global__ void small_kernel(long long cycles, int* sink) {
  const long long t0 = clock64();
  int acc = threadIdx.x;
  while (clock64() - t0 < cycles) acc = acc * 1664525 + 1013904223;
  if (acc == 0x7fffffff) sink[1] = acc;
}

int main() {
  int sms = 0;
  cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, 0);
  int* sink;
  cudaMalloc(&sink, 64);
  cudaStream_t s1, s2;
  cudaStreamCreateWithFlags(&s1, cudaStreamNonBlocking);
  cudaStreamCreateWithFlags(&s2, cudaStreamNonBlocking);
  cudaEvent_t a0, a1, b0, b1;
  cudaEventCreate(&a0); cudaEventCreate(&a1); cudaEventCreate(&b0); cudaEventCreate(&b1);
  const long long big_cycles = 2000000, small_cycles = 600000;  // ~1 ms, ~0.3 ms
  for (int hint = 0; hint < 2; ++hint)
  for (int smem_kb : {200, 220}) {
    for (int small_smem : {0, 1024}) {
      if (hint) cudaFuncSetAttribute(small_kernel<0>, cudaFuncAttributePreferredSharedMemoryCarveout, (int)cudaSharedmemCarveoutMaxShared);
      for (int order = 0; order < 2; ++order) {
        const int smem = smem_kb * 1024;
        cudaFuncSetAttribute(big_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
        cudaDeviceSynchronize();
        float ta = 0, tb = 0, tot = 0;
        cudaEventRecord(a0, s1);
        cudaStreamWaitEvent(s2, a0, 0);
        cudaEventRecord(b0, s2);
        if (order == 0) {
          big_kernel<<<sms, 256, smem, s1>>>(big_cycles, sink);
          small_kernel<0><<<300, 32, small_smem, s2>>>(small_cycles, sink);
        } else {
          small_kernel<0><<<300, 32, small_smem, s2>>>(small_cycles, sink);
          big_kernel<<<sms, 256, smem, s1>>>(big_cycles, sink);
        }
        cudaEventRecord(a1, s1);
        cudaEventRecord(b1, s2);
        cudaDeviceSynchronize();
        cudaEventElapsedTime(&ta, a0, a1);
        cudaEventElapsedTime(&tb, b0, b1);
        cudaEventElapsedTime(&tot, a0, b1);
        printf("hint %d big %3d KB, small smem %5d B, %s first: big %.3f ms, small %.3f ms (isolated: ~%.2f / ~%.2f ms) err=%s\n", hint, smem_kb, small_smem,
               order == 0 ? "big" : "small", ta, tb, big_cycles / 1.9e6, small_cycles / 1.9e6, cudaGetErrorString(cudaGetLastError()));
      }
    }
  }
  return 0;
}
