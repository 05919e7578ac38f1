// Write-only HBM bandwidth by store flavour (B200).  Build: nvcc -arch=sm_100a -O3 -o write_bw write_bw.cu
#include <cuda_runtime.h>
#include <cstdio>
#include <cstdint>
#include <cuda/ptx>
#define CK(x) do { cudaError_t e = (x); if (e != cudaSuccess) { printf("%s: %s\n", #x, cudaGetErrorString(e)); return 1; } } while (0)

template <int MODE>
__global__ void k_store(uint4* out, size_t n16) {
  const uint4 v = make_uint4(threadIdx.x, blockIdx.x, 3, 4);
  for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < n16; i += (size_t)gridDim.x * blockDim.x) {
    if (MODE == 0) out[i] = v;
    if (MODE == 1) __stcs(out + i, v);
    if (MODE == 2) __stwt(out + i, v);
    if (MODE == 3) __stcg(out + i, v);
  }
}
// each CTA writes contiguous chunks of `chunk16` uint4 (like one tile of observations)
template <int MODE>
__global__ void k_store_chunk(uint4* out, size_t n16, int chunk16) {
  const uint4 v = make_uint4(threadIdx.x, blockIdx.x, 3, 4);
  const size_t nchunks = n16 / chunk16;
  for (size_t c = blockIdx.x; c < nchunks; c += gridDim.x) {
    uint4* p = out + c * chunk16;
    for (int i = threadIdx.x; i < chunk16; i += blockDim.x) {
      if (MODE == 0) p[i] = v; else __stcs(p + i, v);
    }
  }
}
// TMA bulk store: every CTA fills shared memory once and streams it out repeatedly
__global__ void k_bulk(uint8_t* out, size_t nbytes, int chunk) {
  extern __shared__ __align__(128) uint8_t sm[];
  for (int i = threadIdx.x; i < chunk; i += blockDim.x) sm[i] = (uint8_t)i;
  asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
  __syncthreads();
  const size_t nchunks = nbytes / chunk;
  if (threadIdx.x == 0) {
    for (size_t c = blockIdx.x; c < nchunks; c += gridDim.x) {
      asm volatile("cp.async.bulk.global.shared::cta.bulk_group [%0], [%1], %2;" ::
                   "l"(out + c * chunk), "r"((uint32_t)__cvta_generic_to_shared(sm)), "r"(chunk) : "memory");
      asm volatile("cp.async.bulk.commit_group;" ::: "memory");
      asm volatile("cp.async.bulk.wait_group.read 4;" ::: "memory");
    }
    asm volatile("cp.async.bulk.wait_group 0;" ::: "memory");
  }
}
// mixed: read r bytes per 3r written (the step kernel's mix)
__global__ void k_mix(const uint4* in, uint4* out, size_t n16) {
  for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < n16 / 4; i += (size_t)gridDim.x * blockDim.x) {
    uint4 v = in[i];
    __stcs(out + 4 * i, v); __stcs(out + 4 * i + 1, v); __stcs(out + 4 * i + 2, v); __stcs(out + 4 * i + 3, v);
  }
}

template <class F> float best_ms(F f) {
  cudaEvent_t a, b; cudaEventCreate(&a); cudaEventCreate(&b);
  float best = 1e9f;
  for (int r = 0; r < 8; ++r) {
    cudaEventRecord(a); f(); cudaEventRecord(b); cudaEventSynchronize(b);
    float ms; cudaEventElapsedTime(&ms, a, b); if (r >= 2 && ms < best) best = ms;
  }
  return best;
}

int main() {
  const size_t nbytes = 1ull << 30, n16 = nbytes / 16;
  uint4 *out, *in; CK(cudaMalloc(&out, nbytes)); CK(cudaMalloc(&in, nbytes)); CK(cudaMemset(in, 1, nbytes));
  printf("memset            %7.0f GB/s\n", nbytes / best_ms([&] { cudaMemset(out, 0, nbytes); }) / 1e6);
  for (int ctas : {148 * 2, 148 * 8, 148 * 16}) for (int thr : {128, 512}) {
    printf("grid %5d x %3d  plain %7.0f  .cs %7.0f  .wt %7.0f  .cg %7.0f GB/s\n", ctas, thr,
           nbytes / best_ms([&] { k_store<0><<<ctas, thr>>>(out, n16); }) / 1e6,
           nbytes / best_ms([&] { k_store<1><<<ctas, thr>>>(out, n16); }) / 1e6,
           nbytes / best_ms([&] { k_store<2><<<ctas, thr>>>(out, n16); }) / 1e6,
           nbytes / best_ms([&] { k_store<3><<<ctas, thr>>>(out, n16); }) / 1e6);
  }
  for (int chunk : {1024, 5264, 16384}) {
    printf("chunked %6d B/CTA-chunk: plain %7.0f  .cs %7.0f GB/s\n", chunk * 16,
           nbytes / best_ms([&] { k_store_chunk<0><<<148 * 6, 128>>>(out, n16, chunk); }) / 1e6,
           nbytes / best_ms([&] { k_store_chunk<1><<<148 * 6, 128>>>(out, n16, chunk); }) / 1e6);
  }
  for (int chunk : {16384, 65536}) for (int ctas : {148, 148 * 2, 148 * 3}) {
    CK(cudaFuncSetAttribute(k_bulk, cudaFuncAttributeMaxDynamicSharedMemorySize, chunk));
    float ms = best_ms([&] { k_bulk<<<ctas, 128, chunk>>>((uint8_t*)out, nbytes, chunk); });
    CK(cudaGetLastError());
    printf("TMA bulk store chunk %6d, %4d CTAs: %7.0f GB/s\n", chunk, ctas, nbytes / ms / 1e6);
  }
  printf("mix 1r:4w (bytes moved = 1.25 x out) %7.0f GB/s total, writes alone %7.0f GB/s\n",
         1.25 * nbytes / best_ms([&] { k_mix<<<148 * 8, 256>>>(in, out, n16); }) / 1e6,
         nbytes / best_ms([&] { k_mix<<<148 * 8, 256>>>(in, out, n16); }) / 1e6);
  CK(cudaDeviceSynchronize());
  return 0;
}
