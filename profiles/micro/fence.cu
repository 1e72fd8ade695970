// latency of CTA-scope synchronisation primitives for one thread that has a shared-memory store in flight (B200):
// what a warp pays per hand-off to another warp of its CTA.  nvcc -arch=sm_100a -o fence fence.cu && ./fence
#include <cstdio>
#include <cstdint>
__global__ void k(long long* out, int n) {
  __shared__ volatile unsigned flag[32];
  __shared__ double data[64];
  __shared__ __align__(8) uint64_t bar;
  unsigned barA = (unsigned)__cvta_generic_to_shared(&bar);
  unsigned flagA = (unsigned)__cvta_generic_to_shared((void*)&flag[0]);
  if (threadIdx.x == 0) { asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(barA)); }
  __syncthreads();
  if (threadIdx.x != 0) return;
  long long t0, t1; double acc = 0;
  // 0: baseline store + volatile flag store
  t0 = clock64();
  for (int i = 0; i < n; ++i) { data[i & 63] = acc; flag[0] = i; acc += 1.0; }
  t1 = clock64(); out[0] = t1 - t0;
  t0 = clock64();
  for (int i = 0; i < n; ++i) { data[i & 63] = acc; __threadfence_block(); flag[0] = i; acc += 1.0; }
  t1 = clock64(); out[1] = t1 - t0;
  t0 = clock64();
  for (int i = 0; i < n; ++i) { data[i & 63] = acc; asm volatile("fence.acq_rel.cta;" ::: "memory"); flag[0] = i; acc += 1.0; }
  t1 = clock64(); out[2] = t1 - t0;
  t0 = clock64();
  for (int i = 0; i < n; ++i) { data[i & 63] = acc; asm volatile("st.release.cta.shared.u32 [%0], %1;" ::"r"(flagA), "r"(i) : "memory"); acc += 1.0; }
  t1 = clock64(); out[3] = t1 - t0;
  t0 = clock64();
  for (int i = 0; i < n; ++i) {
    data[i & 63] = acc;
    uint64_t st; asm volatile("mbarrier.arrive.shared::cta.b64 %0, [%1];" : "=l"(st) : "r"(barA) : "memory");
    unsigned done = 0;
    while (!done) asm volatile("{ .reg .pred p; mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2; selp.u32 %0, 1, 0, p; }" : "=r"(done) : "r"(barA), "r"(i & 1) : "memory");
    acc += 1.0;
  }
  t1 = clock64(); out[4] = t1 - t0;
  // 5: dependent volatile load after store (round trip through the shared pipe)
  t0 = clock64();
  for (int i = 0; i < n; ++i) { flag[1] = i; unsigned v = flag[1]; acc += v; }
  t1 = clock64(); out[5] = t1 - t0;
  t0 = clock64();
  for (int i = 0; i < n; ++i) { unsigned v; asm volatile("ld.acquire.cta.shared.u32 %0, [%1];" : "=r"(v) : "r"(flagA) : "memory"); acc += v + data[i & 63]; }
  t1 = clock64(); out[6] = t1 - t0;
  data[0] = acc;
}
int main() {
  long long* d; cudaMalloc(&d, 64); const int n = 2000;
  k<<<1, 64>>>(d, n); k<<<1, 64>>>(d, n);
  long long h[8]; cudaMemcpy(h, d, 56, cudaMemcpyDeviceToHost);
  const char* nm[] = {"store + volatile flag store", "store + __threadfence_block + flag", "store + fence.acq_rel.cta + flag", "store + st.release.cta.shared", "store + mbarrier arrive + try_wait (own)", "volatile store then dependent load", "ld.acquire.cta.shared + dependent load"};
  for (int i = 0; i < 7; ++i) printf("%-45s %.1f cycles per iteration\n", nm[i], (double)h[i] / n);
  return 0;
}
