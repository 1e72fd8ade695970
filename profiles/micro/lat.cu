// Dependent-chain latencies of the operations the history kernel is built from (1 warp, clock64).
#include <cstdio>
#include <cuda_runtime.h>
#include <stdint.h>
#define N 2048
template <int OP> __global__ void k(double a, double b, uint64_t s0, double* out, long long* cyc) {
  double x = a; uint64_t s = s0; int q = (int)a;
  long long t0 = clock64();
#pragma unroll 16
  for (int i = 0; i < N; ++i) {
    if (OP == 0) x = __dadd_rn(x, b);
    if (OP == 1) x = __dmul_rn(x, b);
    if (OP == 2) x = __fma_rn(x, b, a);
    if (OP == 3) x = __ddiv_rn(a, x) + b;          // div + add
    if (OP == 4) x = __dsqrt_rn(x) + b;            // sqrt + add
    if (OP == 5) s = (2806196910506780709ULL * s + 1ULL) & 0x7fffffffffffffffULL;
    if (OP == 6) x = floor(x * b) + a;
    if (OP == 7) { q = (int)x; x = (double)q + b; }            // F2I + I2F + add
    if (OP == 8) x = fmax(fabs(x - a) - b, x);
    if (OP == 9) { float f = (float)x; f = __fmaf_rn(f, 1.0001f, 0.5f); x = (double)f; }   // F2F both ways + ffma
    if (OP == 10) { x = __longlong_as_double(__double_as_longlong(x) + 1); x = __dadd_rn(x, b); }
  }
  long long t1 = clock64();
  if (threadIdx.x == 0) { out[OP] = x + (double)s + q; cyc[OP] = t1 - t0; }
}
__global__ void klds(const int* g, int* out, long long* cyc, int slot) {
  __shared__ int sm[1024];
  for (int i = threadIdx.x; i < 1024; i += 32) sm[i] = g[i];
  __syncwarp();
  int j = threadIdx.x;
  long long t0 = clock64();
#pragma unroll 16
  for (int i = 0; i < N; ++i) j = sm[j];
  long long t1 = clock64();
  if (threadIdx.x == 0) { out[0] = j; cyc[slot] = t1 - t0; }
}
__global__ void kldg(const int* g, int* out, long long* cyc, int slot) {
  int j = threadIdx.x;
  long long t0 = clock64();
#pragma unroll 16
  for (int i = 0; i < N; ++i) j = __ldg(g + j);
  long long t1 = clock64();
  if (threadIdx.x == 0) { out[1] = j; cyc[slot] = t1 - t0; }
}
int main() {
  double* out; long long* cyc; int* g; int* io;
  cudaMalloc(&out, 64 * 8); cudaMallocManaged(&cyc, 64 * 8); cudaMalloc(&g, 4096); cudaMalloc(&io, 64);
  int hg[1024]; for (int i = 0; i < 1024; ++i) hg[i] = (i * 33 + 7) & 1023;
  cudaMemcpy(g, hg, 4096, cudaMemcpyHostToDevice);
  const char* names[] = {"DADD", "DMUL", "DFMA", "DDIV+DADD", "DSQRT+DADD", "LCG 64-bit mul-add-and", "DMUL+floor+DADD", "F2I+I2F+DADD", "sub,abs,sub,max", "F2F.32<-64,FFMA,F2F.64<-32", "IADD64(bits)+DADD"};
  for (int rep = 0; rep < 2; ++rep) {
    k<0><<<1, 32>>>(1.0, 1e-9, 1, out, cyc); k<1><<<1, 32>>>(1.0, 1.0000001, 1, out, cyc); k<2><<<1, 32>>>(1.0, 0.5, 1, out, cyc);
    k<3><<<1, 32>>>(1.5, 0.25, 1, out, cyc); k<4><<<1, 32>>>(1.5, 0.25, 1, out, cyc); k<5><<<1, 32>>>(1.5, 0.25, 12345, out, cyc);
    k<6><<<1, 32>>>(1.5, 0.75, 1, out, cyc); k<7><<<1, 32>>>(1.5, 0.75, 1, out, cyc); k<8><<<1, 32>>>(1.5, 0.75, 1, out, cyc);
    k<9><<<1, 32>>>(1.5, 0.75, 1, out, cyc); k<10><<<1, 32>>>(1.5, 0.75, 1, out, cyc);
    klds<<<1, 32>>>(g, io, cyc, 11); kldg<<<1, 32>>>(g, io, cyc, 12);
    cudaDeviceSynchronize();
  }
  for (int i = 0; i < 11; ++i) printf("%-30s %7.1f cycles per iteration\n", names[i], (double)cyc[i] / N);
  printf("%-30s %7.1f cycles per iteration\n", "LDS dependent", (double)cyc[11] / N);
  printf("%-30s %7.1f cycles per iteration\n", "LDG (L1 hit) dependent", (double)cyc[12] / N);
  return 0;
}
