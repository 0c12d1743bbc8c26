// Issue-slot microbenchmark for B200 (sm_100a): what does an ALU / FMA-lite / FP64 instruction cost when it is mixed into a
// stream of carry-chained IMAD.WIDE.U32 (the stream every kernel of this library is made of)?
// Loop body = the multiplier's building block (two independent carry chains of 4 wide products each: 8 IMAD.WIDE per
// iteration, as tools/imad_peak.cu mode 3) + K independent instructions of one kind on other registers.  4 warps per SM
// sub-partition (4 CTAs x 128 threads per SM, one wave), like the library's kernels.  Reported per (kind, K):
//   cycles per warp-iteration on one sub-partition = clock64 delta x 4 sub-partitions / (iterations x 16 warps per SM)
// If the extra instructions were free (other pipe, spare issue slots) the figure would stay at K = 0's; the slope of the
// fit is the cost of one such instruction in sub-partition cycles.  Prints one JSON object.
#include <cstdio>
#include <cstdlib>
#include <cstdint>
#include <cuda_runtime.h>

#define CK(x) do { cudaError_t e = (x); if (e != cudaSuccess) { \
  fprintf(stderr, "CUDA error %s at %s:%d\n", cudaGetErrorString(e), __FILE__, __LINE__); return 1; } } while (0)

enum { ADD = 0, XOR = 1, ADDX = 2, SHF = 3, IMADLO = 4, DFMA = 5, SEL = 6 };

template <int KIND, int K>
__global__ void __launch_bounds__(128, 4) k(uint32_t* out, long long* cycles, uint32_t seed, int iters) {
  uint32_t a = seed + threadIdx.x, b = seed * 2654435761u + blockIdx.x;
  uint32_t b1 = b ^ 0x55555555u, b2 = b + 77u, b3 = b * 3u;
  uint32_t w[16], x[8];
  double d[4];
#pragma unroll
  for (int i = 0; i < 8; i++) { x[i] = a ^ (i * 0x9e3779b9u); w[2 * i] = x[i] * 3u; w[2 * i + 1] = x[i] ^ b; }
#pragma unroll
  for (int i = 0; i < 4; i++) d[i] = 1.0 + 1e-3 * (double)(x[i] & 1023);
  long long t0 = clock64();
#pragma unroll 4
  for (int it = 0; it < iters; it++) {
#pragma unroll
    for (int h = 0; h < 2; h++) {
      uint32_t* y = w + 8 * h;
      uint32_t mul = w[8 * (1 - h)];
      asm volatile(
          "mad.lo.cc.u32 %0, %8, %9, %0;\n\t"
          "madc.hi.cc.u32 %1, %8, %9, %1;\n\t"
          "madc.lo.cc.u32 %2, %8, %10, %2;\n\t"
          "madc.hi.cc.u32 %3, %8, %10, %3;\n\t"
          "madc.lo.cc.u32 %4, %8, %11, %4;\n\t"
          "madc.hi.cc.u32 %5, %8, %11, %5;\n\t"
          "madc.lo.cc.u32 %6, %8, %12, %6;\n\t"
          "madc.hi.u32 %7, %8, %12, %7;"
          : "+r"(y[0]), "+r"(y[1]), "+r"(y[2]), "+r"(y[3]), "+r"(y[4]), "+r"(y[5]), "+r"(y[6]), "+r"(y[7])
          : "r"(mul), "r"(b), "r"(b1), "r"(b2), "r"(b3));
#pragma unroll
      for (int j = 0; j < K / 2; j++) {
        const int c = (h * (K / 2) + j) % 8;
        const int step = (h * (K / 2) + j) / 8;
        if (KIND == ADD) asm volatile("add.u32 %0, %0, %1;" : "+r"(x[c]) : "r"((step & 1) ? a : b));
        else if (KIND == XOR) asm volatile("xor.b32 %0, %0, %1;" : "+r"(x[c]) : "r"((step & 1) ? a : b));
        else if (KIND == ADDX) {  // a pair = add.cc + addc (IADD3 + IADD3.X); counts as two instructions
          if (((h * (K / 2) + j) & 1) == 0) asm volatile("add.cc.u32 %0, %0, %2;\n\taddc.u32 %1, %1, %3;" : "+r"(x[c]), "+r"(x[(c + 1) % 8]) : "r"(a), "r"(b));
        } else if (KIND == SHF) asm volatile("shf.l.wrap.b32 %0, %0, %1, 1;" : "+r"(x[c]) : "r"(x[(c + 3) % 8])); else if (KIND == IMADLO) asm volatile("mad.lo.u32 %0, %0, %1, %2;" : "+r"(x[c]) : "r"(a), "r"(b));
        else if (KIND == DFMA) asm volatile("fma.rn.f64 %0, %0, %1, %2;" : "+d"(d[c % 4]) : "d"(1.0000001), "d"(1e-9));
        else if (KIND == SEL && ((h * (K / 2) + j) & 1) == 0) asm volatile("{\n\t.reg .pred p;\n\tsetp.lt.u32 p, %0, %1;\n\tselp.u32 %0, %2, %0, p;\n\t}" : "+r"(x[c]) : "r"(b), "r"(a));
      }
    }
  }
  long long t1 = clock64();
  uint32_t acc = 0;
#pragma unroll
  for (int i = 0; i < 8; i++) acc ^= w[2 * i] ^ w[2 * i + 1] ^ x[i];
#pragma unroll
  for (int i = 0; i < 4; i++) acc ^= (uint32_t)__double2loint(d[i]);
  out[blockIdx.x * blockDim.x + threadIdx.x] = acc;
  if (threadIdx.x == 0) cycles[blockIdx.x] = t1 - t0;
}

static int nsm;
static uint32_t* d_out;
static long long* d_cyc;
static bool first = true;

template <int KIND, int K>
int run(const char* kind) {
  int grid = nsm * 4, iters = 1 << 15;
  cudaEvent_t e0, e1; CK(cudaEventCreate(&e0)); CK(cudaEventCreate(&e1));
  float ms = 0;
  for (int pass = 0; pass < 2; pass++) {
    CK(cudaEventRecord(e0));
    k<KIND, K><<<grid, 128>>>(d_out, d_cyc, 999u + pass, iters);
    CK(cudaEventRecord(e1)); CK(cudaEventSynchronize(e1));
    CK(cudaEventElapsedTime(&ms, e0, e1));
  }
  static long long h[148 * 8];
  CK(cudaMemcpy(h, d_cyc, sizeof(long long) * grid, cudaMemcpyDeviceToHost));
  double avg = 0;
  for (int i = 0; i < grid; i++) avg += (double)h[i];
  avg /= grid;
  double per_warp_iter = avg * 4.0 / ((double)iters * 16.0);
  printf("%s\n  {\"kind\": \"%s\", \"extra_per_8_wide\": %d, \"cycles_per_warp_iteration\": %.3f, \"ms\": %.2f, \"wide_per_s\": %.4e}", first ? "" : ",",
         kind, K, per_warp_iter, ms, 8.0 * 32.0 * iters * 4.0 * grid / (ms * 1e-3));
  first = false;
  return 0;
}

template <int KIND>
int sweep(const char* kind) {
  if (run<KIND, 0>(kind)) return 1;
  if (run<KIND, 2>(kind)) return 1;
  if (run<KIND, 4>(kind)) return 1;
  if (run<KIND, 8>(kind)) return 1;
  if (run<KIND, 12>(kind)) return 1;
  if (run<KIND, 16>(kind)) return 1;
  if (run<KIND, 24>(kind)) return 1;
  if (run<KIND, 32>(kind)) return 1;
  return 0;
}

int main() {
  cudaDeviceProp p; CK(cudaGetDeviceProperties(&p, 0));
  nsm = p.multiProcessorCount;
  CK(cudaMalloc(&d_out, sizeof(uint32_t) * nsm * 4 * 128)); CK(cudaMalloc(&d_cyc, sizeof(long long) * nsm * 8));
  printf("{\"gpu\": \"%s\", \"sms\": %d, \"warps_per_subpartition\": 4, \"body\": \"8 carry-chained IMAD.WIDE.U32 + K extra instructions\", \"points\": [", p.name, nsm);
  if (sweep<ADD>("IADD3 (add.u32)")) return 1;
  if (sweep<XOR>("LOP3 (xor.b32)")) return 1;
  if (sweep<ADDX>("IADD3 + IADD3.X pairs (add.cc/addc)")) return 1;
  if (sweep<SHF>("SHF (funnel shift)")) return 1;
  if (sweep<IMADLO>("IMAD (mad.lo.u32)")) return 1;
  if (sweep<DFMA>("DFMA (fma.rn.f64)")) return 1;
  if (sweep<SEL>("ISETP + SEL pairs")) return 1;
  printf("\n]}\n");
  return 0;
}
