// Integer-multiply roofline microbenchmark for B200 (sm_100a).
// Measures sustained issue rates of the instructions the Fq/Fr Montgomery code is made of:
//   mad.lo.u32 (IMAD), mad.hi.u32 (IMAD.HI), mad.wide.u32 (IMAD.WIDE), the carry-chained
//   mad.lo.cc/madc.hi.cc pair (IMAD.WIDE.X), add.cc chains (IADD3), and a 1:1 wide+add mix.
// Prints one JSON object: per-instruction results/clk/SM (from clock64 inside the kernel) and
// results/s for the whole GPU (from CUDA events).  Used as the denominator of roofline.frac.
#include <cstdio>
#include <cstdlib>
#include <cstdint>
#include <cuda_runtime.h>

#define SB_MUL_NOINLINE 0
#include "../schnorr_b200/csrc/fq.cuh"

#define CK(x) do { cudaError_t e = (x); if (e != cudaSuccess) { \
  fprintf(stderr, "CUDA error %s at %s:%d\n", cudaGetErrorString(e), __FILE__, __LINE__); return 1; } } while (0)

constexpr int CHAINS = 8;

__device__ __forceinline__ unsigned long long gtimer() {
  unsigned long long t;
  asm volatile("mov.u64 %0, %globaltimer;" : "=l"(t));
  return t;
}
// cycles[4*cta + {0,1,2,3}] = clock64 delta, globaltimer delta (ns), %smid, unused
template <int MODE>
__global__ void __launch_bounds__(256) k(uint32_t* out, long long* cycles, uint32_t seed, int ITERS) {
  uint32_t a = seed + threadIdx.x, b = seed * 2654435761u + blockIdx.x;
  uint32_t b1 = b ^ 0x55555555u, b2 = b + 77u, b3 = b * 3u;
  uint32_t x[CHAINS];   // 32-bit chains
  uint64_t v[CHAINS];   // 64-bit chains (aligned register pairs)
  double d[CHAINS];
  uint32_t w[16];      // 32-bit pairs for the carry-chain mode
#pragma unroll
  for (int i = 0; i < CHAINS; i++) {
    x[i] = a ^ (i * 0x9e3779b9u);
    v[i] = ((uint64_t)(b + i) << 32) | x[i];
    d[i] = 1.0 + 1e-3 * (double)(x[i] & 1023);
    w[2 * i] = x[i] * 3u; w[2 * i + 1] = x[i] ^ b;
  }
  unsigned long long g0 = gtimer();
  long long t0 = clock64();
#pragma unroll (MODE == 7 ? 1 : 8)
  for (int it = 0; it < ITERS; it++) {
    if (MODE == 0) {  // IMAD (mad.lo), multiplicand = running value so ptxas cannot hoist the product
#pragma unroll
      for (int c = 0; c < CHAINS; c++) asm volatile("mad.lo.u32 %0, %0, %1, %2;" : "+r"(x[c]) : "r"(a), "r"(b));
    } else if (MODE == 1) {  // IMAD.HI
#pragma unroll
      for (int c = 0; c < CHAINS; c++) asm volatile("mad.hi.u32 %0, %0, %1, %2;" : "+r"(x[c]) : "r"(a), "r"(b));
    } else if (MODE == 2) {  // IMAD.WIDE.U32: 32x32 + 64 -> 64, no carry out
#pragma unroll
      for (int c = 0; c < CHAINS; c++)
        asm volatile("{\n\t.reg .u32 l,h;\n\tmov.b64 {l,h}, %2;\n\tmad.wide.u32 %0, l, %1, %0;\n\t}" : "+l"(v[c]) : "r"(a), "l"(v[(c + 3) % CHAINS]));
    } else if (MODE == 3) {
      // the multiplier's building block: two independent carry chains of 4 products each,
      // (mad.lo.cc, madc.hi.cc) x4 -> IMAD.WIDE.U32 + 3x IMAD.WIDE.U32.X, accumulating in place.
      // The multiplier operand is a running value so ptxas cannot hoist a product out of the loop.
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
      }
    } else if (MODE == 4) {  // 64-bit add as add.cc/addc pair (ALU-pipe IADD3 [+ IMAD.X])
#pragma unroll
      for (int c = 0; c < CHAINS; c++)
        asm volatile("{\n\t.reg .u32 l,h;\n\tmov.b64 {l,h}, %0;\n\tadd.cc.u32 l, l, %1;\n\taddc.u32 h, h, %2;\n\tmov.b64 %0, {l,h};\n\t}"
                     : "+l"(v[c]) : "r"(a), "r"(b));
    } else if (MODE == 5) {  // mix: 4 wide products + 4 lop3/shf-type ALU ops on other registers
#pragma unroll
      for (int c = 0; c < 4; c++) {
        asm volatile("{\n\t.reg .u32 l,h;\n\tmov.b64 {l,h}, %2;\n\tmad.wide.u32 %0, l, %1, %0;\n\t}" : "+l"(v[c]) : "r"(a), "l"(v[(c + 1) % 4]));
        asm volatile("add.u32 %0, %0, %1;" : "+r"(x[c]) : "r"(b));
        asm volatile("xor.b32 %0, %0, %1;" : "+r"(x[4 + c]) : "r"(a));
      }
    } else if (MODE == 7) {  // the library's Fq Montgomery multiplier, inlined: 120 wide products per call
      sb200::fq fa, fb;
#pragma unroll
      for (int c = 0; c < 8; c++) { fa.v[c] = w[c]; fb.v[c] = w[8 + c]; }
      fa = sb200::fq_mul_inl(fa, fb);
      fb = sb200::fq_mul_inl(fb, fa);
#pragma unroll
      for (int c = 0; c < 8; c++) { w[c] = fa.v[c]; w[8 + c] = fb.v[c]; }
    } else if (MODE == 8 || MODE == 9) {
      // dual-pipe probe: the carry-chain block of MODE 3 (8 wide products on the FMA-heavy pipe) interleaved with
      // 4 independent DFMA (FP64 pipe) [and, MODE 9, 4 add.cc/addc pairs on the ALU pipe].  If the pipes are
      // independent and power allows, the wide-product rate stays at MODE 3's.
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
        for (int c = 0; c < 2; c++) asm volatile("fma.rz.f64 %0, %0, %1, %2;" : "+d"(d[2 * h + c]) : "d"(1.0000001), "d"(1e-9));
        if (MODE == 9) {
#pragma unroll
          for (int c = 0; c < 2; c++)
            asm volatile("{\n\t.reg .u32 l,h;\n\tmov.b64 {l,h}, %0;\n\tadd.cc.u32 l, l, %1;\n\taddc.u32 h, h, %2;\n\tmov.b64 %0, {l,h};\n\t}"
                         : "+l"(v[2 * h + c]) : "r"(a), "r"(b));
        }
      }
    } else if (MODE == 10) {  // DFMA.RZ (the rounding mode the 52-bit-limb product trick needs)
#pragma unroll
      for (int c = 0; c < CHAINS; c++) asm volatile("fma.rz.f64 %0, %0, %1, %2;" : "+d"(d[c]) : "d"(1.0000001), "d"(1e-9));
    } else if (MODE == 11) {  // DADD
#pragma unroll
      for (int c = 0; c < CHAINS; c++) asm volatile("add.rn.f64 %0, %0, %1;" : "+d"(d[c]) : "d"(1e-9));
    } else if (MODE == 6) {  // DFMA
#pragma unroll
      for (int c = 0; c < CHAINS; c++) asm volatile("fma.rn.f64 %0, %0, %1, %2;" : "+d"(d[c]) : "d"(1.0000001), "d"(1e-9));
    }
  }
  long long t1 = clock64();
  unsigned long long g1 = gtimer();
  uint32_t acc = 0;
#pragma unroll
  for (int i = 0; i < CHAINS; i++) acc ^= w[2 * i] ^ w[2 * i + 1] ^ x[i] ^ (uint32_t)v[i] ^ (uint32_t)(v[i] >> 32) ^ (uint32_t)__double2loint(d[i]);
  out[blockIdx.x * blockDim.x + threadIdx.x] = acc;
  if (threadIdx.x == 0) {
    unsigned smid;
    asm volatile("mov.u32 %0, %smid;" : "=r"(smid));
    cycles[4 * blockIdx.x] = t1 - t0;
    cycles[4 * blockIdx.x + 1] = (long long)(g1 - g0);
    cycles[4 * blockIdx.x + 2] = smid;
  }
}

// One wave: grid = SMs x resident CTAs, every CTA resident for the whole launch and doing identical work, so
//   results/clk/SM = ops per thread x threads per SM / (clock64 delta of a CTA)            [SM-clock domain]
//   SM clock held  = clock64 delta / globaltimer delta                                      [measured in the kernel]
//   results/s      = total ops / CUDA-event time of the launch                              [wall domain]
// `target_ms` sizes the launch (>= 1 s for the committed numbers: long enough that the clock record is the
// clock under sustained load, not a boost transient).
template <int MODE>
int run(const char* name, double ops_per_iter, int nsm, int /*unused*/, uint32_t* out, long long* cyc, bool last, int target_ms) {
  int block = 256, resident = 0;
  CK(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&resident, k<MODE>, block, 0));
  if (MODE == 7 && resident > 4) resident = 4;  // the library's kernels run 4 CTAs x 128 threads; same warps/SM here: 4 x 256 / 2
  int grid = nsm * resident;
  cudaEvent_t e0, e1; CK(cudaEventCreate(&e0)); CK(cudaEventCreate(&e1));
  int iters = 1 << 12;
  float ms = 0;
  for (int pass = 0; pass < 3; pass++) {  // calibrate, then the measured launch
    CK(cudaEventRecord(e0));
    k<MODE><<<grid, block>>>(out, cyc, 999u + pass, iters);
    CK(cudaEventRecord(e1)); CK(cudaEventSynchronize(e1));
    CK(cudaEventElapsedTime(&ms, e0, e1));
    if (pass < 2) {
      double want = pass == 0 ? 50.0 : (double)target_ms;
      double f = want / (ms > 1e-3 ? ms : 1e-3);
      iters = (int)(iters * f); iters = (iters + 7) & ~7; if (iters < 8) iters = 8;
    }
  }
  long long* h = (long long*)malloc(sizeof(long long) * 4 * grid);
  CK(cudaMemcpy(h, cyc, sizeof(long long) * 4 * grid, cudaMemcpyDeviceToHost));
  double cyc_avg = 0, cyc_max = 0, ns_avg = 0; int per_sm[1024] = {0};
  for (int i = 0; i < grid; i++) { cyc_avg += (double)h[4 * i]; ns_avg += (double)h[4 * i + 1]; if (h[4 * i] > cyc_max) cyc_max = (double)h[4 * i]; per_sm[h[4 * i + 2] & 1023]++; }
  cyc_avg /= grid; ns_avg /= grid;
  int sm_lo = 1 << 30, sm_hi = 0;
  for (int i = 0; i < nsm; i++) { if (per_sm[i] < sm_lo) sm_lo = per_sm[i]; if (per_sm[i] > sm_hi) sm_hi = per_sm[i]; }
  free(h);
  double total_ops = ops_per_iter * iters * (double)block * grid;
  double per_clk_sm = ops_per_iter * iters * (double)block * resident / cyc_avg;
  double mhz = cyc_avg / ns_avg * 1e3;
  printf("  \"%s\": {\"per_s\": %.4e, \"per_clk_sm\": %.3f, \"ms\": %.2f, \"sm_mhz_in_kernel\": %.1f, \"resident_ctas\": %d, "
         "\"ctas_per_sm_min_max\": [%d, %d], \"cycles_max_over_avg\": %.4f, \"per_s_from_cycles\": %.4e}%s\n", name,
         total_ops / (ms * 1e-3), per_clk_sm, ms, mhz, resident, sm_lo, sm_hi, cyc_max / cyc_avg, per_clk_sm * nsm * mhz * 1e6, last ? "" : ",");
  return 0;
}

int main(int argc, char** argv) {
  int ITERS = argc > 1 ? atoi(argv[1]) : 1000;  // target milliseconds per measured launch
  cudaDeviceProp p; CK(cudaGetDeviceProperties(&p, 0));
  int nsm = p.multiProcessorCount, cps = 8;  // 8 CTAs x 256 thr = 64 warps/SM (full occupancy)
  uint32_t* out; long long* cyc;
  CK(cudaMalloc(&out, sizeof(uint32_t) * nsm * cps * 256)); CK(cudaMalloc(&cyc, sizeof(long long) * 4 * nsm * cps));
  printf("{\n  \"gpu\": \"%s\", \"sms\": %d, \"clock_khz_max\": %d,\n", p.name, nsm, p.clockRate);
  if (run<0>("imad_lo", CHAINS, nsm, cps, out, cyc, false, ITERS)) return 1;
  if (run<1>("imad_hi", CHAINS, nsm, cps, out, cyc, false, ITERS)) return 1;
  if (run<2>("imad_wide", CHAINS, nsm, cps, out, cyc, false, ITERS)) return 1;
  if (run<3>("imad_wide_cc", 8, nsm, cps, out, cyc, false, ITERS)) return 1;   // 8 wide products / iter
  if (run<4>("iadd64_pair", CHAINS, nsm, cps, out, cyc, false, ITERS)) return 1;  // 16 adds / iter
  if (run<5>("mix_wide4_alu8", 4, nsm, cps, out, cyc, false, ITERS)) return 1;  // counts the 4 wide products
  if (run<6>("dfma", CHAINS, nsm, cps, out, cyc, false, ITERS)) return 1;
  if (run<10>("dfma_rz", CHAINS, nsm, cps, out, cyc, false, ITERS)) return 1;
  if (run<11>("dadd", CHAINS, nsm, cps, out, cyc, false, ITERS)) return 1;
  if (run<8>("wide8_dfma4__wide", 8, nsm, cps, out, cyc, false, ITERS)) return 1;          // counts the 8 wide products
  if (run<9>("wide8_dfma4_iadd4__wide", 8, nsm, cps, out, cyc, false, ITERS)) return 1;    // counts the 8 wide products
  if (run<7>("fq_mul_wide_products", 240, nsm, 4, out, cyc, true, ITERS)) return 1;  // 2 muls x 120 products / iter
  printf("}\n");
  return 0;
}
