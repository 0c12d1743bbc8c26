// Microbenchmark behind DESIGN.md 4.2 "operands in shared memory": the hot loop of the curve kernel
// (Straus over two 9-entry tables, 34 four-bit windows: 132 doublings + 68 additions) in two forms
//   A  the library's form: completed point in registers, out-of-line multiplier with register arguments
//      (ptxas shuffles ~27 registers per call, half of them as IMAD.MOV on the saturated FMA-heavy pipe);
//   B  operands resident in shared memory (one conflict-free 16-byte column per thread and half-element),
//      out-of-line field operations that take SLOT NUMBERS: no argument marshalling at all.
// Same field operations in the same order, so the outputs must be bit-identical (checked).
// usage: curve_vm_bench [log2n] [reps]
#include <cstdio>
#include <cstdlib>
#include <cstdint>
#include <vector>
#include <cuda_runtime.h>

#include "../schnorr_b200/csrc/ed.cuh"
using namespace sb200;

#define CK(x) do { cudaError_t e = (x); if (e != cudaSuccess) { fprintf(stderr, "CUDA error %s at %s:%d\n", cudaGetErrorString(e), __FILE__, __LINE__); return 1; } } while (0)

constexpr int TPB = 128;
#ifndef VM_CTAS
#define VM_CTAS 6
#endif
#ifndef VM_SLOTS
#define VM_SLOTS 8
#endif

__device__ __forceinline__ fq ldg_fq(const uint32_t* p) {
  const uint4* q = reinterpret_cast<const uint4*>(p);
  uint4 a = __ldg(q), b = __ldg(q + 1);
  fq r = {{a.x, a.y, a.z, a.w, b.x, b.y, b.z, b.w}};
  return r;
}
__device__ __forceinline__ void stg_fq(uint32_t* p, const fq& v) {
  uint4* q = reinterpret_cast<uint4*>(p);
  q[0] = make_uint4(v.v[0], v.v[1], v.v[2], v.v[3]);
  q[1] = make_uint4(v.v[4], v.v[5], v.v[6], v.v[7]);
}

// in: pts [n][2][16] (two affine "points" per tuple), ks [n][2][8] (offset-recoded scalars); out [n][32] = E,F,G,H
__global__ void __launch_bounds__(TPB, 4) k_regs(int64_t n, const uint32_t* pts, const uint32_t* ks, uint32_t* out, int nwin) {
  int64_t i = (int64_t)blockIdx.x * TPB + threadIdx.x;
  if (i >= n) i = n - 1;
  pniels tabs[18];
#pragma unroll 1
  for (int t = 0; t < 2; t++) vartable_build(tabs + 9 * t, affine_to_ext(ldg_fq(pts + i * 32 + 16 * t), ldg_fq(pts + i * 32 + 16 * t + 8)));
  uint32_t k1[8], k2[8];
#pragma unroll
  for (int j = 0; j < 8; j++) { k1[j] = ks[i * 16 + j]; k2[j] = ks[i * 16 + 8 + j]; }
  p1p1 c = ed_mul_var2_rolled(tabs, k1, tabs + 9, k2, nwin);
  stg_fq(out + i * 32, c.E); stg_fq(out + i * 32 + 8, c.F); stg_fq(out + i * 32 + 16, c.G); stg_fq(out + i * 32 + 24, c.H);
}

// ---- form B -----------------------------------------------------------------------------------------------------
extern __shared__ uint4 vm_smem[];  // [slot][half][thread]
__device__ __forceinline__ fq vm_ld(int s) {
  const uint4* p = vm_smem + (s * 2) * TPB + threadIdx.x;
  uint4 a = p[0], b = p[TPB];
  fq r = {{a.x, a.y, a.z, a.w, b.x, b.y, b.z, b.w}};
  return r;
}
__device__ __forceinline__ void vm_st(int s, const fq& v) {
  uint4* p = vm_smem + (s * 2) * TPB + threadIdx.x;
  p[0] = make_uint4(v.v[0], v.v[1], v.v[2], v.v[3]);
  p[TPB] = make_uint4(v.v[4], v.v[5], v.v[6], v.v[7]);
}
__device__ __noinline__ void vm_mul(int d, int a, int b) { vm_st(d, fq_mul_inl(vm_ld(a), vm_ld(b))); }
__device__ __noinline__ void vm_sqr(int d, int a) { vm_st(d, fq_sqr_inl(vm_ld(a))); }
__device__ __noinline__ void vm_sqr_sum(int d, int a, int b) { vm_st(d, fq_sqr_inl(fq_add(vm_ld(a), vm_ld(b)))); }
// d = (a + sign b) * (*q)   with q in thread-private (local / global) memory
__device__ __noinline__ void vm_mul_pm_g(int d, int a, int b, int minus, const fq* q) {
  fq x = vm_ld(a), y = vm_ld(b);
  fq s = minus ? fq_sub(x, y) : fq_add(x, y);
  vm_st(d, fq_mul_inl(s, *q));
}
__device__ __noinline__ void vm_mul_g(int d, int a, const fq* q) { vm_st(d, fq_mul_inl(vm_ld(a), *q)); }
// slots 0..3 = A, B, C, S  ->  E, F, G, H of the doubled point
__device__ __noinline__ void vm_tail_dbl() {
  fq A = vm_ld(0), B = vm_ld(1), C = fq_dbl(vm_ld(2)), S = vm_ld(3);
  fq H = fq_add(A, B), G = fq_sub(B, A);
  vm_st(3, H); vm_st(2, G); vm_st(0, fq_sub(S, H)); vm_st(1, fq_sub(C, G));
}
// slots 0..3 = A, B, C, D  ->  E, F, G, H of the sum; neg: the table entry was to be subtracted (C = T * 2dT' changes sign)
__device__ __noinline__ void vm_tail_add(int neg) {
  fq A = vm_ld(0), B = vm_ld(1), C = vm_ld(2), D = fq_dbl(vm_ld(3));
  fq F = fq_sub(D, C), G = fq_add(D, C);
  vm_st(0, fq_sub(B, A)); vm_st(3, fq_add(B, A));
  vm_st(1, neg ? G : F); vm_st(2, neg ? F : G);
}
__device__ __forceinline__ void vm_dbl() {
  vm_mul(4, 0, 1); vm_mul(5, 2, 3); vm_mul(6, 1, 2);
  vm_sqr(0, 4); vm_sqr(1, 5); vm_sqr(2, 6); vm_sqr_sum(3, 4, 5);
  vm_tail_dbl();
}
__device__ __forceinline__ void vm_add(const pniels* tab, int dgt) {
  const int neg = dgt < 0;
  const pniels* q = tab + (neg ? -dgt : dgt);
  vm_mul(4, 0, 1); vm_mul(5, 2, 3); vm_mul(6, 1, 2); vm_mul(7, 0, 3);  // X, Y, Z, T
  vm_mul_pm_g(0, 5, 4, 1, neg ? &q->YpX : &q->YmX);
  vm_mul_pm_g(1, 5, 4, 0, neg ? &q->YmX : &q->YpX);
  vm_mul_g(2, 7, &q->T2d);
  vm_mul_g(3, 6, &q->Z);
  vm_tail_add(neg);
}
__global__ void __launch_bounds__(TPB, VM_CTAS) k_smem(int64_t n, const uint32_t* pts, const uint32_t* ks, uint32_t* out, int nwin) {
  int64_t i = (int64_t)blockIdx.x * TPB + threadIdx.x;
  if (i >= n) i = n - 1;
  pniels tabs[18];
#pragma unroll 1
  for (int t = 0; t < 2; t++) vartable_build(tabs + 9 * t, affine_to_ext(ldg_fq(pts + i * 32 + 16 * t), ldg_fq(pts + i * 32 + 16 * t + 8)));
  uint32_t k1[8], k2[8];
#pragma unroll
  for (int j = 0; j < 8; j++) { k1[j] = ks[i * 16 + j]; k2[j] = ks[i * 16 + 8 + j]; }
  {  // top window, as ed_mul_var2_rolled: identity + entry, then + entry
    p1p1 c = ed_add(ext_identity(), vartable_lookup(tabs, recode_digit<4>(k1, nwin - 1)));
    c = ed_add(p1p1_to_ext(c), vartable_lookup(tabs + 9, recode_digit<4>(k2, nwin - 1)));
    vm_st(0, c.E); vm_st(1, c.F); vm_st(2, c.G); vm_st(3, c.H);
  }
#pragma unroll 1
  for (int w = nwin - 2; w >= 0; w--) {
#pragma unroll 1
    for (int k = 0; k < 4; k++) vm_dbl();
#pragma unroll 1
    for (int t = 0; t < 2; t++) vm_add(t ? tabs + 9 : tabs, recode_digit<4>(t ? k2 : k1, w));
  }
  stg_fq(out + i * 32, vm_ld(0)); stg_fq(out + i * 32 + 8, vm_ld(1)); stg_fq(out + i * 32 + 16, vm_ld(2)); stg_fq(out + i * 32 + 24, vm_ld(3));
}

int main(int argc, char** argv) {
  const int lg = argc > 1 ? atoi(argv[1]) : 20, reps = argc > 2 ? atoi(argv[2]) : 3, nwin = 34;
  const int64_t n = 1ll << lg;
  std::vector<uint32_t> hp((size_t)n * 32), hk((size_t)n * 16);
  uint64_t s = 88172645463325252ull;
  auto rnd = [&]() { s ^= s << 13; s ^= s >> 7; s ^= s << 17; return (uint32_t)(s >> 16); };
  for (auto& x : hp) x = rnd();
  for (size_t j = 7; j < hp.size(); j += 8) hp[j] &= 0x3fffffffu;  // < q
  for (auto& x : hk) x = rnd();
  uint32_t *dp, *dk, *o1, *o2;
  CK(cudaMalloc(&dp, hp.size() * 4)); CK(cudaMalloc(&dk, hk.size() * 4)); CK(cudaMalloc(&o1, (size_t)n * 128)); CK(cudaMalloc(&o2, (size_t)n * 128));
  CK(cudaMemcpy(dp, hp.data(), hp.size() * 4, cudaMemcpyHostToDevice)); CK(cudaMemcpy(dk, hk.data(), hk.size() * 4, cudaMemcpyHostToDevice));
  const size_t smem = (size_t)VM_SLOTS * 2 * TPB * 16;
  CK(cudaFuncSetAttribute(k_smem, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  const unsigned grid = (unsigned)((n + TPB - 1) / TPB);
  cudaEvent_t e0, e1; CK(cudaEventCreate(&e0)); CK(cudaEventCreate(&e1));
  float ms[2];
  for (int v = 0; v < 2; v++) {
    for (int r = -1; r < reps; r++) {
      if (r == 0) CK(cudaEventRecord(e0));
      if (v == 0) k_regs<<<grid, TPB>>>(n, dp, dk, o1, nwin); else k_smem<<<grid, TPB, smem>>>(n, dp, dk, o2, nwin);
    }
    CK(cudaEventRecord(e1)); CK(cudaEventSynchronize(e1)); CK(cudaGetLastError());
    CK(cudaEventElapsedTime(&ms[v], e0, e1)); ms[v] /= reps;
  }
  std::vector<uint32_t> h1((size_t)n * 32), h2((size_t)n * 32);
  CK(cudaMemcpy(h1.data(), o1, h1.size() * 4, cudaMemcpyDeviceToHost)); CK(cudaMemcpy(h2.data(), o2, h2.size() * 4, cudaMemcpyDeviceToHost));
  size_t bad = 0;
  for (size_t j = 0; j < h1.size(); j++) bad += h1[j] != h2[j];
  int occ = 0; cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ, k_smem, TPB, smem);
  cudaFuncAttributes fa; cudaFuncGetAttributes(&fa, k_smem);
  printf("{\"log2n\": %d, \"regs_ms\": %.3f, \"smem_ms\": %.3f, \"speedup\": %.4f, \"mismatching_words\": %zu, \"smem_ctas_per_sm\": %d, \"smem_regs\": %d, \"smem_local_bytes\": %zu}\n",
         lg, ms[0], ms[1], ms[0] / ms[1], bad, occ, fa.numRegs, (size_t)fa.localSizeBytes);
  return bad != 0;
}
