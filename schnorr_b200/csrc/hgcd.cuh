// Half-size scalars for verification (Antipa, Brown, Gallant, Lambert, Struik, Vanstone: "Accelerated verification
// of ECDSA signatures", SAC 2005), made exact on the WHOLE curve so that verdicts stay those of
// `PublicKey::verify` (/root/reference/src/keys/public.rs:121-130) for every input, torsion components included.
//
// The reference accepts iff  u G + c PK == R  in the curve group, whose order is N = 8 r.  For any integers (a, b)
// with  a = b c (mod N)  and  gcd(b, N) = 1,  multiplication by b is an automorphism of the group and a PK = b c PK
// for EVERY curve point PK (N kills the group), hence
//        u G + c PK == R    <=>    (b u mod r) G + a PK - b R == identity.
// The lattice {(a, b) : a = b c mod N} has determinant N ~ 2^255, so a vector with both entries ~2^128 exists: the
// extended Euclidean algorithm on (N, c), stopped when the remainder drops below 2^128, yields it
// (remainder_i = t_i c mod N, |t_i| <= N / remainder_(i-1)).  b must be odd (gcd with the cofactor 8; |b| < r makes it
// coprime to r): consecutive cofactors t_i are coprime, so if t_i is even the algorithm simply runs on to the next
// one.  The double-scalar multiplication over the two VARIABLE points then needs 132 doublings instead of 248.
//
// One step per iteration, all lanes of a warp in lock-step: the quotient is under-estimated from the top 64 bits
// (a partial quotient is still a valid Euclid step: the pair is only swapped once the remainder is smaller), the
// cofactor magnitudes add (their signs alternate).  If a cofactor outgrows the 134-bit window budget (pathological
// c: a very short lattice vector), `ok` is false and the caller runs the full-size multiplication for that tuple.
#pragma once
#include "fq.cuh"

namespace sb200 {

#if defined(__CUDA_ARCH__)
#define SB_WARP_ANY(x) __any_sync(0xffffffffu, (x))
#else
#define SB_WARP_ANY(x) (x)
#endif

struct hgcd_res {
  uint32_t a[8];  // 0 <= a < 2^134
  uint32_t b[8];  // |b| < 2^134, odd
  bool bneg;      // b < 0
  bool ok;
};

SB_HD int hgcd_clz(uint32_t x) {
#if defined(__CUDA_ARCH__)
  return __clz((int)x);
#else
  return x ? __builtin_clz(x) : 32;
#endif
}

SB_HD hgcd_res half_gcd_8r(const uint32_t* c) {
  const uint32_t n8r[8] = SB200_8R_INIT;
  uint32_t r0[8], r1[8], m0[5], m1[5];
#pragma unroll
  for (int i = 0; i < 8; i++) {
    r0[i] = n8r[i];
    r1[i] = c[i];
  }
#pragma unroll
  for (int i = 0; i < 5; i++) m0[i] = m1[i] = 0;
  m1[0] = 1;
  bool neg1 = false, overflow = false;  // sign of t1 (t0 has the opposite sign; t0 = 0, t1 = +1 initially)
#pragma unroll 1
  for (int it = 0; it < 320; it++) {
    const bool small = (r1[4] | r1[5] | r1[6] | r1[7]) == 0;
    const bool zero = small && (r1[0] | r1[1] | r1[2] | r1[3]) == 0;
    // the previous pair (r0, t0) is the alternative when t1 is even (t0 is then odd): it fits if r0 < 2^134
    const bool prev_fits = (r0[5] | r0[6] | r0[7]) == 0 && r0[4] < 64u && (m0[0] & 1u);
    const bool done = (small && ((m1[0] & 1u) || prev_fits)) || overflow || (zero && it > 0);
    if (!SB_WARP_ANY(!done)) break;
    if (done) continue;
    // top 64 bits of r0 (r0 >= r1, r0 >= 2^128 unless the pair is in its extra "make b odd" step) and r1 alongside
    int h = 7;
#pragma unroll
    for (int k = 7; k >= 3; k--)
      if (r0[k] == 0 && h == k) h = k - 1;
    uint32_t w2 = 0, w1 = 0, w0 = 0, v2 = 0, v1 = 0, v0 = 0;
#pragma unroll
    for (int k = 2; k < 8; k++) {
      if (h == k) {
        w2 = r0[k]; w1 = r0[k - 1]; w0 = r0[k - 2];
        v2 = r1[k]; v1 = r1[k - 1]; v0 = r1[k - 2];
      }
    }
    const int sh = hgcd_clz(w2) & 31;  // w2 != 0 except when r0 < 2^96: then the estimate below is exact enough anyway
    uint64_t R0 = ((uint64_t)w2 << 32) | w1, R1 = ((uint64_t)v2 << 32) | v1;
    if (sh) {
      R0 = (R0 << sh) | (w0 >> (32 - sh));
      R1 = (R1 << sh) | (v0 >> (32 - sh));
    }
    // q <= floor(R0 / (R1 + 1)) <= floor(r0 / r1), from a double-precision quotient scaled down by 2^-50 (the three roundings
    // -- two conversions and the division -- are each within 2^-53): ~25 instructions instead of the ~90 of a 64-bit integer
    // division, and the same value on the host build (IEEE division, no contraction possible).  An estimate that is low by one
    // only costs an extra step.
    uint64_t q64 = (uint64_t)((double)R0 / ((double)R1 + 1.0) * 0.99999999999999911182158029987);
    if (q64 < 1) q64 = 1;               // r0 >= r1: one subtraction is always possible
    if (q64 > 0x7fffffffull) q64 = 0x7fffffffull;
    const uint32_t q = (uint32_t)q64;
    SB_COUNT(wide, 13);  // 8 + 5 limb products of this step (the roofline's work count includes them)
    // r0 -= q * r1
    uint64_t carry = 0;
    uint32_t borrow = 0;
#pragma unroll
    for (int i = 0; i < 8; i++) {
      uint64_t p = (uint64_t)q * r1[i] + carry;
      carry = p >> 32;
      uint64_t d = (uint64_t)r0[i] - (uint32_t)p - borrow;
      r0[i] = (uint32_t)d;
      borrow = (uint32_t)(d >> 32) & 1u;
    }
    // |t0| += q * |t1|
    carry = 0;
#pragma unroll
    for (int i = 0; i < 5; i++) {
      uint64_t p = (uint64_t)q * m1[i] + m0[i] + carry;
      m0[i] = (uint32_t)p;
      carry = p >> 32;
    }
    overflow |= carry != 0;
    // swap when the remainder dropped below the divisor
    uint32_t t[8];
    const bool lt = sub8(t, r0, r1) != 0;
    if (lt) {
#pragma unroll
      for (int i = 0; i < 8; i++) {
        uint32_t x = r0[i];
        r0[i] = r1[i];
        r1[i] = x;
      }
#pragma unroll
      for (int i = 0; i < 5; i++) {
        uint32_t x = m0[i];
        m0[i] = m1[i];
        m1[i] = x;
      }
      neg1 = !neg1;
    }
  }
  hgcd_res res;
  const bool small = (r1[4] | r1[5] | r1[6] | r1[7]) == 0;
  const bool cur = small && (m1[0] & 1u);  // (r1, t1) if t1 is odd, else the previous pair (r0, t0)
  const bool prev_fits = (r0[5] | r0[6] | r0[7]) == 0 && r0[4] < 64u && (m0[0] & 1u);
#pragma unroll
  for (int i = 0; i < 8; i++) {
    res.a[i] = cur ? (i < 4 ? r1[i] : 0u) : (i < 5 ? r0[i] : 0u);
    res.b[i] = i < 5 ? (cur ? m1[i] : m0[i]) : 0u;
  }
  res.bneg = cur ? neg1 : !neg1;
  res.ok = small && (cur || prev_fits) && !overflow && res.b[4] < 64u;
  return res;
}

}  // namespace sb200
