// Field inversion by a two-level (Lehmer-style) Euclidean algorithm instead of Fermat's a^(q-2).
//
// `fq_inv` (fq.cuh) is 256 squarings + 78 multiplications = 30 864 limb products and ~80 000 other instructions: a third of
// a signature's curve work even when shared between four signatures, a quarter of a key generation's.  The extended Euclid on
// (q, A) needs ~150 nearest-integer quotient steps; done like lat3.cuh in two levels -- INNER rounds on double-precision
// copies of the two remainders and on the 2 x 2 integer transform accumulated so far (exact in doubles), OUTER passes that
// apply the transform EXACTLY to the 288-bit two's-complement remainders and cofactors and refresh the doubles -- it costs
// ~13 outer passes of 8 multiply-subtracts each plus ~150 inner rounds of a division and three fused multiply-adds: about
// 13 000 instructions.  Floating-point error can only make a quotient sub-optimal: every step is a unimodular row operation
// applied exactly, so the invariant  r_i = t_i * A (mod q)  always holds, and the loop ends exactly when a remainder is 0
// (a freshly converted non-zero integer never converts to 0.0), the other being +-gcd = +-1.
// Data-dependent iteration count: NOT constant-time -- the address-oblivious signer (SB200_SIGN_OBLIVIOUS) keeps Fermat.
#pragma once
#include "lat3.cuh"

namespace sb200 {

#ifndef SB_INV_UNROLL
#define SB_INV_UNROLL 1
#endif
constexpr int INV_MAX_OUTER = 24;   // measured 9-14 on random input
constexpr int INV_MAX_INNER = 32;

// a / b to ~40 bits on the device (hardware reciprocal estimate + one Newton step: 4 instructions instead of the ~35 of an IEEE
// division) -- enough for quotients below 2^31, and a quotient that is off by one is only a sub-optimal step; exact on the host.
SB_HD double inv_div(double a, double b) {
#if defined(__CUDA_ARCH__)
  double r;
  asm("rcp.approx.ftz.f64 %0, %1;" : "=d"(r) : "d"(b));
  r = lat3_fma(r, lat3_fma(-b, r, 1.0), r);
  return a * r;
#else
  return a / b;
#endif
}

// A = Montgomery representation a * 2^256 mod q (as the integer it is)  ->  Montgomery representation of a^-1;  0 -> 0
SB_HD fq fq_inv_euclid(const fq& A, bool& ok) {
  const uint32_t qq[8] = SB200_FQ_MOD_INIT;
  uint32_t W[2][2][LAT3_LIMBS];  // [row][r, t][limb], two's complement; rows (q, 0) and (A, 1)
#pragma unroll
  for (int i = 0; i < LAT3_LIMBS; i++) W[0][0][i] = W[0][1][i] = W[1][0][i] = W[1][1][i] = 0;
#pragma unroll
  for (int i = 0; i < 8; i++) {
    W[0][0][i] = qq[i];
    W[1][0][i] = A.v[i];
  }
  W[1][1][0] = 1;
  bool done = false, giveup = false;
#pragma unroll 1
  for (int outer = 0; outer < INV_MAX_OUTER; outer++) {
    if (!SB_WARP_ANY(!done)) break;
#if defined(SB_INV_STATS)
    SB_INV_STATS(outer);
#endif
    double F[2], T[2][2];
    F[0] = lat3_to_double(W[0][0]);
    F[1] = lat3_to_double(W[1][0]);
    T[0][0] = T[1][1] = 1.0;
    T[0][1] = T[1][0] = 0.0;
    bool any_change = false, stop = done;
#pragma unroll 1
    for (int inner = 0; inner < INV_MAX_INNER; inner++) {
      if (!SB_WARP_ANY(!stop)) break;
      if (stop) continue;
      // row 0 = the larger remainder
      const bool sw = fabs(F[1]) > fabs(F[0]);
      {
        double x = F[0], y = F[1];
        F[0] = sw ? y : x; F[1] = sw ? x : y;
#pragma unroll
        for (int k = 0; k < 2; k++) {
          x = T[0][k]; y = T[1][k];
          T[0][k] = sw ? y : x; T[1][k] = sw ? x : y;
        }
      }
      bool changed = sw;
      if (F[1] != 0.0) {
        // quotients clamped to 2^31 (a clamped quotient is a partial, still valid, step): |T| < 2^20 before a round keeps
        // |T| < 2^52 after it -- exact in doubles and within lat3_submul's 52-bit multiplier -- and a round that pushed T past
        // 2^20 is the last of its pass.  A quotient beyond 2^40 (a remainder 40 bits shorter than its predecessor: probability
        // 2^-40 per step on random input, certain for a tiny A) would need ~2^9 clamped steps and more: such a lane gives up
        // and the caller inverts it by Fermat.
        const double qf = inv_div(F[0], F[1]);
        giveup |= fabs(qf) > 1099511627776.0;
        const double q = lat3_clamp(lat3_rint(qf), 2147483648.0);
        if (q != 0.0) {
          changed = true;
          F[0] = lat3_fma(-q, F[1], F[0]);
          T[0][0] = lat3_fma(-q, T[1][0], T[0][0]);
          T[0][1] = lat3_fma(-q, T[1][1], T[0][1]);
        }
      }
      any_change |= changed;
      const double tmax = fmax(fmax(fabs(T[0][0]), fabs(T[0][1])), fmax(fabs(T[1][0]), fabs(T[1][1])));
      stop = !changed || tmax >= 1048576.0;
    }
    // (r, t) rows <- T (r, t) rows, exactly; finished lanes carry the identity.  Fully unrolled: W is indexed statically
    // everywhere and stays in registers (the function is out of line with its own allocation).
#if SB_INV_UNROLL
#pragma unroll
#else
#pragma unroll 1
#endif
    for (int k = 0; k < 2; k++) {
      uint32_t nw[2][LAT3_LIMBS];
#pragma unroll
      for (int v = 0; v < 2; v++) {
#pragma unroll
        for (int i = 0; i < LAT3_LIMBS; i++) nw[v][i] = 0;
#pragma unroll
        for (int j = 0; j < 2; j++) lat3_submul(nw[v], W[j][k], done ? (v == j ? -1.0 : 0.0) : -T[v][j]);
      }
#pragma unroll
      for (int v = 0; v < 2; v++)
#pragma unroll
        for (int i = 0; i < LAT3_LIMBS; i++) W[v][k][i] = nw[v][i];
    }
    done |= !any_change || giveup;
  }
  // one remainder is 0, the other +-gcd: x = +-t of that row
  uint32_t z0 = 0;
#pragma unroll
  for (int i = 0; i < LAT3_LIMBS; i++) z0 |= W[0][0][i];
  const bool use1 = z0 == 0;  // row 0 is the zero remainder: the gcd sits in row 1
  uint32_t r[LAT3_LIMBS], x[LAT3_LIMBS];
#pragma unroll
  for (int i = 0; i < LAT3_LIMBS; i++) {
    r[i] = use1 ? W[1][0][i] : W[0][0][i];
    x[i] = use1 ? W[1][1][i] : W[0][1][i];
  }
  uint32_t hi_or = 0, hi_and = 0xffffffffu;
#pragma unroll
  for (int i = 1; i < LAT3_LIMBS; i++) { hi_or |= r[i]; hi_and &= r[i]; }
  const bool plus = r[0] == 1u && hi_or == 0, minus = r[0] == 0xffffffffu && hi_and == 0xffffffffu;
  // x <- -x when the gcd came out as -1
  {
    uint32_t c = minus ? 1u : 0u;
#pragma unroll
    for (int i = 0; i < LAT3_LIMBS; i++) {
      uint64_t t = (uint64_t)(minus ? ~x[i] : x[i]) + c;
      x[i] = (uint32_t)t;
      c = (uint32_t)(t >> 32);
    }
  }
  // |t| < q: one conditional addition of q brings x into [0, q)
  {
    const bool neg = (x[LAT3_LIMBS - 1] >> 31) != 0;
    uint32_t c = 0;
#pragma unroll
    for (int i = 0; i < 8; i++) {
      uint64_t t = (uint64_t)x[i] + (neg ? qq[i] : 0u) + c;
      x[i] = (uint32_t)t;
      c = (uint32_t)(t >> 32);
    }
  }
  fq xi;
#pragma unroll
  for (int i = 0; i < 8; i++) xi.v[i] = (plus || minus) ? x[i] : 0u;
  // gcd = q (A = 0) is a result (0); anything else that did not reach +-1 -- a lane that gave up, an exhausted pass budget -- is not
  ok = plus || minus || fq_is_zero(A);
  // x = A^-1 = a^-1 2^-256 as a plain integer; the Montgomery product with 2^768 gives a^-1 2^256
  const fq r3 = {{0x439b73afu, 0xc62c1807u, 0x8cf06990u, 0x1b3e0d18u, 0xc7b5f418u, 0x73d13c71u, 0xc8db33e9u, 0x6e2a5bb9u}};  // 2^768 mod q;
  return fq_mul(xi, r3);
}

// a^-1 by the Euclid, with Fermat's form for the lanes that gave up (decided by a warp vote: all lanes stay in lock-step)
SB_HD fq fq_inv_euclid_checked(const fq& a) {
  bool ok;
  fq r = fq_inv_euclid(a, ok);
  if (SB_WARP_ANY(!ok)) {
    const fq slow = fq_inv(a);
    r = fq_select(r, slow, !ok);
  }
  return r;
}
#if defined(__CUDA_ARCH__)
// out of line with its own register allocation (see lat3.cuh SB_LAT3_OOL)
static __device__ __noinline__ void fq_inv_euclid_ool(const fq* a, fq* r) { *r = fq_inv_euclid_checked(*a); }
SB_HD fq fq_inv_fast(const fq& a) {
  fq r;
  fq_inv_euclid_ool(&a, &r);
  return r;
}
#else
SB_HD fq fq_inv_fast(const fq& a) { return fq_inv_euclid_checked(a); }
#endif

}  // namespace sb200
