// Short scalars for the variable-generator verification (`PublicKeyVarGen::verify`,
// /root/reference/src/keys/public.rs:401-415):  u Gen + c PK == R  with BOTH bases per-tuple variables.
//
// hgcd.cuh halves ONE variable-base scalar; here there are two, so the same idea needs a 3-dimensional lattice.
// The curve group has order N = 8 r.  For integers (b, a, d) with
//        a = b c (mod N),   d = b u (mod N),   gcd(b, N) = 1
// multiplication by b is an automorphism of the group and a PK = b c PK, d Gen = b u Gen for EVERY pair of curve
// points (N kills the group), hence
//        u Gen + c PK == R    <=>    d Gen + a PK - b R == identity
// exactly, torsion components included.  The lattice L = {(b, a, d)} has the basis (1, c, u), (0, N, 0), (0, 0, N) and
// determinant N^2 ~ 2^510, so its short vectors have entries ~2^170: a triple-scalar Straus multiplication with 44
// four-bit windows (172 doublings) replaces the double-scalar one with 64 (252 doublings).
//
// The reduction is a greedy (Semaev-style) one, shaped for one tuple per thread with all lanes of a warp in lock-step:
// every round sorts the three basis vectors by length, takes the middle one modulo the shortest (a Gauss step) and the
// longest modulo the plane of the other two, with the integer quotients taken from a double-precision Gram matrix of
// the CURRENT vectors and applied EXACTLY to the 288-bit integer vectors.  Floating-point error can only make a step
// sub-optimal, never wrong: every basis vector is an integer combination of lattice vectors at all times.  ~70 rounds
// (measured: 50-87 over random inputs) bring the entries from 2^255 down to <= 2^173; the shortest vector with an odd b
// is used (the b's of a basis cannot all be even: (1, c, u) is in the lattice).  If none fits the 174-bit window budget
// `ok` is false and the caller runs the full-size multiplication for that tuple.
#pragma once
#include "fq.cuh"
#include "hgcd.cuh"

#if !defined(__CUDA_ARCH__)
#include <cmath>
#endif

namespace sb200 {

constexpr int LAT3_LIMBS = 9;        // 288-bit two's complement
constexpr int LAT3_WINDOWS = 44;     // 4-bit windows of the triple-scalar multiplication: magnitudes < 2^174
constexpr int LAT3_MAX_ROUNDS = 160;

struct lat3_res {
  uint32_t a[8], b[8], d[8];  // magnitudes, < 2^174
  bool aneg, bneg, dneg;
  bool ok;
};

SB_HD double lat3_fma(double x, double y, double z) {
#if defined(__CUDA_ARCH__)
  return __fma_rn(x, y, z);
#else
  return std::fma(x, y, z);
#endif
}
SB_HD double lat3_rint(double x) {
#if defined(__CUDA_ARCH__)
  return rint(x);
#else
  return std::nearbyint(x);
#endif
}

// two's complement 288-bit integer -> double (Horner from the top limb: the sign-extension limbs cancel exactly)
SB_HD double lat3_to_double(const uint32_t* v) {
  double d = (double)(int32_t)v[LAT3_LIMBS - 1];
#pragma unroll
  for (int i = LAT3_LIMBS - 2; i >= 0; i--) d = lat3_fma(d, 4294967296.0, (double)v[i]);
  return d;
}

// w -= q * x   for an integer-valued double |q| <= 2^52, exact modulo 2^288 (two's complement)
SB_HD void lat3_submul(uint32_t* w, const uint32_t* x, double q) {
  const bool neg = q < 0.0;
  const uint64_t m = (uint64_t)(neg ? -q : q);
  const uint32_t m0 = (uint32_t)m, m1 = (uint32_t)(m >> 32);
  uint32_t p[LAT3_LIMBS];
  uint64_t carry = 0;
#pragma unroll
  for (int i = 0; i < LAT3_LIMBS; i++) {
    uint64_t t = (uint64_t)m0 * x[i] + carry;
    p[i] = (uint32_t)t;
    carry = t >> 32;
  }
  carry = 0;
#pragma unroll
  for (int i = 0; i + 1 < LAT3_LIMBS; i++) {
    uint64_t t = (uint64_t)m1 * x[i] + p[i + 1] + carry;
    p[i + 1] = (uint32_t)t;
    carry = t >> 32;
  }
  SB_COUNT(wide, 2 * LAT3_LIMBS - 1);
  // w = neg ? w + p : w - p
  uint32_t c = neg ? 0u : 1u;  // subtraction as w + ~p + 1
#pragma unroll
  for (int i = 0; i < LAT3_LIMBS; i++) {
    uint64_t t = (uint64_t)w[i] + (neg ? p[i] : ~p[i]) + c;
    w[i] = (uint32_t)t;
    c = (uint32_t)(t >> 32);
  }
}

SB_HD void lat3_cswap(uint32_t (*A)[LAT3_LIMBS], uint32_t (*B)[LAT3_LIMBS], double* fa, double* fb, double& na, double& nb) {
  const bool sw = na > nb;
#pragma unroll
  for (int k = 0; k < 3; k++) {
#pragma unroll
    for (int i = 0; i < LAT3_LIMBS; i++) {
      uint32_t x = A[k][i], y = B[k][i];
      A[k][i] = sw ? y : x;
      B[k][i] = sw ? x : y;
    }
    double x = fa[k], y = fb[k];
    fa[k] = sw ? y : x;
    fb[k] = sw ? x : y;
  }
  double x = na, y = nb;
  na = sw ? y : x;
  nb = sw ? x : y;
}

// |v| < 2^174 ?  and |v| as 8 limbs
SB_HD bool lat3_abs_fits(const uint32_t* v, uint32_t* mag, bool& neg) {
  neg = (v[LAT3_LIMBS - 1] >> 31) != 0;
  uint32_t c = neg ? 1u : 0u, hi = 0;
#pragma unroll
  for (int i = 0; i < LAT3_LIMBS; i++) {
    uint64_t t = (uint64_t)(neg ? ~v[i] : v[i]) + c;
    uint32_t x = (uint32_t)t;
    c = (uint32_t)(t >> 32);
    if (i < 8) mag[i] = x;
    if (i == 5) hi |= x >> 14;   // bits 174..191
    if (i > 5) hi |= x;
  }
  return hi == 0;
}

// c < 2^252, u < 2^252 (canonical limbs)
SB_HD lat3_res lattice3_8r(const uint32_t* c, const uint32_t* u) {
  const uint32_t n8r[8] = SB200_8R_INIT;
  uint32_t W[3][3][LAT3_LIMBS];  // [vector][coordinate b, a, d][limb]
#pragma unroll
  for (int v = 0; v < 3; v++)
#pragma unroll
    for (int k = 0; k < 3; k++)
#pragma unroll
      for (int i = 0; i < LAT3_LIMBS; i++) W[v][k][i] = 0;
  W[0][0][0] = 1;
#pragma unroll
  for (int i = 0; i < 8; i++) {
    W[0][1][i] = c[i];
    W[0][2][i] = u[i];
    W[1][1][i] = n8r[i];
    W[2][2][i] = n8r[i];
  }
  bool done = false;
#pragma unroll 1
  for (int round = 0; round < LAT3_MAX_ROUNDS; round++) {
    if (!SB_WARP_ANY(!done)) break;
    double F[3][3], nrm[3];
#pragma unroll
    for (int v = 0; v < 3; v++) {
#pragma unroll
      for (int k = 0; k < 3; k++) F[v][k] = lat3_to_double(W[v][k]);
      nrm[v] = lat3_fma(F[v][0], F[v][0], lat3_fma(F[v][1], F[v][1], F[v][2] * F[v][2]));
    }
    // ascending by length
    lat3_cswap(W[0], W[1], F[0], F[1], nrm[0], nrm[1]);
    lat3_cswap(W[1], W[2], F[1], F[2], nrm[1], nrm[2]);
    lat3_cswap(W[0], W[1], F[0], F[1], nrm[0], nrm[1]);
    if (done) continue;  // (the swaps above are idempotent on a finished lane)
    bool changed = false;
    const double lim = 4503599627370496.0;  // 2^52
    // Gauss step: the middle vector modulo the shortest
    double g00 = nrm[0];
    double g01 = lat3_fma(F[0][0], F[1][0], lat3_fma(F[0][1], F[1][1], F[0][2] * F[1][2]));
    if (g00 > 0.0) {
      double q = lat3_rint(g01 / g00);
      q = q > lim ? lim : (q < -lim ? -lim : q);
      if (q != 0.0) {
        changed = true;
#pragma unroll
        for (int k = 0; k < 3; k++) {
          lat3_submul(W[1][k], W[0][k], q);
          F[1][k] = lat3_fma(-q, F[0][k], F[1][k]);
        }
      }
    }
    // the longest vector modulo the plane of the other two (normal equations of the 2 x 2 Gram matrix)
    g01 = lat3_fma(F[0][0], F[1][0], lat3_fma(F[0][1], F[1][1], F[0][2] * F[1][2]));
    const double g11 = lat3_fma(F[1][0], F[1][0], lat3_fma(F[1][1], F[1][1], F[1][2] * F[1][2]));
    const double g02 = lat3_fma(F[0][0], F[2][0], lat3_fma(F[0][1], F[2][1], F[0][2] * F[2][2]));
    const double g12 = lat3_fma(F[1][0], F[2][0], lat3_fma(F[1][1], F[2][1], F[1][2] * F[2][2]));
    const double det = lat3_fma(g00, g11, -(g01 * g01));
    if (det > 0.0) {
      const double inv = 1.0 / det;
      double q0 = lat3_rint(lat3_fma(g02, g11, -(g12 * g01)) * inv);
      double q1 = lat3_rint(lat3_fma(g12, g00, -(g02 * g01)) * inv);
      q0 = q0 > lim ? lim : (q0 < -lim ? -lim : q0);
      q1 = q1 > lim ? lim : (q1 < -lim ? -lim : q1);
      if (q0 != 0.0 || q1 != 0.0) {
        changed = true;
#pragma unroll
        for (int k = 0; k < 3; k++) {
          lat3_submul(W[2][k], W[0][k], q0);
          lat3_submul(W[2][k], W[1][k], q1);
        }
      }
    }
    done = !changed;
  }
  // the shortest basis vector with an odd b that fits the window budget
  lat3_res res;
  res.ok = false;
  res.aneg = res.bneg = res.dneg = false;
#pragma unroll
  for (int i = 0; i < 8; i++) res.a[i] = res.b[i] = res.d[i] = 0;
#pragma unroll
  for (int v = 2; v >= 0; v--) {  // later (shorter) candidates overwrite earlier ones
    uint32_t mb[8], ma[8], md[8];
    bool nb, na, nd;
    bool fits = lat3_abs_fits(W[v][0], mb, nb);
    fits &= lat3_abs_fits(W[v][1], ma, na);
    fits &= lat3_abs_fits(W[v][2], md, nd);
    fits &= (mb[0] & 1u) != 0;
#pragma unroll
    for (int i = 0; i < 8; i++) {
      res.b[i] = fits ? mb[i] : res.b[i];
      res.a[i] = fits ? ma[i] : res.a[i];
      res.d[i] = fits ? md[i] : res.d[i];
    }
    res.bneg = fits ? nb : res.bneg;
    res.aneg = fits ? na : res.aneg;
    res.dneg = fits ? nd : res.dneg;
    res.ok |= fits;
  }
  return res;
}

}  // namespace sb200
