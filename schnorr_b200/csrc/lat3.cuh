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
// The reduction is a greedy (Semaev-style) one, shaped for one tuple per thread with all lanes of a warp in lock-step,
// in two levels like Lehmer's gcd.  INNER rounds work on a double-precision copy F of the three basis vectors and on the
// 3 x 3 integer transform T accumulated so far (also held in doubles, exactly): sort the vectors by length, take the
// middle one modulo the shortest (a Gauss step) and the longest modulo the plane of the other two, quotients rounded
// from the Gram matrix of F.  When T's entries reach 2^20 (F has then lost ~40 of its 53 bits to cancellation) or
// nothing changes, the OUTER level applies T EXACTLY to the 288-bit integer vectors and refreshes F from them.
// Floating-point error can only make a step sub-optimal, never wrong: T is unimodular by construction and is applied
// with exact integer arithmetic, so the three vectors are a basis of the lattice at all times.  ~70 inner rounds in
// 9-10 outer passes (measured over random inputs: 50-88 / 9-10) bring the entries from 2^255 down to <= 2^173; the
// shortest basis vector with an odd b is used (the b's of a basis cannot all be even: (1, c, u) is in the lattice).
// If none fits the 174-bit window budget -- structured inputs such as u = r - 1, where (8, ., -8) is in the lattice
// and every short vector has an even b -- `ok` is false and the caller runs the full-size multiplication for that
// tuple.  (A one-level version that updated the 288-bit vectors every round was measured first: its ~1 300
// instructions per round ate the whole gain, 15.8 -> 16.3 M/s)
#pragma once
#include "fq.cuh"
#include "hgcd.cuh"

#if !defined(__CUDA_ARCH__)
#include <cmath>
#endif

namespace sb200 {

constexpr int LAT3_LIMBS = 9;        // 288-bit two's complement
constexpr int LAT3_WINDOWS = 44;     // 4-bit windows of the triple-scalar multiplication: magnitudes < 2^174
constexpr int LAT3_MAX_OUTER = 20;    // exact refreshes (measured 9-10)
constexpr int LAT3_MAX_INNER = 32;    // double-precision rounds between two refreshes

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

SB_HD void lat3_cswap(double* fa, double* fb, double* ta, double* tb, double& na, double& nb) {
  const bool sw = na > nb;
#pragma unroll
  for (int k = 0; k < 3; k++) {
    double x = fa[k], y = fb[k];
    fa[k] = sw ? y : x;
    fb[k] = sw ? x : y;
    x = ta[k]; y = tb[k];
    ta[k] = sw ? y : x;
    tb[k] = sw ? x : y;
  }
  double x = na, y = nb;
  na = sw ? y : x;
  nb = sw ? x : y;
}
SB_HD double lat3_dot(const double* a, const double* b) { return lat3_fma(a[0], b[0], lat3_fma(a[1], b[1], a[2] * b[2])); }
SB_HD double lat3_clamp(double q, double lim) { return q > lim ? lim : (q < -lim ? -lim : q); }

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
  // [vector][coordinate b, a, d][limb]; indexed by loop variables below, so it lives in (L1-resident) local memory:
  // every lane uses the same index, the accesses are coalesced, and 81 registers stay free
  uint32_t W[3][3][LAT3_LIMBS];
#pragma unroll 1
  for (int v = 0; v < 3; v++)
#pragma unroll 1
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
  for (int outer = 0; outer < LAT3_MAX_OUTER; outer++) {
    if (!SB_WARP_ANY(!done)) break;
    double F[3][3], T[3][3], nrm[3];
#pragma unroll
    for (int v = 0; v < 3; v++) {
#pragma unroll
      for (int k = 0; k < 3; k++) {
        F[v][k] = lat3_to_double(W[v][k]);
        T[v][k] = v == k ? 1.0 : 0.0;
      }
    }
    bool any_change = false, stop = done;
#pragma unroll 1
    for (int inner = 0; inner < LAT3_MAX_INNER; inner++) {
      if (!SB_WARP_ANY(!stop)) break;
      if (stop) continue;
#pragma unroll
      for (int v = 0; v < 3; v++) nrm[v] = lat3_dot(F[v], F[v]);
      lat3_cswap(F[0], F[1], T[0], T[1], nrm[0], nrm[1]);  // ascending by length
      lat3_cswap(F[1], F[2], T[1], T[2], nrm[1], nrm[2]);
      lat3_cswap(F[0], F[1], T[0], T[1], nrm[0], nrm[1]);
      bool changed = false;
      // quotients are clamped to 2^15 (a clamped quotient is a partial, still valid, step): with |T| < 2^20 before a
      // round, |T| <= 2^20 (1 + 2 q + q^2) < 2^51 after it, so T stays an exact integer matrix in doubles and its entries
      // fit lat3_submul's 52-bit multiplier
      const double qlim = 32768.0, tlim = 1048576.0;
      const double g00 = nrm[0];
      if (g00 > 0.0) {  // Gauss step: the middle vector modulo the shortest
        const double q = lat3_clamp(lat3_rint(lat3_dot(F[0], F[1]) / g00), qlim);
        if (q != 0.0) {
          changed = true;
#pragma unroll
          for (int k = 0; k < 3; k++) {
            F[1][k] = lat3_fma(-q, F[0][k], F[1][k]);
            T[1][k] = lat3_fma(-q, T[0][k], T[1][k]);
          }
        }
      }
      // the longest vector modulo the plane of the other two (normal equations of the 2 x 2 Gram matrix)
      const double g01 = lat3_dot(F[0], F[1]), g11 = lat3_dot(F[1], F[1]);
      const double g02 = lat3_dot(F[0], F[2]), g12 = lat3_dot(F[1], F[2]);
      const double det = lat3_fma(g00, g11, -(g01 * g01));
      if (det > 0.0) {
        const double inv = 1.0 / det;
        const double q0 = lat3_clamp(lat3_rint(lat3_fma(g02, g11, -(g12 * g01)) * inv), qlim);
        const double q1 = lat3_clamp(lat3_rint(lat3_fma(g12, g00, -(g02 * g01)) * inv), qlim);
        if (q0 != 0.0 || q1 != 0.0) {
          changed = true;
#pragma unroll
          for (int k = 0; k < 3; k++) {
            F[2][k] = lat3_fma(-q1, F[1][k], lat3_fma(-q0, F[0][k], F[2][k]));
            T[2][k] = lat3_fma(-q1, T[1][k], lat3_fma(-q0, T[0][k], T[2][k]));
          }
        }
      }
      any_change |= changed;
      double tmax = 0.0;
#pragma unroll
      for (int v = 0; v < 3; v++)
#pragma unroll
        for (int k = 0; k < 3; k++) tmax = fmax(tmax, fabs(T[v][k]));
      stop = !changed || tmax >= tlim;
    }
    // W <- T W, exactly, one coordinate at a time (lanes that are done carry T = identity)
    double Tm[3][3];
#pragma unroll
    for (int v = 0; v < 3; v++)
#pragma unroll
      for (int k = 0; k < 3; k++) Tm[v][k] = done ? (v == k ? 1.0 : 0.0) : T[v][k];
#pragma unroll 1
    for (int k = 0; k < 3; k++) {
      uint32_t nw[3][LAT3_LIMBS];
#pragma unroll
      for (int v = 0; v < 3; v++) {
#pragma unroll
        for (int i = 0; i < LAT3_LIMBS; i++) nw[v][i] = 0;
#pragma unroll
        for (int j = 0; j < 3; j++) lat3_submul(nw[v], W[j][k], -Tm[v][j]);
      }
#pragma unroll
      for (int v = 0; v < 3; v++)
#pragma unroll
        for (int i = 0; i < LAT3_LIMBS; i++) W[v][k][i] = nw[v][i];
    }
    done |= !any_change;
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

#ifndef SB_LAT3_OOL
#define SB_LAT3_OOL 1  // inlined into the three-table curve kernel the reduction was miscompiled (wrong vectors on the GPU, right ones in its own kernel and on the host): tests/test_gpu_robustness.py::test_lattice3_on_the_device, test_vargen_*
#endif
#if defined(__CUDA_ARCH__) && SB_LAT3_OOL
// out of line, arguments and result through local memory: the reduction keeps its own register allocation
static __device__ __noinline__ void lattice3_8r_ool(const uint32_t* c, const uint32_t* u, lat3_res* r) { *r = lattice3_8r(c, u); }
SB_HD lat3_res lattice3_8r_call(const uint32_t* c, const uint32_t* u) {
  lat3_res r;
  lattice3_8r_ool(c, u, &r);
  return r;
}
#else
SB_HD lat3_res lattice3_8r_call(const uint32_t* c, const uint32_t* u) { return lattice3_8r(c, u); }
#endif

}  // namespace sb200
