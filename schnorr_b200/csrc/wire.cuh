// Wire formats on the device: the reference's serialised forms are accepted / produced
// directly, so a caller holding bytes never decodes on the host.
//
//   JubJubAffine::to_bytes / from_bytes   (32 B: v little-endian, bit 255 = low bit of u; no subgroup check)
//       as used by Signature::{to,from}_bytes   /root/reference/src/signatures.rs:106-123
//       and PublicKey::{to,from}_bytes          /root/reference/src/keys/public.rs:87-101
//   JubJubScalar::from_bytes              (32 B canonical, rejects >= r)     /root/reference/src/keys/secret.rs:96-102
//   BlsScalar::from_bytes                 (32 B canonical, rejects >= q)
//   Field::random = from_bytes_wide       (64 B little-endian reduced mod r / mod q)  /root/reference/src/keys/secret.rs:83,155
#pragma once
#include "core.cuh"

namespace sb200 {

SB_HD bool lt_q(const uint32_t* k) {
  uint32_t t[8];
  const uint32_t qq[8] = SB200_FQ_MOD_INIT;
  return sub8(t, k, qq) != 0;
}

// x^e for a constant little-endian exponent of `bits` bits, 4-bit fixed windows
SB_HD fq fq_pow_const(const fq& x, const uint32_t* e, int bits) {
  fq tab[16];
  tab[0] = fq_one();
  tab[1] = x;
#pragma unroll 1
  for (int i = 2; i < 16; i++) tab[i] = fq_mul(tab[i - 1], x);
  fq acc = fq_one();
#pragma unroll 1
  for (int w = (bits + 3) / 4 - 1; w >= 0; w--) {
    acc = fq_sqr(fq_sqr(fq_sqr(fq_sqr(acc))));
    acc = fq_mul(acc, tab[(e[w >> 3] >> ((w & 7) * 4)) & 15u]);
  }
  return acc;
}

// Tonelli-Shanks square root in F_q (2-adicity 32).  Returns false if x is a non-residue.  Variable time
// (inputs are public: keys and signatures).
SB_HD bool fq_sqrt(const fq& x, fq& root) {
  const uint32_t e[8] = SB200_FQ_SQRT_EXP_INIT;
  const fq g = {SB200_FQ_ROOT_OF_UNITY_INIT};
  const fq one = fq_one();
  fq w = fq_pow_const(x, e, SB200_FQ_SQRT_EXP_BITS);  // x^((t-1)/2)
  fq a = fq_mul(x, w);                                 // x^((t+1)/2)
  fq b = fq_mul(a, w);                                 // x^t, in the 2^32-order subgroup
  fq z = g;
  int v = 32;
  bool ok = true;
  if (fq_is_zero(x)) {
    root = x;
    return true;
  }
#pragma unroll 1
  while (!fq_eq(b, one)) {
    int k = 0;
    fq tmp = b;
#pragma unroll 1
    while (!fq_eq(tmp, one) && k < v) {
      tmp = fq_sqr(tmp);
      k++;
    }
    if (k >= v) {  // order of b does not divide 2^(v-1): non-residue
      ok = false;
      break;
    }
    fq ww = z;
#pragma unroll 1
    for (int i = 0; i < v - k - 1; i++) ww = fq_sqr(ww);
    a = fq_mul(a, ww);
    z = fq_sqr(ww);
    b = fq_mul(b, z);
    v = k;
  }
  root = a;
  return ok;
}

// sqrt(num / den) without an inversion and without the Tonelli-Shanks search loops (den != 0):
//   y = num den,  a = y^(-(t+1)/2) (one 255-bit exponentiation),  b = a^2 y = y^-t lies in the subgroup of order 2^32 = <g>;
//   the discrete logarithm k of b (b = g^k) is read off in eight 4-bit digits, Pohlig-Hellman style -- digit i is
//   the position of (b g^-(k mod 16^i))^(2^(28-4i)) among the 16 sixteenth roots of unity W[] -- and then
//   sqrt(1/y) = a g^(-k/2), sqrt(num/den) = num sqrt(1/y).  112 squarings + 16 multiplications replace ~500 squarings
//   of the bit-serial search (whose trip counts also diverge across the lanes of a warp), and the separate field
//   inversion (255 squarings) disappears: 2.6x fewer multiplications per decompressed point.
// Branch-free; a non-residue (odd k) or garbage input yields a wrong candidate that the closing check u^2 den == num
// rejects, so the function is self-verifying.  Returns the candidate in `u` (either root; the caller fixes the sign).
#ifndef SB_FAST_DECOMPRESS
#define SB_FAST_DECOMPRESS 1
#endif
#if defined(__CUDACC__)
__device__ const uint32_t d_fq_dlog_w[16][8] = SB200_FQ_DLOG_W_INIT;
__device__ const uint32_t d_fq_dlog_p[128][8] = SB200_FQ_DLOG_P_INIT;
__device__ const uint32_t d_fq_dlog_q[128][8] = SB200_FQ_DLOG_Q_INIT;
#endif
static const uint32_t h_fq_dlog_w[16][8] = SB200_FQ_DLOG_W_INIT;
static const uint32_t h_fq_dlog_p[128][8] = SB200_FQ_DLOG_P_INIT;
static const uint32_t h_fq_dlog_q[128][8] = SB200_FQ_DLOG_Q_INIT;

SB_HD bool fq_sqrt_ratio(const fq& num, const fq& den, fq& u) {
  const uint32_t e[8] = SB200_FQ_ISQRT_EXP_INIT;
  fq y = fq_mul(num, den);
  fq a = fq_pow_const(y, e, SB200_FQ_ISQRT_EXP_BITS);  // y^(-(t+1)/2)
  fq b = fq_mul(fq_sqr(a), y);                          // y^-t
#pragma unroll 1
  for (int i = 0; i < 8; i++) {
    fq h = b;
#pragma unroll 1
    for (int s = 0; s < 28 - 4 * i; s++) h = fq_sqr(h);
    int d = 0;
#pragma unroll 1
    for (int j = 1; j < 16; j++) d = fq_eq(h, ld8(SB_CONST(fq_dlog_w)[j])) ? j : d;
    b = fq_mul(b, ld8(SB_CONST(fq_dlog_p)[16 * i + d]));
    a = fq_mul(a, ld8(SB_CONST(fq_dlog_q)[16 * i + d]));
  }
  u = fq_mul(num, a);
  return fq_eq(fq_mul(fq_sqr(u), den), num);
}

// 32 bytes (as 8 LE words) -> affine point, Montgomery.  false <=> the reference's from_bytes returns None.
SB_HD bool point_decompress(const uint32_t* in, fq& u, fq& v) {
  uint32_t vb[8];
#pragma unroll
  for (int i = 0; i < 8; i++) vb[i] = in[i];
  uint32_t sign = vb[7] >> 31;
  vb[7] &= 0x7fffffffu;
  bool ok = lt_q(vb);
  fq vc;
#pragma unroll
  for (int i = 0; i < 8; i++) vc.v[i] = ok ? vb[i] : 0u;
  v = fq_to_mont(vc);
  fq v2 = fq_sqr(v);
  fq num = fq_sub(v2, fq_one());
  fq den = fq_add(fq_one(), fq_mul(ed_d(), v2));
#if SB_FAST_DECOMPRESS
  ok &= fq_sqrt_ratio(num, den, u);  // den = 1 + d v^2 != 0: d is a non-square, -1 a square
#else
  fq x = fq_mul(num, fq_inv(den));  // den = 0 -> inverse "0" -> x = 0, as invert().unwrap_or(zero)
  ok &= fq_sqrt(x, u);
#endif
  uint32_t parity = fq_from_mont(u).v[0] & 1u;
  u = fq_select(u, fq_neg(u), parity != sign);
  return ok;
}

// affine (u, v), Montgomery -> 32 bytes
SB_HD void point_compress(const fq& u, const fq& v, uint32_t* out) {
  fq vc = fq_from_mont(v);
  uint32_t parity = fq_from_mont(u).v[0] & 1u;
#pragma unroll
  for (int i = 0; i < 8; i++) out[i] = vc.v[i];
  out[7] |= parity << 31;
}

// 512-bit little-endian integer (16 words) reduced mod r -> canonical scalar.
// lo mod r = mont(mont(lo, R^2), 1);  hi * 2^256 mod r = mont(hi, R^2);  (mont(a, b) needs only b < r)
SB_HD void fr_from_wide(const uint32_t* w, uint32_t* out) {
  fr lo, hi, one, r2 = {SB200_FR_R2_INIT};
#pragma unroll
  for (int i = 0; i < 8; i++) {
    lo.v[i] = w[i];
    hi.v[i] = w[8 + i];
    one.v[i] = i == 0 ? 1u : 0u;
  }
  fr a = fr_mont_mul(fr_mont_mul(lo, r2), one);
  fr b = fr_mont_mul(hi, r2);
  fr s;
  add8(s.v, a.v, b.v);  // < 2r < 2^256
  cond_sub_p<FrP>(s.v);
#pragma unroll
  for (int i = 0; i < 8; i++) out[i] = s.v[i];
}
// same mod q, result in Montgomery form (what BlsScalar holds):  lo*R + hi*R^2 = mont(lo, R^2) + mont(mont(hi, R^2), R^2)
SB_HD fq fq_from_wide(const uint32_t* w) {
  fq lo, hi;
  const fq r2 = {SB200_FQ_R2_INIT};
#pragma unroll
  for (int i = 0; i < 8; i++) {
    lo.v[i] = w[i];
    hi.v[i] = w[8 + i];
  }
  // reduce the raw halves below q first (they may be >= q: up to 2 subtractions since 2^256 < 3q)
  cond_sub_p<FqP>(lo.v); cond_sub_p<FqP>(lo.v);
  cond_sub_p<FqP>(hi.v); cond_sub_p<FqP>(hi.v);
  return fq_add(fq_mul(lo, r2), fq_mul(fq_mul(hi, r2), r2));
}

// PublicKey::from_bytes + Signature::from_bytes + BlsScalar::from_bytes + verify, one tuple.
// `invalid` <=> any from_bytes of the reference would fail (InvalidData); then the verdict is false.
SB_HD bool verify_bytes_core(const uint32_t* pk32, const uint32_t* sig64, const uint32_t* msg32, const uint32_t* combG,
                             bool& invalid) {
  point_in PK, R;
  PK.affine = R.affine = true;
  PK.Z = R.Z = fq_one();
  bool ok = point_decompress(pk32, PK.U, PK.V);
  ok &= point_decompress(sig64 + 8, R.U, R.V);
  uint32_t u[8], mb[8], c[8];
#pragma unroll
  for (int i = 0; i < 8; i++) {
    u[i] = sig64[i];
    mb[i] = msg32[i];
  }
  ok &= scalar_lt_r(u) & lt_q(mb);
  fq mc;
#pragma unroll
  for (int i = 0; i < 8; i++) mc.v[i] = ok ? mb[i] : 0u;
  invalid = !ok;
  bool verdict = verify_core(PK, u, R, fq_to_mont(mc), combG, c);
  return ok & verdict;
}

// canonical message bytes -> Montgomery field element; false <=> BlsScalar::from_bytes fails (>= q)
SB_HD bool msg_from_bytes(const uint32_t* msg32, fq& m) {
  uint32_t mb[8];
#pragma unroll
  for (int i = 0; i < 8; i++) mb[i] = msg32[i];
  bool ok = lt_q(mb);
  fq mc;
#pragma unroll
  for (int i = 0; i < 8; i++) mc.v[i] = ok ? mb[i] : 0u;
  m = fq_to_mont(mc);
  return ok;
}
SB_HD bool point_from_bytes(const uint32_t* b32, point_in& P) {
  P.affine = true;
  P.Z = fq_one();
  return point_decompress(b32, P.U, P.V);
}

// PublicKeyDouble::from_bytes (pk || pk', /root/reference/src/keys/public.rs:282-299) +
// SignatureDouble::from_bytes (u || R || R', /root/reference/src/signatures.rs:245-270) + verify (public.rs:222-244)
SB_HD bool verify_double_bytes_core(const uint32_t* pk64, const uint32_t* sig96, const uint32_t* msg32, const uint32_t* combG,
                                    const uint32_t* combGp, bool& invalid) {
  point_in PK, PKp, R, Rp;
  bool ok = point_from_bytes(pk64, PK);
  ok &= point_from_bytes(pk64 + 8, PKp);
  ok &= point_from_bytes(sig96 + 8, R);
  ok &= point_from_bytes(sig96 + 16, Rp);
  uint32_t u[8], c[8];
#pragma unroll
  for (int i = 0; i < 8; i++) u[i] = sig96[i];
  fq m;
  ok &= scalar_lt_r(u) & msg_from_bytes(msg32, m);
  invalid = !ok;
  bool verdict = verify_double_core(PK, PKp, u, R, Rp, m, combG, combGp, c);
  return ok & verdict;
}

// PublicKeyVarGen::from_bytes (pk || generator, public.rs:347-372) + SignatureVarGen::from_bytes (u || R,
// signatures.rs:387-404) + verify (public.rs:401-415)
SB_HD bool verify_vargen_bytes_core(const uint32_t* pk64, const uint32_t* sig64, const uint32_t* msg32, bool& invalid) {
  point_in PK, GEN, R;
  bool ok = point_from_bytes(pk64, PK);
  ok &= point_from_bytes(pk64 + 8, GEN);
  ok &= point_from_bytes(sig64 + 8, R);
  uint32_t u[8], c[8];
#pragma unroll
  for (int i = 0; i < 8; i++) u[i] = sig64[i];
  fq m;
  ok &= scalar_lt_r(u) & msg_from_bytes(msg32, m);
  invalid = !ok;
  bool verdict = verify_vargen_core(PK, GEN, u, R, m, c);
  return ok & verdict;
}

// The from_bytes of a signing tuple: SecretKey::from_bytes / JubJubScalar::from_bytes (reject >= r,
// /root/reference/src/keys/secret.rs:96-102) for sk and the nonce, BlsScalar::from_bytes (reject >= q) for the message.
// Returns false where the reference would return Err(InvalidData); the values are then replaced by 0 so the
// arithmetic below stays in range (a nonce >= 2^252 would overflow the window recoding), and the caller emits no
// signature for that tuple.
SB_HD bool sign_inputs_from_bytes(const uint32_t* sk32, const uint32_t* nonce32, const uint32_t* msg32, uint32_t* sk,
                                  uint32_t* nonce, fq& m) {
  uint32_t a[8], b[8];
#pragma unroll
  for (int i = 0; i < 8; i++) {
    a[i] = sk32[i];
    b[i] = nonce32[i];
  }
  bool ok = scalar_lt_r(a) & scalar_lt_r(b);
  ok &= msg_from_bytes(msg32, m);
#pragma unroll
  for (int i = 0; i < 8; i++) {
    sk[i] = ok ? a[i] : 0u;
    nonce[i] = ok ? b[i] : 0u;
  }
  if (!ok) m = fq_zero();
  return ok;
}

// SecretKey::sign_double -> SignatureDouble::to_bytes (secret.rs:217-240, signatures.rs:245-258): sig96 = u || R || R'.
// false <=> a from_bytes fails; sig96 is then all zero.
SB_HD bool sign_double_bytes_core(const uint32_t* sk32, const uint32_t* msg32, const uint32_t* nonce32, const uint32_t* combG,
                                  const uint32_t* combGp, uint32_t* sig96) {
  fq m, Ru, Rv, Rpu, Rpv;
  uint32_t c[8], sk[8], nonce[8];
  bool ok = sign_inputs_from_bytes(sk32, nonce32, msg32, sk, nonce, m);
  sign_double_core(sk, nonce, m, combG, combGp, sig96, Ru, Rv, Rpu, Rpv, c);
  point_compress(Ru, Rv, sig96 + 8);
  point_compress(Rpu, Rpv, sig96 + 16);
#pragma unroll
  for (int i = 0; i < 24; i++) sig96[i] = ok ? sig96[i] : 0u;
  return ok;
}

// SecretKeyVarGen::from_bytes (sk || generator, secret.rs:313-336) + sign (secret.rs:433-451) +
// SignatureVarGen::to_bytes: sig64 = u || R.  false <=> a from_bytes fails (sk, nonce, message, or the generator does
// not decode); sig64 is then all zero.
SB_HD bool sign_vargen_bytes_core(const uint32_t* sk64, const uint32_t* msg32, const uint32_t* nonce32, uint32_t* sig64) {
  point_in GEN;
  bool ok = point_from_bytes(sk64 + 8, GEN);
  fq m, Ru, Rv;
  uint32_t c[8], sk[8], nonce[8];
  ok &= sign_inputs_from_bytes(sk64, nonce32, msg32, sk, nonce, m);
  sign_vargen_core(sk, GEN, nonce, m, sig64, Ru, Rv, c);
  point_compress(Ru, Rv, sig64 + 8);
#pragma unroll
  for (int i = 0; i < 16; i++) sig64[i] = ok ? sig64[i] : 0u;
  return ok;
}

}  // namespace sb200
