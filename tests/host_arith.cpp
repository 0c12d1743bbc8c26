// Host-side build of the arithmetic headers (the emulated carry-chain blocks), exported with C
// linkage so pytest can drive it through ctypes and compare with the big-int oracle on the CPU.
// Test scaffolding: never linked into libschnorr_b200.so.
#include <cstring>
#include <vector>
#include "../include/schnorr_b200.h"
#include "../schnorr_b200/csrc/wire.cuh"
#include "../schnorr_b200/csrc/params_host.cuh"
using namespace sb200;

static fq L(const uint32_t* p) { fq r; memcpy(r.v, p, 32); return r; }
static void S(uint32_t* p, const fq& a) { memcpy(p, a.v, 32); }

extern "C" {
// Hades tables of this build: derived from the given round constants / MDS exactly as sb200_init_ex derives them
int h_setup_hades(const uint32_t* rc335, const uint32_t* mds25) {
  return params::derive_hades_tables(reinterpret_cast<const uint32_t(*)[8]>(rc335), reinterpret_cast<const uint32_t(*)[5][8]>(mds25), h_hades) ? 0 : -1;
}
void h_counts_reset() { emu::cnt() = {}; }
void h_counts_get(unsigned long long* o) { auto& c = emu::cnt(); o[0] = c.wide; o[1] = c.fq_mul; o[2] = c.fq_sqr; o[3] = c.fq_addsub; o[4] = c.fr_mul; o[5] = c.fq_dot5; o[6] = c.dfma; o[7] = c.fd_mul; o[8] = c.fd_sqr; o[9] = c.fd_dot5; }
void h_fq_mul(const uint32_t* a, const uint32_t* b, uint32_t* r) { S(r, fq_mul(L(a), L(b))); }
void h_fq_sqr(const uint32_t* a, uint32_t* r) { S(r, fq_sqr(L(a))); }
void h_fq_dot5(const uint32_t* c40, const uint32_t* s40, uint32_t* r) {
  S(r, fq_dot5(reinterpret_cast<const uint32_t(*)[8]>(c40), L(s40), L(s40 + 8), L(s40 + 16), L(s40 + 24), L(s40 + 32)));
}
void h_fq_add(const uint32_t* a, const uint32_t* b, uint32_t* r) { S(r, fq_add(L(a), L(b))); }
void h_fq_sub(const uint32_t* a, const uint32_t* b, uint32_t* r) { S(r, fq_sub(L(a), L(b))); }
void h_fq_inv(const uint32_t* a, uint32_t* r) { S(r, fq_inv(L(a))); }
void h_fr_mul(const uint32_t* a, const uint32_t* b, uint32_t* r) {
  fr x, y; memcpy(x.v, a, 32); memcpy(y.v, b, 32); fr z = fr_mul(x, y); memcpy(r, z.v, 32);
}
#if SB_EXPERIMENTAL_FD
// ---- FP64-pipe arithmetic (fd.cuh / hades_fd.cuh), DFMA pair emulated by a 128-bit product --------------------
static fdd FD(const uint32_t* p) { return fd_todbl(fd_split(p)); }   // plain 256-bit integer, limbs re-sliced
static void FS(uint32_t* p, const fd& a) { fd_join(a, p); }           // lazily reduced result (< 2^256)
void h_fd_mul(const uint32_t* a, const uint32_t* b, uint32_t* r) { FS(r, fd_mul(FD(a), FD(b))); }
void h_fd_sqr(const uint32_t* a, uint32_t* r) { FS(r, fd_sqr(FD(a))); }
void h_fd_mulc_add(const uint32_t* cst, const uint32_t* x, const uint32_t* c, uint32_t* r) {
  fdd cd = FD(cst);
  FS(r, fd_mulc_add(cd.d, FD(x), fd_split(c)));
}
void h_fd_dot5(const uint32_t* c40, const uint32_t* s40, const uint32_t* add, uint32_t* r) {
  double cst[25];
  for (int j = 0; j < 5; j++) { fdd t = FD(c40 + 8 * j); for (int k = 0; k < 5; k++) cst[5 * j + k] = t.d[k]; }
  fd ad; if (add) ad = fd_split(add);
  FS(r, fd_dot5(cst, add ? ad.l : nullptr, FD(s40), FD(s40 + 8), FD(s40 + 16), FD(s40 + 24), FD(s40 + 32)));
}
void h_fd_roundtrip(const uint32_t* a, uint32_t* r) { fd_to_canonical(fd_from_fq(L(a)), r); }  // Montgomery-256 in, canonical out
void h_hades_fd(uint32_t* st) {  // Montgomery-256 words in, canonical integers out
  fd s[5]; for (int i = 0; i < 5; i++) s[i] = fd_from_fq(L(st + 8 * i));
  hades_perm_fd(s);
  for (int i = 0; i < 5; i++) fd_to_canonical(s[i], st + 8 * i);
}
#endif
void h_challenge3(const uint32_t* ru, const uint32_t* rv, const uint32_t* m, int fdpath, uint32_t* c) {
#if SB_EXPERIMENTAL_FD
  double slots[FDH_SLOTS * 10 * 3];
  if (fdpath == 2) { challenge3_fd_p(L(ru), L(rv), L(m), c, slots, 1); return; }
  if (fdpath == 3) { challenge3_fd_p(L(ru), L(rv), L(m), c, slots + 1, 3); return; }  // strided layout
  if (fdpath) { challenge3_fd(L(ru), L(rv), L(m), c); return; }
#endif
  challenge3(L(ru), L(rv), L(m), c);
}
#if SB_EXPERIMENTAL_FD
void h_challenge3_pair(const uint32_t* ru, const uint32_t* rv, const uint32_t* m, uint32_t* c) {  // 2 x (8 words) each
  fq a[2] = {L(ru), L(ru + 8)}, b[2] = {L(rv), L(rv + 8)}, mm[2] = {L(m), L(m + 8)};
  challenge3_fd2(a, b, mm, reinterpret_cast<uint32_t(*)[8]>(c));
}
#endif
void h_challenge5(const uint32_t* ru, const uint32_t* rv, const uint32_t* rpu, const uint32_t* rpv, const uint32_t* m, int fdpath, uint32_t* c) {
#if SB_EXPERIMENTAL_FD
  if (fdpath) { challenge5_fd(L(ru), L(rv), L(rpu), L(rpv), L(m), c); return; }
#endif
  challenge5(L(ru), L(rv), L(rpu), L(rpv), L(m), c);
}
void h_hades(uint32_t* st, int dense) {
  fq s[5]; for (int i = 0; i < 5; i++) s[i] = L(st + 8 * i);
  if (dense) hades_perm_dense(s); else hades_perm(s);
  for (int i = 0; i < 5; i++) S(st + 8 * i, s[i]);
}
// comb table for base (u, v): 32*129*24 limbs
void h_comb_build(const uint32_t* bu, const uint32_t* bv, uint32_t* table) {
  for (int j = 0; j < COMB_WINDOWS; j++)
    for (int e = 0; e < COMB_ENTRIES; e++) comb_build_entry(L(bu), L(bv), j, e, table + (size_t)(j * COMB_ENTRIES + e) * 24);
}
static point_in P(const uint32_t* p, int affine) {
  point_in r; r.U = L(p); r.V = L(p + 8); r.affine = affine != 0; r.Z = affine ? fq_one() : L(p + 16); return r;
}
int h_verify(const uint32_t* pk, const uint32_t* u, const uint32_t* R, const uint32_t* m, int affine, const uint32_t* combG, uint32_t* c) {
  return verify_core(P(pk, affine), u, P(R, affine), L(m), combG, c);
}
int h_verify_vargen(const uint32_t* pk, const uint32_t* gen, const uint32_t* u, const uint32_t* R, const uint32_t* m, int affine, uint32_t* c) {
  return verify_vargen_core(P(pk, affine), P(gen, affine), u, P(R, affine), L(m), c);
}
int h_verify_double(const uint32_t* pk, const uint32_t* pkp, const uint32_t* u, const uint32_t* R, const uint32_t* Rp, const uint32_t* m,
                    int affine, const uint32_t* combG, const uint32_t* combGp, uint32_t* c) {
  return verify_double_core(P(pk, affine), P(pkp, affine), u, P(R, affine), P(Rp, affine), L(m), combG, combGp, c);
}
void h_sign(const uint32_t* sk, const uint32_t* nonce, const uint32_t* m, const uint32_t* combG, uint32_t* u, uint32_t* Ruv, uint32_t* c) {
  fq a, b; sign_core(sk, nonce, L(m), combG, u, a, b, c); S(Ruv, a); S(Ruv + 8, b);
}
void h_sign_double(const uint32_t* sk, const uint32_t* nonce, const uint32_t* m, const uint32_t* combG, const uint32_t* combGp,
                   uint32_t* u, uint32_t* Ruv, uint32_t* Rpuv, uint32_t* c) {
  fq a, b, cc, d; sign_double_core(sk, nonce, L(m), combG, combGp, u, a, b, cc, d, c);
  S(Ruv, a); S(Ruv + 8, b); S(Rpuv, cc); S(Rpuv + 8, d);
}
void h_sign_vargen(const uint32_t* sk, const uint32_t* gen, int affine, const uint32_t* nonce, const uint32_t* m, uint32_t* u, uint32_t* Ruv, uint32_t* c) {
  fq a, b; sign_vargen_core(sk, P(gen, affine), nonce, L(m), u, a, b, c); S(Ruv, a); S(Ruv + 8, b);
}
// PLONK witness rows (core.cuh witness_core): scheme 0 / 1 / 2 -> 11 / 19 / 13 field elements
void h_witness(int scheme, const uint32_t* sk, const uint32_t* nonce, const uint32_t* m, const uint32_t* gen, int affine,
               const uint32_t* combG, const uint32_t* combGp, uint32_t* rows) {
  auto emit = [&](int k, const fq& v) { S(rows + 8 * k, v); };
  if (scheme == 0) witness_core<0>(sk, nonce, L(m), point_in(), combG, combGp, emit);
  else if (scheme == 1) witness_core<1>(sk, nonce, L(m), point_in(), combG, combGp, emit);
  else witness_core<2>(sk, nonce, L(m), P(gen, affine), combG, combGp, emit);
}
// address-oblivious scalar multiplication (SB200_SIGN_OBLIVIOUS): 4-bit comb read by masked scan, scanned window table
void h_comb4_build(const uint32_t* bu, const uint32_t* bv, uint32_t* table) {  // 64 x 8 x 24 limbs
  for (int j = 0; j < CT_WINDOWS; j++)
    for (int e = 1; e <= CT_ENTRIES; e++) comb4_build_entry(L(bu), L(bv), j, e, table + (size_t)(j * CT_ENTRIES + e - 1) * 24);
}
void h_fixed_mul_oblivious(const uint32_t* comb4, const uint32_t* k, uint32_t* uv) {
  fq a, b; ext_to_affine(fixed_base_mul_oblivious(comb4, k), a, b); S(uv, a); S(uv + 8, b);
}
void h_var_mul_oblivious(const uint32_t* p, int affine, const uint32_t* k, uint32_t* uv) {
  fq a, b; ext_to_affine(var_base_mul_oblivious(P(p, affine), k), a, b); S(uv, a); S(uv + 8, b);
}
void h_fq_inv_fast(const uint32_t* a, uint32_t* r) { S(r, fq_inv_fast(L(a))); }
void h_fq_inv_fermat(const uint32_t* a, uint32_t* r) { S(r, fq_inv(L(a))); }
// 3-dimensional short vector for the variable-generator verification: out = a[8] | b[8] | d[8] | aneg bneg dneg ok
void h_lattice3(const uint32_t* c, const uint32_t* u, uint32_t* out) {
  lat3_res r = lattice3_8r(c, u);
  memcpy(out, r.a, 32); memcpy(out + 8, r.b, 32); memcpy(out + 16, r.d, 32);
  out[24] = r.aneg; out[25] = r.bneg; out[26] = r.dneg; out[27] = r.ok;
}
int h_verify_vargen_ec(const uint32_t* pk, const uint32_t* gen, const uint32_t* u, const uint32_t* R, const uint32_t* c, int affine, int fast, int* fast_ok) {
  bool fo = true;
  bool ok = fast ? verify_vargen_ec_fast(P(pk, affine), P(gen, affine), u, P(R, affine), c, fo) : verify_vargen_ec(P(pk, affine), P(gen, affine), u, P(R, affine), c);
  *fast_ok = fo;
  return ok;
}
int h_point_well_formed(const uint32_t* p, int affine) { return point_well_formed(P(p, affine)); }
int h_decompress(const uint32_t* b, uint32_t* uv) { fq u, v; bool ok = point_decompress(b, u, v); S(uv, u); S(uv + 8, v); return ok; }
void h_compress(const uint32_t* uv, uint32_t* b) { point_compress(L(uv), L(uv + 8), b); }
void h_fr_from_wide(const uint32_t* w, uint32_t* r) { fr_from_wide(w, r); }
void h_fq_from_wide(const uint32_t* w, uint32_t* r) { S(r, fq_from_wide(w)); }
int h_fq_sqrt_ratio(const uint32_t* n, const uint32_t* d, uint32_t* r) { fq x; bool ok = fq_sqrt_ratio(L(n), L(d), x); S(r, x); return ok; }
int h_fq_sqrt(const uint32_t* a, uint32_t* r) { fq x; bool ok = fq_sqrt(L(a), x); S(r, x); return ok; }
int h_verify_bytes(const uint32_t* pk, const uint32_t* sig, const uint32_t* msg, const uint32_t* combG, int* invalid) {
  bool inv; bool ok = verify_bytes_core(pk, sig, msg, combG, inv); *invalid = inv; return ok;
}
int h_verify_double_bytes(const uint32_t* pk, const uint32_t* sig, const uint32_t* msg, const uint32_t* combG, const uint32_t* combGp, int* invalid) {
  bool inv; bool ok = verify_double_bytes_core(pk, sig, msg, combG, combGp, inv); *invalid = inv; return ok;
}
int h_verify_vargen_bytes(const uint32_t* pk, const uint32_t* sig, const uint32_t* msg, int* invalid) {
  bool inv; bool ok = verify_vargen_bytes_core(pk, sig, msg, inv); *invalid = inv; return ok;
}
void h_sign_double_bytes(const uint32_t* sk, const uint32_t* msg, const uint32_t* nonce, const uint32_t* combG, const uint32_t* combGp, uint32_t* sig96) {
  sign_double_bytes_core(sk, msg, nonce, combG, combGp, sig96);
}
int h_sign_vargen_bytes(const uint32_t* sk64, const uint32_t* msg, const uint32_t* nonce, uint32_t* sig64) { return sign_vargen_bytes_core(sk64, msg, nonce, sig64); }
// half-size scalars: out = a[8] | b[8] | bneg | ok
void h_half_gcd(const uint32_t* c, uint32_t* out) {
  hgcd_res r = half_gcd_8r(c);
  memcpy(out, r.a, 32); memcpy(out + 8, r.b, 32); out[16] = r.bneg; out[17] = r.ok;
}
int h_verify_ec(const uint32_t* pk, const uint32_t* u, const uint32_t* R, const uint32_t* c, int affine, const uint32_t* combG, int fast, int* fast_ok) {
  bool fo = true;
  bool ok = fast ? verify_ec_core_fast(P(pk, affine), u, P(R, affine), c, combG, fo) : verify_ec_core(P(pk, affine), u, P(R, affine), c, combG);
  *fast_ok = fo;
  return ok;
}
void h_fixed_mul(const uint32_t* comb, const uint32_t* k, uint32_t* uv) {
  fq a, b; ext_to_affine(fixed_base_mul(comb, k), a, b); S(uv, a); S(uv + 8, b);
}
void h_var_mul(const uint32_t* p, int affine, const uint32_t* k, uint32_t* uv) {
  fq a, b; ext_to_affine(var_base_mul(P(p, affine), k), a, b); S(uv, a); S(uv + 8, b);
}
}
