// Hades252 permutation (width 5, x^5, 8 full + 59 partial rounds) and the Poseidon sponge shapes the
// Schnorr challenge uses.  Replaces `dusk_poseidon::sponge::truncated::hash` as called from
// challenge_hash / challenge_hash_double: /root/reference/src/signatures.rs:127-134, 275-290.
//
// One permutation per thread, state in registers (5 x 8 limbs), constants broadcast from
// __constant__ memory (every lane reads the same address in lock-step).
// The 59 partial rounds run in the sparse factorisation derived at context creation (params_host.cuh;
// 9 products per round instead of 25; algebraically identical, hence bit-exact); the dense form is
// kept as hades_perm_dense for the parity tests and the self-check of the derivation.
#pragma once
#include "fq.cuh"

namespace sb200 {

#define SB_HADES_W 5

// Every Hades table is an INPUT of context creation (include/schnorr_b200.h sb200_params -> params_host.cuh
// derive_hades_tables), uploaded once per device; nothing numeric from dusk-hades is baked into the kernels.
struct HadesTables {
  uint32_t rc[335][8];         // ROUND_CONSTANTS[0..335): 67 rounds x 5 words, consumed in order
  uint32_t mds[25][8];         // MDS_MATRIX, row-major
  uint32_t pre[5][8];          // sparse form: added to the state before the partial rounds
  uint32_t sparse[59 * 11][8]; // per partial round: key (1), row (5), column (5; entry 4 unused)
  uint32_t post[25][8];        // dense matrix applied once after the partial rounds
};
#if defined(__CUDACC__)
__constant__ HadesTables d_hades;
#endif
static HadesTables h_hades;  // host build (parameter derivation self-check, tests/host_arith): filled by the same derivation

#if defined(__CUDA_ARCH__)
#define SB_CONST(name) d_##name
#define SB_HADES(field) d_hades.field
#else
#define SB_CONST(name) h_##name
#define SB_HADES(field) h_hades.field
#endif

SB_HD fq ld8(const uint32_t* p) {
  fq r;
#pragma unroll
  for (int i = 0; i < 8; i++) r.v[i] = p[i];
  return r;
}

SB_HD fq fq_pow5(const fq& x) {
  fq x2 = fq_sqr(x);
  fq x4 = fq_sqr(x2);
  return fq_mul(x4, x);
}

// s <- M * s with M a dense 5x5 matrix in constant memory (row-major)
// (each row is one lazily reduced dot product: 5 x 64 + 56 wide products instead of 5 x 120)
#define SB_MATMUL5(s, mat)                                                                  \
  do {                                                                                      \
    fq _r[5];                                                                               \
    _Pragma("unroll") for (int _k = 0; _k < 5; _k++)                                        \
        _r[_k] = fq_dot5(&SB_HADES(mat)[_k * 5], s[0], s[1], s[2], s[3], s[4]);             \
    _Pragma("unroll") for (int _k = 0; _k < 5; _k++) s[_k] = _r[_k];                        \
  } while (0)

// plain version (one reduction per product) for the reference-shaped dense permutation
#define SB_MATMUL5_PLAIN(s, mat)                                        \
  do {                                                                  \
    fq _r[5];                                                           \
    _Pragma("unroll 1") for (int _k = 0; _k < 5; _k++) {                \
      fq _a = fq_mul(ld8(SB_HADES(mat)[_k * 5]), s[0]);                 \
      _Pragma("unroll") for (int _j = 1; _j < 5; _j++)                  \
          _a = fq_add(_a, fq_mul(ld8(SB_HADES(mat)[_k * 5 + _j]), s[_j])); \
      _r[_k] = _a;                                                      \
    }                                                                   \
    _Pragma("unroll") for (int _k = 0; _k < 5; _k++) s[_k] = _r[_k];    \
  } while (0)

template <bool LAZY = true>
SB_HD void hades_full_round(fq* s, int rc_base) {
#pragma unroll
  for (int k = 0; k < 5; k++) s[k] = fq_pow5(fq_add(s[k], ld8(SB_HADES(rc)[rc_base + k])));
  if (LAZY) SB_MATMUL5(s, mds); else SB_MATMUL5_PLAIN(s, mds);
}

// Reference-shaped permutation (ScalarStrategy::perm of dusk-hades): used by parity tests.
SB_HD void hades_perm_dense(fq* s) {
  int rc = 0;
#pragma unroll 1
  for (int r = 0; r < 4; r++, rc += 5) hades_full_round<false>(s, rc);
#pragma unroll 1
  for (int r = 0; r < 59; r++, rc += 5) {
#pragma unroll
    for (int k = 0; k < 5; k++) s[k] = fq_add(s[k], ld8(SB_HADES(rc)[rc + k]));
    s[4] = fq_pow5(s[4]);
    SB_MATMUL5_PLAIN(s, mds);
  }
#pragma unroll 1
  for (int r = 0; r < 4; r++, rc += 5) hades_full_round<false>(s, rc);
}

// Production permutation: sparse partial rounds.
SB_HD void hades_perm(fq* s) {
  int rc = 0;
#pragma unroll 1
  for (int r = 0; r < 4; r++, rc += 5) hades_full_round(s, rc);
#pragma unroll
  for (int k = 0; k < 4; k++) s[k] = fq_add(s[k], ld8(SB_HADES(pre)[k]));
#pragma unroll 1
  for (int t = 0; t < 59; t++) {
    const uint32_t(*c)[8] = &SB_HADES(sparse)[t * 11];
    s[4] = fq_pow5(fq_add(s[4], ld8(c[0])));
    fq np = fq_dot5(&c[1], s[0], s[1], s[2], s[3], s[4]);  // row . state, one reduction
#pragma unroll
    for (int j = 0; j < 4; j++) s[j] = fq_add(s[j], fq_mul(ld8(c[6 + j]), s[4]));
    s[4] = np;
  }
  SB_MATMUL5(s, post);
  rc += 59 * 5;
#pragma unroll 1
  for (int r = 0; r < 4; r++, rc += 5) hades_full_round(s, rc);
}

// low 250 bits of the canonical integer (dusk_poseidon::sponge::truncated), as a scalar
SB_HD void truncate250(const fq& mont_word, uint32_t* c) {
  fq x = fq_from_mont(mont_word);
#pragma unroll
  for (int i = 0; i < 8; i++) c[i] = x.v[i];
  c[7] &= 0x03ffffffu;
}

// c = H(Ru, Rv, m): sponge state [0, Ru, Rv, m, 1], one permutation, word 1.
template <bool DENSE = false>
SB_HD void challenge3(const fq& Ru, const fq& Rv, const fq& m, uint32_t* c) {
  fq s[5] = {fq_zero(), Ru, Rv, m, fq_one()};
  if (DENSE) hades_perm_dense(s); else hades_perm(s);
  truncate250(s[1], c);
}

// c = H(Ru, Rv, R'u, R'v, m): [0, Ru, Rv, R'u, R'v] -> perm -> word1 += m, word2 += 1 -> perm -> word 1.
template <bool DENSE = false>
SB_HD void challenge5(const fq& Ru, const fq& Rv, const fq& Rpu, const fq& Rpv, const fq& m, uint32_t* c) {
  fq s[5] = {fq_zero(), Ru, Rv, Rpu, Rpv};
  if (DENSE) hades_perm_dense(s); else hades_perm(s);
  s[1] = fq_add(s[1], m);
  s[2] = fq_add(s[2], fq_one());
  if (DENSE) hades_perm_dense(s); else hades_perm(s);
  truncate250(s[1], c);
}

}  // namespace sb200
