// Hades252 / Poseidon challenge hash on the FP64 pipe (see fd.cuh for why and how).
// Same permutation as hades.cuh (sparse partial rounds, S-box on the LAST word, dusk-hades), same sponge shapes as
// `dusk_poseidon::sponge::truncated::hash` called from /root/reference/src/signatures.rs:127-134, 275-290; only the
// number representation differs (5 x 52-bit limbs, Montgomery radix 2^260), so the challenge scalar is bit-identical
// (tests: host_arith `h_hades_fd`, GPU `test_hades[fd]`, and every sign / verify parity test).
//
// Every addition of the permutation is folded into the preceding lazy reduction:
//   * round keys ride on the MDS / sparse-row dot product that produces the word (`add` of fd_dot5),
//   * the running words of the sparse rounds are updated by fd_mulc_add (col * x / 2^260 + word),
// so the only stand-alone additions are the first round's keys and the sponge's second absorb.
#pragma once
#include "fd.cuh"
#include "../hades.cuh"

namespace sb200 {

// Constant tables in Montgomery-2^260 form (tools/gen_constants.py):
//   fdh_add[45][5]     integer limbs, 9 groups of 5 words: keys of round 0 | rounds 1, 2, 3 | (pre_0..3, k_0) |
//                      round 63 | rounds 64, 65, 66
//   fdh_mds[25][5]     doubles, row-major MDS;  fdh_post[25][5] doubles (dense matrix after the sparse rounds)
//   fdh_sp_add[59][5]  integer limbs: k_(t+1), carried by round t's row product (zeros for t = 58)
//   fdh_sp_mat[59*9][5] doubles: per round row_t[0..4], col_t[0..3]
#if defined(__CUDACC__)
__device__ const uint64_t d_fdh_add[45][5] = SB200_FDH_ADD_INIT;
__device__ const double d_fdh_mds[25][5] = SB200_FDH_MDS_INIT;
__device__ const double d_fdh_post[25][5] = SB200_FDH_POST_INIT;
__device__ const uint64_t d_fdh_sp_add[59][5] = SB200_FDH_SP_ADD_INIT;
__device__ const double d_fdh_sp_mat[59 * 9][5] = SB200_FDH_SP_MAT_INIT;
#endif
static const uint64_t h_fdh_add[45][5] = SB200_FDH_ADD_INIT;
static const double h_fdh_mds[25][5] = SB200_FDH_MDS_INIT;
static const double h_fdh_post[25][5] = SB200_FDH_POST_INIT;
static const uint64_t h_fdh_sp_add[59][5] = SB200_FDH_SP_ADD_INIT;
static const double h_fdh_sp_mat[59 * 9][5] = SB200_FDH_SP_MAT_INIT;

SB_HD fd fd_pow5(const fd& x) {
  fdd xd = fd_todbl(x);
  fd x2 = fd_sqr(xd);
  fd x4 = fd_sqr(fd_todbl(x2));
  return fd_mul(fd_todbl(x4), xd);
}

// s <- M s + add  (dense 5 x 5, one lazily reduced dot product per row; add = 5 words of 5 limbs or nullptr)
SB_HD void fdh_matmul(fd* s, const double (*mat)[5], const uint64_t (*add)[5]) {
  fdd d[5];
#pragma unroll
  for (int k = 0; k < 5; k++) d[k] = fd_todbl(s[k]);
#pragma unroll 1
  for (int k = 0; k < 5; k++) s[k] = fd_dot5(&mat[5 * k][0], add ? &add[k][0] : nullptr, d[0], d[1], d[2], d[3], d[4]);
}

// state in Montgomery-2^260 form, lazily reduced (< 2^257) in and out
SB_HD void hades_perm_fd(fd* s) {
#pragma unroll
  for (int k = 0; k < 5; k++) s[k] = fd_add(s[k], SB_CONST(fdh_add)[k]);
  // 4 full rounds; the MDS product of round r carries the keys of round r + 1 (after round 3: pre_0..3 and k_0)
#pragma unroll 1
  for (int r = 0; r < 4; r++) {
#pragma unroll 1
    for (int k = 0; k < 5; k++) s[k] = fd_pow5(s[k]);
    fdh_matmul(s, SB_CONST(fdh_mds), &SB_CONST(fdh_add)[5 * (r + 1)]);
  }
  // 59 sparse partial rounds
#pragma unroll 1
  for (int t = 0; t < 59; t++) {
    const double(*m)[5] = &SB_CONST(fdh_sp_mat)[9 * t];
    fdd x = fd_todbl(fd_pow5(s[4]));
    {
      fdd d0 = fd_todbl(s[0]), d1 = fd_todbl(s[1]), d2 = fd_todbl(s[2]), d3 = fd_todbl(s[3]);
      s[4] = fd_dot5(&m[0][0], SB_CONST(fdh_sp_add)[t], d0, d1, d2, d3, x);
    }
#pragma unroll 1
    for (int j = 0; j < 4; j++) s[j] = fd_mulc_add(&m[5 + j][0], x, s[j]);
  }
  fdh_matmul(s, SB_CONST(fdh_post), &SB_CONST(fdh_add)[25]);
#pragma unroll 1
  for (int r = 0; r < 4; r++) {
#pragma unroll 1
    for (int k = 0; k < 5; k++) s[k] = fd_pow5(s[k]);
    fdh_matmul(s, SB_CONST(fdh_mds), r < 3 ? &SB_CONST(fdh_add)[5 * (6 + r)] : nullptr);
  }
}

// ---- two permutations per thread, interleaved (fd.cuh "paired" primitives) -------------------------------------
SB_HD fd2 fd_pow5_2(const fd2& x) {
  fdd2 xd = fd_todbl2(x);
  fd2 x2 = fd_reduce2(fd_sqr_cols2(xd));
  fd2 x4 = fd_reduce2(fd_sqr_cols2(fd_todbl2(x2)));
  return fd_reduce2(fd_macv2(fd_cols_init2(1), fd_todbl2(x4), xd));
}
SB_HD void fdh_matmul2(fd2* s, const double (*mat)[5], const uint64_t (*add)[5]) {
  fdd2 d[5];
#pragma unroll
  for (int k = 0; k < 5; k++) d[k] = fd_todbl2(s[k]);
#pragma unroll 1
  for (int r = 0; r < 5; r++) {
    fdc2 c = fd_cols_init2(5);
    if (add) {
      fd_cols_add(c.a, add[r]);
      fd_cols_add(c.b, add[r]);
    }
#pragma unroll
    for (int j = 0; j < 5; j++) c = fd_macc2(c, d[j], &mat[5 * r + j][0]);
    s[r] = fd_reduce2(c);
  }
}
SB_HD void hades_perm_fd2(fd2* s) {
#pragma unroll
  for (int k = 0; k < 5; k++) {
    s[k].a = fd_add(s[k].a, SB_CONST(fdh_add)[k]);
    s[k].b = fd_add(s[k].b, SB_CONST(fdh_add)[k]);
  }
#pragma unroll 1
  for (int r = 0; r < 4; r++) {
#pragma unroll 1
    for (int k = 0; k < 5; k++) s[k] = fd_pow5_2(s[k]);
    fdh_matmul2(s, SB_CONST(fdh_mds), &SB_CONST(fdh_add)[5 * (r + 1)]);
  }
#pragma unroll 1
  for (int t = 0; t < 59; t++) {
    const double(*m)[5] = &SB_CONST(fdh_sp_mat)[9 * t];
    fdd2 x = fd_todbl2(fd_pow5_2(s[4]));
    {
      fdc2 c = fd_cols_init2(5);
      fd_cols_add(c.a, SB_CONST(fdh_sp_add)[t]);
      fd_cols_add(c.b, SB_CONST(fdh_sp_add)[t]);
#pragma unroll
      for (int j = 0; j < 4; j++) c = fd_macc2(c, fd_todbl2(s[j]), &m[j][0]);
      c = fd_macc2(c, x, &m[4][0]);
      s[4] = fd_reduce2(c);
    }
#pragma unroll 1
    for (int j = 0; j < 4; j++) {
      fdc2 c = fd_cols_init2(1);
      fd_cols_add_csub(c.a, s[j].a);
      fd_cols_add_csub(c.b, s[j].b);
      s[j] = fd_reduce2(fd_macc2(c, x, &m[5 + j][0]));
    }
  }
  fdh_matmul2(s, SB_CONST(fdh_post), &SB_CONST(fdh_add)[25]);
#pragma unroll 1
  for (int r = 0; r < 4; r++) {
#pragma unroll 1
    for (int k = 0; k < 5; k++) s[k] = fd_pow5_2(s[k]);
    fdh_matmul2(s, SB_CONST(fdh_mds), r < 3 ? &SB_CONST(fdh_add)[5 * (6 + r)] : nullptr);
  }
}

// ---- memory-operand permutation: state and temporaries live in 8 slots (fd.cuh), stride `ls` between limbs -----
// slots 0..4 = state words, 5..7 = S-box temporaries; slot j starts at base + j * 10 * ls doubles.
constexpr int FDH_SLOTS = 8;
SB_HD double* fdh_slot(double* base, int ls, int j) { return base + (size_t)j * 10 * ls; }

// word <- word^5 (in place; uses temporaries 5 and 6)
SB_HD void fdh_pow5_p(double* base, int ls, int w) {
  double *x = fdh_slot(base, ls, w), *x2 = fdh_slot(base, ls, 5), *x4 = fdh_slot(base, ls, 6);
  fd_reduce_p(fd_sqr_p(x, ls), x2, ls);
  fd_reduce_p(fd_sqr_p(x2, ls), x4, ls);
  fd_reduce_p(fd_mac_p(fd_start_p(1, nullptr, 0, false), x4, ls, x, ls), x, ls);
}

// state <- M state + add.  Every output word needs all five input words, so the operand forms of the inputs are
// first copied aside into the temporaries' space (slots 5..7 = 30 words >= 25) and the rows overwrite the state.
SB_HD void fdh_matmul_p(double* base, int ls, const double (*mat)[5], const uint64_t (*add)[5]) {
  double* tmp = fdh_slot(base, ls, 5);
#pragma unroll 1
  for (int k = 0; k < 25; k++) tmp[k * ls] = fdh_slot(base, ls, k / 5)[(k % 5) * ls];
#pragma unroll 1
  for (int r = 0; r < 5; r++) {
    fdc c = fd_start_p(5, add ? &add[r][0] : nullptr, 1, false);
#pragma unroll 1
    for (int j = 0; j < 5; j++) c = fd_mac_p(c, tmp + 5 * j * ls, ls, &mat[5 * r + j][0], 1);
    fd_reduce_p(c, fdh_slot(base, ls, r), ls);
  }
}

SB_HD void hades_perm_fd_p(double* base, int ls) {
  // round-0 keys
#pragma unroll 1
  for (int k = 0; k < 5; k++) fd_slot_store(fdh_slot(base, ls, k), ls, fd_add(fd_slot_load(fdh_slot(base, ls, k), ls), SB_CONST(fdh_add)[k]));
#pragma unroll 1
  for (int r = 0; r < 4; r++) {
#pragma unroll 1
    for (int k = 0; k < 5; k++) fdh_pow5_p(base, ls, k);
    fdh_matmul_p(base, ls, SB_CONST(fdh_mds), &SB_CONST(fdh_add)[5 * (r + 1)]);
  }
  double *s4 = fdh_slot(base, ls, 4), *x = fdh_slot(base, ls, 7);
#pragma unroll 1
  for (int t = 0; t < 59; t++) {
    const double(*m)[5] = &SB_CONST(fdh_sp_mat)[9 * t];
    {  // x = s4^5 into slot 7
      double *x2 = fdh_slot(base, ls, 5), *x4 = fdh_slot(base, ls, 6);
      fd_reduce_p(fd_sqr_p(s4, ls), x2, ls);
      fd_reduce_p(fd_sqr_p(x2, ls), x4, ls);
      fd_reduce_p(fd_mac_p(fd_start_p(1, nullptr, 0, false), x4, ls, s4, ls), x, ls);
    }
    fdc c = fd_start_p(5, SB_CONST(fdh_sp_add)[t], 1, false);
#pragma unroll 1
    for (int j = 0; j < 4; j++) c = fd_mac_p(c, fdh_slot(base, ls, j), ls, &m[j][0], 1);
    c = fd_mac_p(c, x, ls, &m[4][0], 1);
    fd_reduce_p(c, s4, ls);
#pragma unroll 1
    for (int j = 0; j < 4; j++) {
      double* sj = fdh_slot(base, ls, j);
      fd_reduce_p(fd_mac_p(fd_start_p(1, fd_slot_ints(sj, ls), ls, true), x, ls, &m[5 + j][0], 1), sj, ls);
    }
  }
  fdh_matmul_p(base, ls, SB_CONST(fdh_post), &SB_CONST(fdh_add)[25]);
#pragma unroll 1
  for (int r = 0; r < 4; r++) {
#pragma unroll 1
    for (int k = 0; k < 5; k++) fdh_pow5_p(base, ls, k);
    fdh_matmul_p(base, ls, SB_CONST(fdh_mds), r < 3 ? &SB_CONST(fdh_add)[5 * (6 + r)] : nullptr);
  }
}

// c = H(Ru, Rv, m) with the state in caller-provided slot storage (FDH_SLOTS * 10 * ls doubles)
SB_HD void challenge3_fd_p(const fq& Ru, const fq& Rv, const fq& m, uint32_t* c, double* base, int ls);

// low 250 bits of the canonical integer (dusk_poseidon::sponge::truncated)
SB_HD void fd_truncate250(const fd& word, uint32_t* c) {
  fd_to_canonical(word, c);
  c[7] &= 0x03ffffffu;
}

SB_HD fd fd_const_one() {  // Montgomery-2^260 one
  const fd c = {SB200_FD_ONE_U_INIT};
  return c;
}
SB_HD fd fd_const_zero() {
  fd r;
#pragma unroll
  for (int i = 0; i < 5; i++) r.l[i] = 0;
  return r;
}

// c = H(Ru, Rv, m): sponge state [0, Ru, Rv, m, 1], one permutation, word 1.
SB_HD void challenge3_fd(const fq& Ru, const fq& Rv, const fq& m, uint32_t* c) {
  fd s[5] = {fd_const_zero(), fd_from_fq(Ru), fd_from_fq(Rv), fd_from_fq(m), fd_const_one()};
  hades_perm_fd(s);
  fd_truncate250(s[1], c);
}

// two challenges at once: c[w] = H(Ru[w], Rv[w], m[w])
SB_HD void challenge3_fd2(const fq* Ru, const fq* Rv, const fq* m, uint32_t (*c)[8]) {
  fd2 s[5];
  s[0] = {fd_const_zero(), fd_const_zero()};
  s[1] = {fd_from_fq(Ru[0]), fd_from_fq(Ru[1])};
  s[2] = {fd_from_fq(Rv[0]), fd_from_fq(Rv[1])};
  s[3] = {fd_from_fq(m[0]), fd_from_fq(m[1])};
  s[4] = {fd_const_one(), fd_const_one()};
  hades_perm_fd2(s);
  fd_truncate250(s[1].a, c[0]);
  fd_truncate250(s[1].b, c[1]);
}

SB_HD void challenge3_fd_p(const fq& Ru, const fq& Rv, const fq& m, uint32_t* c, double* base, int ls) {
  fd_slot_store(fdh_slot(base, ls, 0), ls, fd_const_zero());
  fd_slot_store(fdh_slot(base, ls, 1), ls, fd_from_fq(Ru));
  fd_slot_store(fdh_slot(base, ls, 2), ls, fd_from_fq(Rv));
  fd_slot_store(fdh_slot(base, ls, 3), ls, fd_from_fq(m));
  fd_slot_store(fdh_slot(base, ls, 4), ls, fd_const_one());
  hades_perm_fd_p(base, ls);
  fd_truncate250(fd_slot_load(fdh_slot(base, ls, 1), ls), c);
}

// c = H(Ru, Rv, R'u, R'v, m): [0, Ru, Rv, R'u, R'v] -> perm -> word1 += m, word2 += 1 -> perm -> word 1.
SB_HD void challenge5_fd(const fq& Ru, const fq& Rv, const fq& Rpu, const fq& Rpv, const fq& m, uint32_t* c) {
  fd s[5] = {fd_const_zero(), fd_from_fq(Ru), fd_from_fq(Rv), fd_from_fq(Rpu), fd_from_fq(Rpv)};
  hades_perm_fd(s);
  fd mm = fd_from_fq(m), one = fd_const_one();
  s[1] = fd_add(s[1], mm.l);
  s[2] = fd_add(s[2], one.l);
  hades_perm_fd(s);
  fd_truncate250(s[1], c);
}

}  // namespace sb200
