/* CPU restatement of dusk-schnorr's sign / verify ALGORITHM -- TEST INFRASTRUCTURE ONLY.
 *
 * Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / `--impl reference` leg may
 * load this library.  The product (schnorr_b200/) never does.
 *
 * PARITY STATUS: parity unpinned against the real crate (see oracle/schnorr_oracle.py's header):
 * the reference's arithmetic lives in un-vendored Rust crates (dusk-bls12_381 ^0.13,
 * dusk-jubjub ^0.14, dusk-poseidon ^0.33 / dusk-hades) that cannot be built here.  This file
 * restates THEIR algorithms the way the reference executes them, so that its run time is a fair
 * stand-in for "the reference's CPU path" and its outputs are a second, independent check of the
 * GPU results at batch sizes the Python oracle cannot reach:
 *   - 4 x 64-bit Montgomery F_q / F_r with 128-bit products (dusk-bls12_381 Scalar, dusk-jubjub Fr);
 *   - `JubJubExtended * JubJubScalar` = 252-step MSB-first double-and-add over the extended-Niels
 *     form with a constant-time select -- no windows, no fixed-base tables;
 *   - `to_hash_inputs` = one field inversion per point (Fermat);
 *   - dense Hades252 (8 full + 59 partial rounds, 25 multiplications per MDS layer);
 *   - projective equality.
 * Call-for-call it follows /root/reference/src/keys/public.rs:121-130, 222-244, 401-415 and
 * /root/reference/src/keys/secret.rs:150-168, 217-240, 433-451.
 * It is validated against oracle/schnorr_oracle.py (tests/test_oracle.py); all numeric constants
 * are handed in by that Python oracle through ref_init(), so there is a single source for them.
 * Batches are split over pthreads (the stand-in for "rayon over the batch").
 */
#include <pthread.h>
#include <stdint.h>
#include <stdlib.h>
#include <string.h>

typedef unsigned __int128 u128;
typedef struct { uint64_t l[4]; } fe;          /* Montgomery form, canonical */
typedef struct { fe u, v, z, t1, t2; } ext_t;  /* JubJubExtended */
typedef struct { fe vpu, vmu, z, t2d; } niels_t;

/* ---- constants, set by ref_init ---------------------------------------------------------- */
static uint64_t QM[4], RM[4];            /* moduli */
static uint64_t QINV, RINV;              /* -p^-1 mod 2^64 */
static fe Q_R1, Q_R2, R_R2;              /* 2^256, 2^512 mod q; 2^512 mod r */
static fe ED_2D;
static ext_t GEN, GEN_NUMS;
static fe RC[335];
static fe MDS[5][5];

/* ---- generic 4-limb Montgomery ------------------------------------------------------------- */
static inline int geq(const uint64_t* a, const uint64_t* m) {
  for (int i = 3; i >= 0; i--) { if (a[i] > m[i]) return 1; if (a[i] < m[i]) return 0; }
  return 1;
}
static inline void sub_mod_raw(uint64_t* a, const uint64_t* m) {
  u128 b = 0;
  for (int i = 0; i < 4; i++) { u128 t = (u128)a[i] - m[i] - (uint64_t)b; a[i] = (uint64_t)t; b = (t >> 64) & 1; }
}
static void mont_mul(uint64_t* r, const uint64_t* a, const uint64_t* b, const uint64_t* m, uint64_t inv) {
  uint64_t t[9] = {0};
  for (int i = 0; i < 4; i++) {
    u128 c = 0;
    for (int j = 0; j < 4; j++) { c += (u128)a[i] * b[j] + t[i + j]; t[i + j] = (uint64_t)c; c >>= 64; }
    t[i + 4] = (uint64_t)c;
  }
  /* montgomery_reduce */
  uint64_t top = 0;
  for (int i = 0; i < 4; i++) {
    uint64_t k = t[i] * inv;
    u128 c = 0;
    for (int j = 0; j < 4; j++) { c += (u128)k * m[j] + t[i + j]; t[i + j] = (uint64_t)c; c >>= 64; }
    for (int j = i + 4; j < 8; j++) { c += t[j]; t[j] = (uint64_t)c; c >>= 64; }
    top += (uint64_t)c;
  }
  uint64_t res[4] = {t[4], t[5], t[6], t[7]};
  if (top || geq(res, m)) sub_mod_raw(res, m);
  memcpy(r, res, 32);
}
static inline fe fq_mul(fe a, fe b) { fe r; mont_mul(r.l, a.l, b.l, QM, QINV); return r; }
static inline fe fq_sqr(fe a) { return fq_mul(a, a); }
static inline fe fq_add(fe a, fe b) {
  fe r; u128 c = 0;
  for (int i = 0; i < 4; i++) { c += (u128)a.l[i] + b.l[i]; r.l[i] = (uint64_t)c; c >>= 64; }
  if (geq(r.l, QM)) sub_mod_raw(r.l, QM);
  return r;
}
static inline fe fq_sub(fe a, fe b) {
  fe r; u128 bw = 0;
  for (int i = 0; i < 4; i++) { u128 t = (u128)a.l[i] - b.l[i] - (uint64_t)bw; r.l[i] = (uint64_t)t; bw = (t >> 64) & 1; }
  if (bw) { u128 c = 0; for (int i = 0; i < 4; i++) { c += (u128)r.l[i] + QM[i]; r.l[i] = (uint64_t)c; c >>= 64; } }
  return r;
}
static inline fe fq_dbl(fe a) { return fq_add(a, a); }
static inline int fq_eq(fe a, fe b) { return memcmp(a.l, b.l, 32) == 0; }
static fe fq_inv(fe a) { /* a^(q-2), plain square-and-multiply over the 255 exponent bits */
  uint64_t e[4] = {QM[0] - 2, QM[1], QM[2], QM[3]};
  fe acc = Q_R1;
  for (int i = 254; i >= 0; i--) {
    acc = fq_sqr(acc);
    if ((e[i >> 6] >> (i & 63)) & 1) acc = fq_mul(acc, a);
  }
  return acc;
}

/* ---- JubJub, as dusk-jubjub executes it ---------------------------------------------------- */
static ext_t ext_identity(void) { ext_t p; memset(&p, 0, sizeof p); p.v = Q_R1; p.z = Q_R1; return p; }
static ext_t completed_into_extended(fe u, fe v, fe z, fe t) {
  ext_t r; r.u = fq_mul(u, t); r.v = fq_mul(v, z); r.z = fq_mul(z, t); r.t1 = u; r.t2 = v; return r;
}
static ext_t ext_double(const ext_t* p) {
  fe uu = fq_sqr(p->u), vv = fq_sqr(p->v), zz2 = fq_dbl(fq_sqr(p->z));
  fe uv2 = fq_sqr(fq_add(p->u, p->v));
  fe vpu = fq_add(vv, uu), vmu = fq_sub(vv, uu);
  return completed_into_extended(fq_sub(uv2, vpu), vpu, vmu, fq_sub(zz2, vmu));
}
static niels_t ext_to_niels(const ext_t* p) {
  niels_t n; n.vpu = fq_add(p->v, p->u); n.vmu = fq_sub(p->v, p->u); n.z = p->z;
  n.t2d = fq_mul(fq_mul(p->t1, p->t2), ED_2D); return n;
}
static ext_t ext_add_niels(const ext_t* p, const niels_t* q) {
  fe a = fq_mul(fq_sub(p->v, p->u), q->vmu), b = fq_mul(fq_add(p->v, p->u), q->vpu);
  fe c = fq_mul(fq_mul(p->t1, p->t2), q->t2d), d = fq_dbl(fq_mul(p->z, q->z));
  return completed_into_extended(fq_sub(b, a), fq_add(b, a), fq_add(d, c), fq_sub(d, c));
}
static ext_t ext_add(const ext_t* p, const ext_t* q) { niels_t n = ext_to_niels(q); return ext_add_niels(p, &n); }
/* `P * s`: 252 iterations over the bits of the canonical scalar, MSB first, skipping the top 4 */
static ext_t ext_mul(const ext_t* p, const uint64_t* s) {
  niels_t zero, base = ext_to_niels(p);
  zero.vpu = Q_R1; zero.vmu = Q_R1; zero.z = Q_R1; memset(&zero.t2d, 0, sizeof(fe));
  ext_t acc = ext_identity();
  for (int i = 251; i >= 0; i--) {
    acc = ext_double(&acc);
    uint64_t bit = (s[i >> 6] >> (i & 63)) & 1, mask = 0 - bit;
    niels_t sel;
    const uint64_t* zp = (const uint64_t*)&zero; const uint64_t* bp = (const uint64_t*)&base; uint64_t* sp = (uint64_t*)&sel;
    for (size_t k = 0; k < sizeof(niels_t) / 8; k++) sp[k] = (zp[k] & ~mask) | (bp[k] & mask);
    acc = ext_add_niels(&acc, &sel);
  }
  return acc;
}
static int ext_eq(const ext_t* a, const ext_t* b) {
  return fq_eq(fq_mul(a->u, b->z), fq_mul(b->u, a->z)) & fq_eq(fq_mul(a->v, b->z), fq_mul(b->v, a->z));
}
static void ext_to_affine(const ext_t* p, fe* u, fe* v) { fe zi = fq_inv(p->z); *u = fq_mul(p->u, zi); *v = fq_mul(p->v, zi); }

/* ---- Hades252 + sponge ---------------------------------------------------------------------- */
static inline fe pow5(fe x) { fe x2 = fq_sqr(x), x4 = fq_sqr(x2); return fq_mul(x4, x); }
static void hades_perm(fe* s) {
  int ci = 0;
  for (int rnd = 0; rnd < 67; rnd++) {
    for (int k = 0; k < 5; k++) s[k] = fq_add(s[k], RC[ci++]);
    if (rnd < 4 || rnd >= 63) { for (int k = 0; k < 5; k++) s[k] = pow5(s[k]); } else s[4] = pow5(s[4]);
    fe r[5];
    for (int k = 0; k < 5; k++) {
      fe a = fq_mul(MDS[k][0], s[0]);
      for (int j = 1; j < 5; j++) a = fq_add(a, fq_mul(MDS[k][j], s[j]));
      r[k] = a;
    }
    memcpy(s, r, sizeof r);
  }
}
static void truncate250(fe word, uint64_t* c) {
  uint64_t one[4] = {1, 0, 0, 0};
  mont_mul(c, word.l, one, QM, QINV); /* out of Montgomery form */
  c[3] &= 0x03ffffffffffffffULL;
}
static void challenge3(fe ru, fe rv, fe m, uint64_t* c) {
  fe s[5]; memset(&s[0], 0, sizeof(fe)); s[1] = ru; s[2] = rv; s[3] = m; s[4] = Q_R1;
  hades_perm(s); truncate250(s[1], c);
}
static void challenge5(fe ru, fe rv, fe pu, fe pv, fe m, uint64_t* c) {
  fe s[5]; memset(&s[0], 0, sizeof(fe)); s[1] = ru; s[2] = rv; s[3] = pu; s[4] = pv;
  hades_perm(s); s[1] = fq_add(s[1], m); s[2] = fq_add(s[2], Q_R1); hades_perm(s); truncate250(s[1], c);
}
/* u = nonce - c * sk mod r */
static void sign_finish(const uint64_t* nonce, const uint64_t* c, const uint64_t* sk, uint64_t* u) {
  uint64_t t[4], p[4];
  mont_mul(t, c, sk, RM, RINV); mont_mul(p, t, R_R2.l, RM, RINV);
  u128 bw = 0;
  for (int i = 0; i < 4; i++) { u128 d = (u128)nonce[i] - p[i] - (uint64_t)bw; u[i] = (uint64_t)d; bw = (d >> 64) & 1; }
  if (bw) { u128 cc = 0; for (int i = 0; i < 4; i++) { cc += (u128)u[i] + RM[i]; u[i] = (uint64_t)cc; cc >>= 64; } }
}

/* ---- ABI-layout loaders (same arrays the GPU library takes) ------------------------------- */
static ext_t load_point(const uint64_t* p, int affine) {
  ext_t e; memcpy(e.u.l, p, 32); memcpy(e.v.l, p + 4, 32);
  if (affine) e.z = Q_R1; else memcpy(e.z.l, p + 8, 32);
  /* t1 * t2 = u*v/z: with (U,V,Z) only, rebuild an equal extended point (U Z : V Z : Z^2 : U V) */
  if (affine) { e.t1 = e.u; e.t2 = e.v; }
  else { fe U = e.u, V = e.v, Z = e.z; e.u = fq_mul(U, Z); e.v = fq_mul(V, Z); e.z = fq_sqr(Z); e.t1 = U; e.t2 = V; }
  return e;
}
static void store_affine(uint64_t* o, const ext_t* p) { fe u, v; ext_to_affine(p, &u, &v); memcpy(o, u.l, 32); memcpy(o + 4, v.l, 32); }

typedef struct {
  int op, affine; int64_t lo, hi;
  const uint64_t *a0, *a1, *a2, *a3, *a4, *a5;
  uint64_t *o0, *o1, *o2, *o3; uint8_t* verdict;
} job_t;

enum { OP_VERIFY, OP_VERIFY_DOUBLE, OP_VERIFY_VARGEN, OP_SIGN, OP_SIGN_DOUBLE, OP_SIGN_VARGEN, OP_KEYGEN, OP_SMUL };

static void* worker(void* arg) {
  job_t* j = (job_t*)arg;
  int pw = j->affine ? 8 : 12;
  for (int64_t i = j->lo; i < j->hi; i++) {
    uint64_t c[4];
    switch (j->op) {
      case OP_VERIFY: { /* a0 pk, a1 u, a2 R, a3 m        public.rs:121-130 */
        ext_t pk = load_point(j->a0 + i * pw, j->affine), R = load_point(j->a2 + i * pw, j->affine);
        fe ru, rv, m; ext_to_affine(&R, &ru, &rv); memcpy(m.l, j->a3 + i * 4, 32);
        challenge3(ru, rv, m, c);
        ext_t p1 = ext_mul(&GEN, j->a1 + i * 4), p2 = ext_mul(&pk, c), s = ext_add(&p1, &p2);
        j->verdict[i] = (uint8_t)ext_eq(&s, &R);
        if (j->o0) memcpy(j->o0 + i * 4, c, 32);
      } break;
      case OP_VERIFY_DOUBLE: { /* a0 pk, a1 pk', a2 u, a3 R, a4 R', a5 m   public.rs:222-244 */
        ext_t pk = load_point(j->a0 + i * pw, j->affine), pkp = load_point(j->a1 + i * pw, j->affine);
        ext_t R = load_point(j->a3 + i * pw, j->affine), Rp = load_point(j->a4 + i * pw, j->affine);
        fe ru, rv, pu, pv, m; ext_to_affine(&R, &ru, &rv); ext_to_affine(&Rp, &pu, &pv); memcpy(m.l, j->a5 + i * 4, 32);
        challenge5(ru, rv, pu, pv, m, c);
        ext_t a = ext_mul(&GEN, j->a2 + i * 4), b = ext_mul(&pk, c), s1 = ext_add(&a, &b);
        ext_t a2 = ext_mul(&GEN_NUMS, j->a2 + i * 4), b2 = ext_mul(&pkp, c), s2 = ext_add(&a2, &b2);
        j->verdict[i] = (uint8_t)(ext_eq(&s1, &R) && ext_eq(&s2, &Rp));
        if (j->o0) memcpy(j->o0 + i * 4, c, 32);
      } break;
      case OP_VERIFY_VARGEN: { /* a0 pk, a1 gen, a2 u, a3 R, a4 m   public.rs:401-415 */
        ext_t pk = load_point(j->a0 + i * pw, j->affine), g = load_point(j->a1 + i * pw, j->affine);
        ext_t R = load_point(j->a3 + i * pw, j->affine);
        fe ru, rv, m; ext_to_affine(&R, &ru, &rv); memcpy(m.l, j->a4 + i * 4, 32);
        challenge3(ru, rv, m, c);
        ext_t a = ext_mul(&g, j->a2 + i * 4), b = ext_mul(&pk, c), s = ext_add(&a, &b);
        j->verdict[i] = (uint8_t)ext_eq(&s, &R);
        if (j->o0) memcpy(j->o0 + i * 4, c, 32);
      } break;
      case OP_SIGN: { /* a0 sk, a1 m, a2 nonce -> o0 u, o1 R(aff), o3 c   secret.rs:150-168 */
        ext_t R = ext_mul(&GEN, j->a2 + i * 4);
        fe ru, rv, m; ext_to_affine(&R, &ru, &rv); memcpy(m.l, j->a1 + i * 4, 32);
        challenge3(ru, rv, m, c);
        sign_finish(j->a2 + i * 4, c, j->a0 + i * 4, j->o0 + i * 4);
        memcpy(j->o1 + i * 8, ru.l, 32); memcpy(j->o1 + i * 8 + 4, rv.l, 32);
        if (j->o3) memcpy(j->o3 + i * 4, c, 32);
      } break;
      case OP_SIGN_DOUBLE: { /* secret.rs:217-240 */
        ext_t R = ext_mul(&GEN, j->a2 + i * 4), Rp = ext_mul(&GEN_NUMS, j->a2 + i * 4);
        fe ru, rv, pu, pv, m; ext_to_affine(&R, &ru, &rv); ext_to_affine(&Rp, &pu, &pv); memcpy(m.l, j->a1 + i * 4, 32);
        challenge5(ru, rv, pu, pv, m, c);
        sign_finish(j->a2 + i * 4, c, j->a0 + i * 4, j->o0 + i * 4);
        memcpy(j->o1 + i * 8, ru.l, 32); memcpy(j->o1 + i * 8 + 4, rv.l, 32);
        memcpy(j->o2 + i * 8, pu.l, 32); memcpy(j->o2 + i * 8 + 4, pv.l, 32);
        if (j->o3) memcpy(j->o3 + i * 4, c, 32);
      } break;
      case OP_SIGN_VARGEN: { /* a0 sk, a1 gen, a2 m, a3 nonce   secret.rs:433-451 */
        ext_t g = load_point(j->a1 + i * pw, j->affine), R = ext_mul(&g, j->a3 + i * 4);
        fe ru, rv, m; ext_to_affine(&R, &ru, &rv); memcpy(m.l, j->a2 + i * 4, 32);
        challenge3(ru, rv, m, c);
        sign_finish(j->a3 + i * 4, c, j->a0 + i * 4, j->o0 + i * 4);
        memcpy(j->o1 + i * 8, ru.l, 32); memcpy(j->o1 + i * 8 + 4, rv.l, 32);
        if (j->o3) memcpy(j->o3 + i * 4, c, 32);
      } break;
      case OP_KEYGEN: { /* a0 sk; a1 = optional per-key generator; o0 pk, o1 pk' (optional)  public.rs:61-67,265-272,337-344 */
        ext_t base = j->a1 ? load_point(j->a1 + i * pw, j->affine) : GEN;
        ext_t p = ext_mul(&base, j->a0 + i * 4); store_affine(j->o0 + i * 8, &p);
        if (j->o1) { ext_t pp = ext_mul(&GEN_NUMS, j->a0 + i * 4); store_affine(j->o1 + i * 8, &pp); }
      } break;
      case OP_SMUL: { /* a0 points, a1 k -> o0 affine */
        ext_t b = load_point(j->a0 + i * pw, j->affine), p = ext_mul(&b, j->a1 + i * 4); store_affine(j->o0 + i * 8, &p);
      } break;
    }
  }
  return 0;
}

static void run_batch(job_t proto, int64_t n, int nthreads) {
  if (nthreads < 1) nthreads = 1;
  if (nthreads > 256) nthreads = 256;
  if ((int64_t)nthreads > n) nthreads = n > 0 ? (int)n : 1;
  pthread_t th[256]; job_t jobs[256];
  int64_t per = (n + nthreads - 1) / nthreads;
  for (int t = 0; t < nthreads; t++) {
    jobs[t] = proto; jobs[t].lo = t * per; jobs[t].hi = (t + 1) * per < n ? (t + 1) * per : n;
    if (jobs[t].lo > n) jobs[t].lo = n;
    pthread_create(&th[t], 0, worker, &jobs[t]);
  }
  for (int t = 0; t < nthreads; t++) pthread_join(th[t], 0);
}

/* blob (u64 words, all field elements already in Montgomery form):
 *   q[4] r[4] qinv rinv  q_r1[4] q_r2[4] r_r2[4]  ed_2d[4]  G.u[4] G.v[4] G'.u[4] G'.v[4]  rc[335][4]  mds[25][4] */
void ref_init(const uint64_t* b) {
  memcpy(QM, b, 32); b += 4; memcpy(RM, b, 32); b += 4; QINV = *b++; RINV = *b++;
  memcpy(Q_R1.l, b, 32); b += 4; memcpy(Q_R2.l, b, 32); b += 4; memcpy(R_R2.l, b, 32); b += 4;
  memcpy(ED_2D.l, b, 32); b += 4;
  fe gu, gv, hu, hv;
  memcpy(gu.l, b, 32); b += 4; memcpy(gv.l, b, 32); b += 4; memcpy(hu.l, b, 32); b += 4; memcpy(hv.l, b, 32); b += 4;
  GEN.u = gu; GEN.v = gv; GEN.z = Q_R1; GEN.t1 = gu; GEN.t2 = gv;
  GEN_NUMS.u = hu; GEN_NUMS.v = hv; GEN_NUMS.z = Q_R1; GEN_NUMS.t1 = hu; GEN_NUMS.t2 = hv;
  for (int i = 0; i < 335; i++) { memcpy(RC[i].l, b, 32); b += 4; }
  for (int i = 0; i < 5; i++) for (int k = 0; k < 5; k++) { memcpy(MDS[i][k].l, b, 32); b += 4; }
}

#define PROTO(OP, AFF) job_t p; memset(&p, 0, sizeof p); p.op = (OP); p.affine = (AFF)
void ref_verify(int64_t n, int affine, const uint64_t* pk, const uint64_t* u, const uint64_t* R, const uint64_t* m,
                uint8_t* verdict, uint64_t* c_out, int nthreads) {
  PROTO(OP_VERIFY, affine); p.a0 = pk; p.a1 = u; p.a2 = R; p.a3 = m; p.verdict = verdict; p.o0 = c_out; run_batch(p, n, nthreads);
}
void ref_verify_double(int64_t n, int affine, const uint64_t* pk, const uint64_t* pkp, const uint64_t* u, const uint64_t* R,
                       const uint64_t* Rp, const uint64_t* m, uint8_t* verdict, uint64_t* c_out, int nthreads) {
  PROTO(OP_VERIFY_DOUBLE, affine); p.a0 = pk; p.a1 = pkp; p.a2 = u; p.a3 = R; p.a4 = Rp; p.a5 = m; p.verdict = verdict; p.o0 = c_out;
  run_batch(p, n, nthreads);
}
void ref_verify_vargen(int64_t n, int affine, const uint64_t* pk, const uint64_t* gen, const uint64_t* u, const uint64_t* R,
                       const uint64_t* m, uint8_t* verdict, uint64_t* c_out, int nthreads) {
  PROTO(OP_VERIFY_VARGEN, affine); p.a0 = pk; p.a1 = gen; p.a2 = u; p.a3 = R; p.a4 = m; p.verdict = verdict; p.o0 = c_out;
  run_batch(p, n, nthreads);
}
void ref_sign(int64_t n, const uint64_t* sk, const uint64_t* m, const uint64_t* nonce, uint64_t* u, uint64_t* R, uint64_t* c, int nthreads) {
  PROTO(OP_SIGN, 1); p.a0 = sk; p.a1 = m; p.a2 = nonce; p.o0 = u; p.o1 = R; p.o3 = c; run_batch(p, n, nthreads);
}
void ref_sign_double(int64_t n, const uint64_t* sk, const uint64_t* m, const uint64_t* nonce, uint64_t* u, uint64_t* R, uint64_t* Rp,
                     uint64_t* c, int nthreads) {
  PROTO(OP_SIGN_DOUBLE, 1); p.a0 = sk; p.a1 = m; p.a2 = nonce; p.o0 = u; p.o1 = R; p.o2 = Rp; p.o3 = c; run_batch(p, n, nthreads);
}
void ref_sign_vargen(int64_t n, int affine, const uint64_t* sk, const uint64_t* gen, const uint64_t* m, const uint64_t* nonce,
                     uint64_t* u, uint64_t* R, uint64_t* c, int nthreads) {
  PROTO(OP_SIGN_VARGEN, affine); p.a0 = sk; p.a1 = gen; p.a2 = m; p.a3 = nonce; p.o0 = u; p.o1 = R; p.o3 = c; run_batch(p, n, nthreads);
}
void ref_keygen(int64_t n, int affine, const uint64_t* sk, const uint64_t* gen_or_null, uint64_t* pk, uint64_t* pkp_or_null, int nthreads) {
  PROTO(OP_KEYGEN, affine); p.a0 = sk; p.a1 = gen_or_null; p.o0 = pk; p.o1 = pkp_or_null; run_batch(p, n, nthreads);
}
void ref_scalar_mul(int64_t n, int affine, const uint64_t* points, const uint64_t* k, uint64_t* out, int nthreads) {
  PROTO(OP_SMUL, affine); p.a0 = points; p.a1 = k; p.o0 = out; run_batch(p, n, nthreads);
}
