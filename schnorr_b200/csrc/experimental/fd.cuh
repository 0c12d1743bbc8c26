// F_q arithmetic on the FP64 pipe (DFMA) for sm_100a: 5 x 52-bit limbs, Montgomery radix 2^260.
//
// Why a second multiplier: the verify / sign kernels saturate the FMA-heavy pipe with IMAD.WIDE (ncu: 87 % active)
// while the FP64 pipe idles.  B200 issues DFMA at twice the IMAD.WIDE rate and the two pipes run concurrently
// (tools/imad_peak.cu modes 8/9: the wide-product rate is unchanged with 0.5 DFMA per product interleaved), so the
// Poseidon challenge hash (27 % of a verification's integer multiplies, 68 % of a signature's) is moved here and
// warps in their hash phase overlap with warps in their curve phase.  Replaces the arithmetic inside
// `dusk_poseidon::sponge::truncated::hash` as called from /root/reference/src/signatures.rs:127-134, 275-290.
//
// Product of two 52-bit limbs held in doubles (exact integers a, b < 2^52), two DFMA + one DADD:
//     h = fma_rz(a, b, 2^104)          = 2^104 + floor(ab / 2^52) * 2^52      (ulp of [2^104, 2^105) is 2^52)
//     l = fma_rz(a, b, 2^104 + 2^52 - h) = 2^52 + (ab mod 2^52)                 (exact)
// The bit patterns of h and l are  0x467<<52 | floor(ab / 2^52)  and  0x433<<52 | (ab mod 2^52): added as 64-bit
// INTEGERS into column accumulators, the exponent fields sum to a per-column constant that is pre-subtracted, so
// one IADD3 pair accumulates two partial products.  All exponent offsets are multiples of 2^52, hence invisible to
// the Montgomery factor m = -col * q^-1 mod 2^52 (q = 1 - 2^32 mod 2^52 makes q^-1 = 1 + 2^32: shifts and adds).
//
// Values are kept lazily reduced (< 2^257 ~ 4.4 q, never conditionally subtracted inside the permutation):
//     fd_mul: a b / 2^260 + q;  2^260 / q ~ 35.3 leaves the head-room.
// Host build: the same code with the DFMA pair emulated by a 128-bit product (tests/host_arith.cpp).
#pragma once
#include "../fq.cuh"
#include "constants_fd_gen.cuh"

#if defined(__CUDACC__)
#define SB_HDC __host__ __device__ constexpr
#else
#define SB_HDC constexpr
#endif

namespace sb200 {

struct fd {  // integer form: 52-bit limbs, l[0..3] < 2^52, l[4] < 2^52 (value < 2^260)
  uint64_t l[5];
};
struct fdd {  // operand form: the same limbs as doubles
  double d[5];
};

constexpr uint64_t FD_M52 = (1ull << 52) - 1;
constexpr uint64_t FD_B52 = 0x4330000000000000ull;   // bit pattern of 2^52
constexpr uint64_t FD_B104 = 0x4670000000000000ull;  // bit pattern of 2^104

SB_HD double fd_todbl(uint64_t x) {  // x < 2^52, exact
#if defined(__CUDA_ARCH__)
  return __dsub_rn(__longlong_as_double((long long)(x | FD_B52)), 0x1p+52);
#else
  return (double)x;
#endif
}
SB_HD fdd fd_todbl(const fd& a) {
  fdd r;
#pragma unroll
  for (int i = 0; i < 5; i++) r.d[i] = fd_todbl(a.l[i]);
  return r;
}

// bit patterns of the two halves of a * b
SB_HD void fd_prod(double a, double b, uint64_t& hi, uint64_t& lo) {
#if defined(__CUDA_ARCH__)
  double h = __fma_rz(a, b, 0x1p+104);
  double s = __dsub_rn(0x1.0000000000001p+104, h);  // 2^104 + 2^52 - h, exact
  double l = __fma_rz(a, b, s);
  hi = (uint64_t)__double_as_longlong(h);
  lo = (uint64_t)__double_as_longlong(l);
#else
  SB_COUNT(dfma, 2);
  unsigned __int128 p = (unsigned __int128)(uint64_t)a * (uint64_t)b;
  hi = FD_B104 + (uint64_t)(p >> 52);
  lo = FD_B52 + ((uint64_t)p & FD_M52);
#endif
}

// exponent-field totals a 5 x 5 limb product (or one Montgomery reduction: same shape) leaves in column k
SB_HDC uint64_t fd_off_mac(int k) {
  uint64_t s = 0;
  for (int i = 0; i < 5; i++)
    for (int j = 0; j < 5; j++) {
      if (i + j == k) s += FD_B52;
      if (i + j + 1 == k) s += FD_B104;
    }
  return s;
}
// ... and a squaring (off-diagonal products counted twice)
SB_HDC uint64_t fd_off_sqr(int k) {
  uint64_t s = 0;
  for (int i = 0; i < 5; i++) {
    for (int j = i + 1; j < 5; j++) {
      if (i + j == k) s += 2 * FD_B52;
      if (i + j + 1 == k) s += 2 * FD_B104;
    }
    if (2 * i == k) s += FD_B52;
    if (2 * i + 1 == k) s += FD_B104;
  }
  return s;
}

// col[0..9] += a * b   (row-wise: the high half of a_i b_j and the low half of a_i b_(j+1) share a column,
// so they enter with one three-input add)
SB_HD void fd_mac(uint64_t* col, const double* a, const double* b) {
#pragma unroll
  for (int i = 0; i < 5; i++) {
    uint64_t h, l, ph;
    fd_prod(a[i], b[0], ph, l);
    col[i] += l;
#pragma unroll
    for (int j = 1; j < 5; j++) {
      fd_prod(a[i], b[j], h, l);
      col[i + j] += ph + l;
      ph = h;
    }
    col[i + 5] += ph;
  }
}

SB_HD void fd_q_dbl(double* q) {
  const double c[5] = SB200_FD_Q_D_INIT;
#pragma unroll
  for (int i = 0; i < 5; i++) q[i] = c[i];
}

// Montgomery reduction of the 10 columns (all exponent offsets, including this function's own, already
// subtracted by the caller): returns (T + M q) / 2^260 with normalised limbs.  Columns 5..9 may hold signed
// limb-wise addends (see fd_mulc_add), hence the arithmetic shifts of the final carry pass.
SB_HD fd fd_reduce(uint64_t* col) {
  double q[5];
  fd_q_dbl(q);
#pragma unroll
  for (int i = 0; i < 5; i++) {
    uint64_t t = col[i];
    uint64_t m = (0ull - (t + (t << 32))) & FD_M52;  // -t * (1 + 2^32) mod 2^52
    double md = fd_todbl(m);
    uint64_t h, l, ph;
    fd_prod(md, q[0], ph, l);
    col[i] += l;  // now a multiple of 2^52
#pragma unroll
    for (int j = 1; j < 5; j++) {
      fd_prod(md, q[j], h, l);
      col[i + j] += ph + l;
      ph = h;
    }
    col[i + 5] += ph;
    col[i + 1] += col[i] >> 52;
  }
  fd r;
#pragma unroll
  for (int k = 0; k < 4; k++) {
    r.l[k] = col[5 + k] & FD_M52;
    col[6 + k] += (uint64_t)((int64_t)col[5 + k] >> 52);
  }
  r.l[4] = col[9];
  return r;
}

// ---- the three primitives the device calls out of line -------------------------------------------------------
// (a challenge hash is ~18 000 limb products: inlined it is hundreds of KB of SASS and stalls on instruction fetch,
// and even one out-of-line body per field operation -- 17 KB for the five-term dot product -- measured
// `stall_no_instruction` 1.07 per issue.  Three small bodies, ~7 KB together, stay resident; the ten column
// accumulators travel in registers, in and out in the same place.)
struct fdc {  // the ten column accumulators
  uint64_t c[10];
};

SB_HD fdc fd_mac_inl(fdc c, const fdd& a, const fdd& b) {  // c += a * b
  fd_mac(c.c, a.d, b.d);
  return c;
}

// columns of a^2, exponent offsets of the squaring AND of the following reduction already subtracted:
// 10 off-diagonal products accumulated once and doubled by the three-input add that also brings in the diagonal
// term (the odd offsets are halved into the accumulator seeds: all are multiples of 2^52)
SB_HD fdc fd_sqr_cols_inl(const fdd& a) {
  uint64_t o[10];
  fdc r;
#pragma unroll
  for (int k = 0; k < 10; k++) o[k] = 0ull - ((fd_off_sqr(k) + fd_off_mac(k)) >> 1);
#pragma unroll
  for (int i = 0; i < 4; i++) {
    uint64_t h, l, ph;
    fd_prod(a.d[i], a.d[i + 1], ph, l);
    o[2 * i + 1] += l;
#pragma unroll
    for (int j = i + 2; j < 5; j++) {
      fd_prod(a.d[i], a.d[j], h, l);
      o[i + j] += ph + l;
      ph = h;
    }
    o[i + 5] += ph;
  }
#pragma unroll
  for (int i = 0; i < 5; i++) {
    uint64_t h, l;
    fd_prod(a.d[i], a.d[i], h, l);
    r.c[2 * i] = o[2 * i] + o[2 * i] + l;
    r.c[2 * i + 1] = o[2 * i + 1] + o[2 * i + 1] + h;
  }
  return r;
}

SB_HD fd fd_reduce_inl(fdc c) { return fd_reduce(c.c); }

#if defined(__CUDACC__)
static __device__ __noinline__ fdc fd_mac_ool(fdc c, fdd a, fdd b) { return fd_mac_inl(c, a, b); }
static __device__ __noinline__ fdc fd_sqr_cols_ool(fdd a) { return fd_sqr_cols_inl(a); }
static __device__ __noinline__ fd fd_reduce_ool(fdc c) { return fd_reduce_inl(c); }
#endif
SB_HD fdc fd_mac_step(const fdc& c, const fdd& a, const fdd& b) {
#if defined(__CUDA_ARCH__)
  return fd_mac_ool(c, a, b);
#else
  return fd_mac_inl(c, a, b);
#endif
}
SB_HD fdc fd_sqr_cols(const fdd& a) {
#if defined(__CUDA_ARCH__)
  return fd_sqr_cols_ool(a);
#else
  return fd_sqr_cols_inl(a);
#endif
}
SB_HD fd fd_reduce_step(const fdc& c) {
#if defined(__CUDA_ARCH__)
  return fd_reduce_ool(c);
#else
  return fd_reduce_inl(c);
#endif
}

// accumulators seeded with minus the exponent offsets of `nprod` products and of the reduction that follows
SB_HD fdc fd_cols_init(int nprod) {
  fdc r;
#pragma unroll
  for (int k = 0; k < 10; k++) r.c[k] = 0ull - (uint64_t)(nprod + 1) * fd_off_mac(k);
  return r;
}
SB_HD fdd fd_ld(const double* p) {  // a constant in operand form
  fdd r;
#pragma unroll
  for (int k = 0; k < 5; k++) r.d[k] = p[k];
  return r;
}

// a * b / 2^260 (+ q slack)
SB_HD fd fd_mul(const fdd& a, const fdd& b) {
  SB_COUNT(fd_mul, 1);
  return fd_reduce_step(fd_mac_step(fd_cols_init(1), a, b));
}
SB_HD fd fd_sqr(const fdd& a) {
  SB_COUNT(fd_sqr, 1);
  return fd_reduce_step(fd_sqr_cols(a));
}

// cst * x / 2^260 + c', with c' = c - q when c >= 2^255 (one conditional subtraction keeps the running words of the
// sparse partial rounds below 2^257: each update adds at most (1 + 0.03) q and the subtraction removes q whenever
// the word exceeds 1.1 q, so the bound creeps by 0.03 q per round: < 4 q after 59 rounds).
// `cst` points at 5 doubles (a canonical constant in operand form).
SB_HD fd fd_mulc_add(const double* cst, const fdd& x, const fd& c) {
  SB_COUNT(fd_mul, 1);
  const uint64_t qu[5] = SB200_FD_Q_U_INIT;
  const uint64_t mask = (c.l[4] >> 47) ? ~0ull : 0ull;
  fdc col = fd_cols_init(1);
#pragma unroll
  for (int k = 0; k < 5; k++) col.c[5 + k] += c.l[k] - (qu[k] & mask);
  return fd_reduce_step(fd_mac_step(col, x, fd_ld(cst)));
}

// sum_j cst[j] * s_j / 2^260 + add   (one reduction for five products; `add` = 5 integer limbs or nullptr)
SB_HD fd fd_dot5(const double* cst, const uint64_t* add, const fdd& s0, const fdd& s1, const fdd& s2, const fdd& s3,
                 const fdd& s4) {
  SB_COUNT(fd_dot5, 1);
  fdc col = fd_cols_init(5);
  if (add) {
#pragma unroll
    for (int k = 0; k < 5; k++) col.c[5 + k] += add[k];
  }
  col = fd_mac_step(col, s0, fd_ld(cst));
  col = fd_mac_step(col, s1, fd_ld(cst + 5));
  col = fd_mac_step(col, s2, fd_ld(cst + 10));
  col = fd_mac_step(col, s3, fd_ld(cst + 15));
  col = fd_mac_step(col, s4, fd_ld(cst + 20));
  return fd_reduce_step(col);
}

// a + b with carry normalisation (values stay far below 2^260)
SB_HD fd fd_add(const fd& a, const uint64_t* b) {
  fd r;
  uint64_t c = 0;
#pragma unroll
  for (int k = 0; k < 4; k++) {
    uint64_t t = a.l[k] + b[k] + c;
    r.l[k] = t & FD_M52;
    c = t >> 52;
  }
  r.l[4] = a.l[4] + b[4] + c;
  return r;
}

// ------------------------------------------------------------------------------------------------
// Two-lane ("paired") primitives: the same operations on two independent elements in one call.  A hash warp runs
// alone on its pipes (the other warps of its sub-partition are curve warps), so nothing hides the latency of the
// DFMA -> DADD -> DFMA -> IADD3 chains and of the five serial rows of a reduction: a single stream measured
// IPC 0.17 with 35-50 % `wait` stalls.  Two independent streams in the same basic block let ptxas interleave them.
// ------------------------------------------------------------------------------------------------
struct fd2 { fd a, b; };
struct fdd2 { fdd a, b; };
struct fdc2 { fdc a, b; };

SB_HD fdd2 fd_todbl2(const fd2& x) { return {fd_todbl(x.a), fd_todbl(x.b)}; }
SB_HD fdc2 fd_cols_init2(int nprod) { return {fd_cols_init(nprod), fd_cols_init(nprod)}; }

SB_HD fdc2 fd_macc2_inl(fdc2 c, const fdd2& x, const double* cst) {  // c += x * constant (same constant for both)
  fdd y = fd_ld(cst);
  fd_mac(c.a.c, x.a.d, y.d);
  fd_mac(c.b.c, x.b.d, y.d);
  return c;
}
SB_HD fdc2 fd_macv2_inl(fdc2 c, const fdd2& x, const fdd2& y) {  // c += x * y
  fd_mac(c.a.c, x.a.d, y.a.d);
  fd_mac(c.b.c, x.b.d, y.b.d);
  return c;
}
SB_HD fdc2 fd_sqr_cols2_inl(const fdd2& x) { return {fd_sqr_cols_inl(x.a), fd_sqr_cols_inl(x.b)}; }
SB_HD fd2 fd_reduce2_inl(fdc2 c) {
  // the two reductions interleaved row by row (each row's Montgomery factor depends on the previous row)
  double q[5];
  fd_q_dbl(q);
  uint64_t* col[2] = {c.a.c, c.b.c};
#pragma unroll
  for (int i = 0; i < 5; i++) {
#pragma unroll
    for (int w = 0; w < 2; w++) {
      uint64_t t = col[w][i];
      uint64_t m = (0ull - (t + (t << 32))) & FD_M52;
      double md = fd_todbl(m);
      uint64_t h, l, ph;
      fd_prod(md, q[0], ph, l);
      col[w][i] += l;
#pragma unroll
      for (int j = 1; j < 5; j++) {
        fd_prod(md, q[j], h, l);
        col[w][i + j] += ph + l;
        ph = h;
      }
      col[w][i + 5] += ph;
      col[w][i + 1] += col[w][i] >> 52;
    }
  }
  fd2 r;
  fd* rr[2] = {&r.a, &r.b};
#pragma unroll
  for (int w = 0; w < 2; w++) {
#pragma unroll
    for (int k = 0; k < 4; k++) {
      rr[w]->l[k] = col[w][5 + k] & FD_M52;
      col[w][6 + k] += (uint64_t)((int64_t)col[w][5 + k] >> 52);
    }
    rr[w]->l[4] = col[w][9];
  }
  return r;
}
#if defined(__CUDACC__)
static __device__ __noinline__ fdc2 fd_macc2_ool(fdc2 c, fdd2 x, const double* cst) { return fd_macc2_inl(c, x, cst); }
static __device__ __noinline__ fdc2 fd_macv2_ool(fdc2 c, fdd2 x, fdd2 y) { return fd_macv2_inl(c, x, y); }
static __device__ __noinline__ fdc2 fd_sqr_cols2_ool(fdd2 x) { return fd_sqr_cols2_inl(x); }
static __device__ __noinline__ fd2 fd_reduce2_ool(fdc2 c) { return fd_reduce2_inl(c); }
#endif
SB_HD fdc2 fd_macc2(const fdc2& c, const fdd2& x, const double* cst) {
#if defined(__CUDA_ARCH__)
  return fd_macc2_ool(c, x, cst);
#else
  return fd_macc2_inl(c, x, cst);
#endif
}
SB_HD fdc2 fd_macv2(const fdc2& c, const fdd2& x, const fdd2& y) {
#if defined(__CUDA_ARCH__)
  return fd_macv2_ool(c, x, y);
#else
  return fd_macv2_inl(c, x, y);
#endif
}
SB_HD fdc2 fd_sqr_cols2(const fdd2& x) {
#if defined(__CUDA_ARCH__)
  return fd_sqr_cols2_ool(x);
#else
  return fd_sqr_cols2_inl(x);
#endif
}
SB_HD fd2 fd_reduce2(const fdc2& c) {
#if defined(__CUDA_ARCH__)
  return fd_reduce2_ool(c);
#else
  return fd_reduce2_inl(c);
#endif
}
// addend (+ the conditional subtraction of q, see fd_mulc_add) into the upper columns
SB_HD void fd_cols_add(fdc& col, const uint64_t* add) {
#pragma unroll
  for (int k = 0; k < 5; k++) col.c[5 + k] += add[k];
}
SB_HD void fd_cols_add_csub(fdc& col, const fd& c) {
  const uint64_t qu[5] = SB200_FD_Q_U_INIT;
  const uint64_t mask = (c.l[4] >> 47) ? ~0ull : 0ull;
#pragma unroll
  for (int k = 0; k < 5; k++) col.c[5 + k] += c.l[k] - (qu[k] & mask);
}

// ------------------------------------------------------------------------------------------------
// Memory-operand form of the same primitives.  A "slot" holds one field element in both forms: doubles at
// p[k * ls] (k < 5) and integer limbs at ((uint64_t*)p)[(5 + k) * ls]; ls = 32 for lane-strided shared memory
// (conflict-free: consecutive lanes, consecutive 8-byte words), 1 for a thread-private array.  Operands and
// results stay in memory, only the ten column accumulators travel in registers -- in and out of every call in the
// same registers -- so a call site is a few address computations instead of 30-40 register moves (which ptxas
// emits half as IMAD.MOV on the FMA-heavy pipe the curve warps are saturating) and the permutation's code shrinks
// from ~60 KB to a few KB (the 32 KB L1.5 instruction cache has to hold the curve loop as well).
// ------------------------------------------------------------------------------------------------
SB_HD const uint64_t* fd_slot_ints(const double* p, int ls) { return reinterpret_cast<const uint64_t*>(p) + 5 * ls; }

SB_HD fdc fd_mac_p_inl(fdc c, const double* a, int as, const double* b, int bs) {
  fdd x, y;
#pragma unroll
  for (int k = 0; k < 5; k++) {
    x.d[k] = a[k * as];
    y.d[k] = b[k * bs];
  }
  fd_mac(c.c, x.d, y.d);
  return c;
}
SB_HD fdc fd_sqr_p_inl(const double* a, int as) {
  fdd x;
#pragma unroll
  for (int k = 0; k < 5; k++) x.d[k] = a[k * as];
  return fd_sqr_cols_inl(x);
}
SB_HD void fd_reduce_p_inl(fdc c, double* dst, int ds) {
  fd r = fd_reduce(c.c);
  uint64_t* u = reinterpret_cast<uint64_t*>(dst) + 5 * ds;
#pragma unroll
  for (int k = 0; k < 5; k++) {
    dst[k * ds] = fd_todbl(r.l[k]);
    u[k * ds] = r.l[k];
  }
}
// accumulators for nprod products + reduction, plus an addend (integer limbs at add[k * as], or nullptr); with
// csub the addend is reduced by q when it is >= 2^255 (see fd_mulc_add)
SB_HD fdc fd_start_p_inl(int nprod, const uint64_t* add, int as, bool csub) {
  fdc r;
  const uint64_t mult = (uint64_t)nprod + 1;
#pragma unroll
  for (int k = 0; k < 10; k++) r.c[k] = 0ull - mult * fd_off_mac(k);
  if (add) {
    const uint64_t qu[5] = SB200_FD_Q_U_INIT;
    uint64_t v[5];
#pragma unroll
    for (int k = 0; k < 5; k++) v[k] = add[k * as];
    const uint64_t mask = (csub && (v[4] >> 47)) ? ~0ull : 0ull;
#pragma unroll
    for (int k = 0; k < 5; k++) r.c[5 + k] += v[k] - (qu[k] & mask);
  }
  return r;
}
#if defined(__CUDACC__)
static __device__ __noinline__ fdc fd_mac_p_ool(fdc c, const double* a, int as, const double* b, int bs) { return fd_mac_p_inl(c, a, as, b, bs); }
static __device__ __noinline__ fdc fd_sqr_p_ool(const double* a, int as) { return fd_sqr_p_inl(a, as); }
static __device__ __noinline__ void fd_reduce_p_ool(fdc c, double* dst, int ds) { fd_reduce_p_inl(c, dst, ds); }
static __device__ __noinline__ fdc fd_start_p_ool(int nprod, const uint64_t* add, int as, bool csub) { return fd_start_p_inl(nprod, add, as, csub); }
#endif
SB_HD fdc fd_mac_p(const fdc& c, const double* a, int as, const double* b, int bs) {
#if defined(__CUDA_ARCH__)
  return fd_mac_p_ool(c, a, as, b, bs);
#else
  return fd_mac_p_inl(c, a, as, b, bs);
#endif
}
SB_HD fdc fd_sqr_p(const double* a, int as) {
#if defined(__CUDA_ARCH__)
  return fd_sqr_p_ool(a, as);
#else
  return fd_sqr_p_inl(a, as);
#endif
}
SB_HD void fd_reduce_p(const fdc& c, double* dst, int ds) {
#if defined(__CUDA_ARCH__)
  fd_reduce_p_ool(c, dst, ds);
#else
  fd_reduce_p_inl(c, dst, ds);
#endif
}
SB_HD fdc fd_start_p(int nprod, const uint64_t* add, int as, bool csub) {
#if defined(__CUDA_ARCH__)
  return fd_start_p_ool(nprod, add, as, csub);
#else
  return fd_start_p_inl(nprod, add, as, csub);
#endif
}
// store an element (integer form) into a slot in both forms
SB_HD void fd_slot_store(double* dst, int ds, const fd& a) {
  uint64_t* u = reinterpret_cast<uint64_t*>(dst) + 5 * ds;
#pragma unroll
  for (int k = 0; k < 5; k++) {
    dst[k * ds] = fd_todbl(a.l[k]);
    u[k * ds] = a.l[k];
  }
}
SB_HD fd fd_slot_load(const double* src, int ss) {
  const uint64_t* u = fd_slot_ints(src, ss);
  fd r;
#pragma unroll
  for (int k = 0; k < 5; k++) r.l[k] = u[k * ss];
  return r;
}

// ------------------------------------------------------------------------------------------------
// conversions: 8 x u32 Montgomery-2^256 form (the ABI's field elements)  <->  5 x 52 Montgomery-2^260 form
// ------------------------------------------------------------------------------------------------
SB_HD fd fd_split(const uint32_t* w) {  // plain re-slicing of a 256-bit integer
  uint64_t v[4];
#pragma unroll
  for (int i = 0; i < 4; i++) v[i] = (uint64_t)w[2 * i] | ((uint64_t)w[2 * i + 1] << 32);
  fd r;
  r.l[0] = v[0] & FD_M52;
  r.l[1] = ((v[0] >> 52) | (v[1] << 12)) & FD_M52;
  r.l[2] = ((v[1] >> 40) | (v[2] << 24)) & FD_M52;
  r.l[3] = ((v[2] >> 28) | (v[3] << 36)) & FD_M52;
  r.l[4] = v[3] >> 16;
  return r;
}
SB_HD void fd_join(const fd& a, uint32_t* w) {  // limbs normalised, value < 2^256
  uint64_t v[4];
  v[0] = a.l[0] | (a.l[1] << 52);
  v[1] = (a.l[1] >> 12) | (a.l[2] << 40);
  v[2] = (a.l[2] >> 24) | (a.l[3] << 28);
  v[3] = (a.l[3] >> 36) | (a.l[4] << 16);
#pragma unroll
  for (int i = 0; i < 4; i++) {
    w[2 * i] = (uint32_t)v[i];
    w[2 * i + 1] = (uint32_t)(v[i] >> 32);
  }
}

// x * 2^256 (8 x u32)  ->  x * 2^260 (5 x 52):  mont260(v, 2^264) = v * 2^4
SB_HD fd fd_from_fq(const fq& a) {
  const fdd kin = {SB200_FD_KIN_D_INIT};
  return fd_mul(fd_todbl(fd_split(a.v)), kin);
}

// x * 2^260 (lazily reduced)  ->  canonical integer x as 8 x u32:  mont260(v, 1) <= q, one conditional subtraction
SB_HD void fd_to_canonical(const fd& a, uint32_t* w) {
  fdd one;
  one.d[0] = 1.0;
#pragma unroll
  for (int i = 1; i < 5; i++) one.d[i] = 0.0;
  fd r = fd_mul(fd_todbl(a), one);
  fd_join(r, w);
  cond_sub_p<FqP>(w);  // [0, q] -> [0, q)
}

}  // namespace sb200
