// libschnorr_b200.so: sm_100a kernels + the C ABI declared in include/schnorr_b200.h.
//
// One tuple per thread, 128 threads per CTA.  Every kernel is integer-multiply bound (~3e5
// IMAD.WIDE per verification against ~300 bytes of input), so the inputs are read straight from the
// tuple-major arrays with two 128-bit loads per field element (every 32-byte sector fully used) and
// the fixed-base comb tables stay L2-resident.  Host buffers are pipelined through two streams per
// device in chunks (H2D / kernel / D2H overlap); tuples shard by index over the context's devices
// with no inter-GPU traffic.
#include <cuda_runtime.h>

#include <algorithm>
#include <atomic>
#include <cstdio>
#include <cstring>
#include <mutex>
#include <string>
#include <vector>

#include "../../include/schnorr_b200.h"
#include "wire.cuh"

using namespace sb200;

namespace {

constexpr int TPB = 128;
#ifndef SB_CHALLENGE_FD
#define SB_CHALLENGE_FD 0  // the stand-alone challenge kernel on the FP64 pipe instead of IMAD.WIDE: measured equal (182.94 vs 182.91 ms per 2^22 verifications)
#endif
#ifndef SB_MIN_CTAS
#define SB_MIN_CTAS 4
#endif
// CTAs per SM the register allocator must allow: 4 (128 registers/thread) everywhere except the two kernels that
// keep a second point table live, which measure faster at 3 (168 registers, fewer spills).  [A/B timed on B200]
#ifndef SB_VERIFY_CTAS
#define SB_VERIFY_CTAS SB_MIN_CTAS
#endif
constexpr int min_ctas(int op) { return (op == 0 || op == 19) ? SB_VERIFY_CTAS : (op == 25 || op == 27) ? SB_MIN_CTAS - 1 : (op == 1 || op == 2) ? SB_MIN_CTAS - 1 : SB_MIN_CTAS; }
constexpr int MAX_IN = 6, MAX_OUT = 4;
constexpr int64_t CHUNK = 1 << 18;  // tuples per pipeline stage

enum Op : int {
  OP_VERIFY = 0, OP_VERIFY_DOUBLE, OP_VERIFY_VARGEN, OP_SIGN, OP_SIGN_DOUBLE, OP_SIGN_VARGEN,
  OP_KEYGEN, OP_KEYGEN_DOUBLE, OP_KEYGEN_VARGEN, OP_DBG_FQ, OP_DBG_FR_MUL, OP_DBG_HADES, OP_DBG_SMUL,
  OP_DECOMPRESS, OP_COMPRESS, OP_FROM_WIDE, OP_VERIFY_BYTES, OP_SIGN_BYTES, OP_CHALLENGE, OP_VERIFY_EC,
  OP_VERIFY_DOUBLE_BYTES, OP_VERIFY_VARGEN_BYTES, OP_SIGN_DOUBLE_BYTES, OP_SIGN_VARGEN_BYTES,
  OP_CHALLENGE_DOUBLE, OP_VERIFY_DOUBLE_EC, OP_CHALLENGE_VARGEN, OP_VERIFY_VARGEN_EC
};

struct KArgs {
  int64_t n;
  uint32_t flags;
  int aux;
  const uint32_t* in[MAX_IN];
  uint32_t* out[MAX_OUT];
  uint32_t* bitmap;
  const uint32_t* combG;
  const uint32_t* combGp;
  struct WsState* ws;
  int nsm;
  pniels* ec_scratch;  // window tables of k_verify_ec_p: EC_P_CTAS * nsm * TPB threads x 18 entries (per stream)
};

// scheduling state of one warp-specialised launch (k_verify_ws): the next-tile counter, reset before every launch,
// and one monotonic CTA-arrival counter per SM that rotates the hash-warp position among an SM's resident CTAs
struct WsState {
  unsigned tile;
  unsigned pad[31];
  unsigned sm_slot[512];
};

__device__ __forceinline__ fq ldg_fq(const uint32_t* p) {
  const uint4* q = reinterpret_cast<const uint4*>(p);
  uint4 a = __ldg(q), b = __ldg(q + 1);
  fq r = {{a.x, a.y, a.z, a.w, b.x, b.y, b.z, b.w}};
  return r;
}
__device__ __forceinline__ void ldg_scalar(const uint32_t* p, uint32_t* k) {
  fq t = ldg_fq(p);
#pragma unroll
  for (int i = 0; i < 8; i++) k[i] = t.v[i];
}
__device__ __forceinline__ void stg8(uint32_t* p, const uint32_t* v) {
  uint4* q = reinterpret_cast<uint4*>(p);
  q[0] = make_uint4(v[0], v[1], v[2], v[3]);
  q[1] = make_uint4(v[4], v[5], v[6], v[7]);
}
__device__ __forceinline__ point_in ldg_point(const uint32_t* base, int64_t i, bool affine) {
  point_in p;
  const uint32_t* q = base + i * (affine ? 16 : 24);
  p.U = ldg_fq(q);
  p.V = ldg_fq(q + 8);
  p.Z = affine ? fq_one() : ldg_fq(q + 16);
  p.affine = affine;
  return p;
}
__device__ __forceinline__ void stg_point(uint32_t* base, int64_t i, const fq& u, const fq& v) {
  stg8(base + i * 16, u.v);
  stg8(base + i * 16 + 8, v.v);
}

template <int OP>
__global__ void __launch_bounds__(TPB, min_ctas(OP)) k_run(const KArgs a) {
  int64_t i = (int64_t)blockIdx.x * TPB + threadIdx.x;
  const bool active = i < a.n;
  if (!active) i = a.n - 1;  // idle lanes redo the last tuple so the warp stays converged
  const bool aff = (a.flags & SB200_POINTS_AFFINE) != 0;
  uint32_t c[8];

  if (OP == OP_VERIFY || OP == OP_VERIFY_DOUBLE || OP == OP_VERIFY_VARGEN) {
    bool ok;
    if (OP == OP_VERIFY) {  // in: pk, u, R, m
      uint32_t u[8];
      ldg_scalar(a.in[1] + i * 8, u);
      ok = verify_core(ldg_point(a.in[0], i, aff), u, ldg_point(a.in[2], i, aff), ldg_fq(a.in[3] + i * 8), a.combG, c);
    } else if (OP == OP_VERIFY_DOUBLE) {  // in: pk, pk', u, R, R', m
      uint32_t u[8];
      ldg_scalar(a.in[2] + i * 8, u);
      ok = verify_double_core(ldg_point(a.in[0], i, aff), ldg_point(a.in[1], i, aff), u, ldg_point(a.in[3], i, aff),
                              ldg_point(a.in[4], i, aff), ldg_fq(a.in[5] + i * 8), a.combG, a.combGp, c);
    } else {  // in: pk, gen, u, R, m
      uint32_t u[8];
      ldg_scalar(a.in[2] + i * 8, u);
      ok = verify_vargen_core(ldg_point(a.in[0], i, aff), ldg_point(a.in[1], i, aff), u, ldg_point(a.in[3], i, aff),
                              ldg_fq(a.in[4] + i * 8), c);
    }
    unsigned word = __ballot_sync(0xffffffffu, ok && active);
    if ((threadIdx.x & 31) == 0 && active) a.bitmap[i >> 5] = word;
    if (a.out[0] && active) stg8(a.out[0] + i * 8, c);
    return;
  }

  // single-key verification as two launches (SB_VERIFY_SPLIT): all warps of an SM are then in the same phase, and each
  // phase's code (hash: 26 KB, curve: 19 KB) fits the 32 KB L1.5 instruction cache by itself
  if (OP == OP_CHALLENGE) {  // in: -, -, R, m -> out0: c
#if SB_CHALLENGE_FD
    {  // the challenge kernel on the FP64 pipe (hades_fd.cuh): nothing else competes for issue slots here
      fq ru, rv;
      point_to_affine(ldg_point(a.in[2], i, aff), ru, rv);
      challenge3_fd(ru, rv, ldg_fq(a.in[3] + i * 8), c);
    }
#else
    verify_hash_core(ldg_point(a.in[2], i, aff), ldg_fq(a.in[3] + i * 8), c);
#endif
    if (active) stg8(a.out[0] + i * 8, c);
    return;
  }
  if (OP == OP_VERIFY_EC) {  // in: pk, u, R, - ; out0 (as input): c -> bitmap
    uint32_t u[8];
    ldg_scalar(a.in[1] + i * 8, u);
    ldg_scalar(a.out[0] + i * 8, c);
    bool ok = verify_ec(ldg_point(a.in[0], i, aff), u, ldg_point(a.in[2], i, aff), c, a.combG);
    unsigned word = __ballot_sync(0xffffffffu, ok && active);
    if ((threadIdx.x & 31) == 0 && active) a.bitmap[i >> 5] = word;
    return;
  }

  if (OP == OP_CHALLENGE_VARGEN) {  // in: -, -, -, R, m -> out0: c
    verify_hash_core(ldg_point(a.in[3], i, aff), ldg_fq(a.in[4] + i * 8), c);
    if (active) stg8(a.out[0] + i * 8, c);
    return;
  }
  if (OP == OP_VERIFY_VARGEN_EC) {  // in: pk, gen, u, R, - ; out0 (as input): c -> bitmap
    uint32_t u[8];
    ldg_scalar(a.in[2] + i * 8, u);
    ldg_scalar(a.out[0] + i * 8, c);
    bool ok = verify_vargen_ec(ldg_point(a.in[0], i, aff), ldg_point(a.in[1], i, aff), u, ldg_point(a.in[3], i, aff), c);
    unsigned word = __ballot_sync(0xffffffffu, ok && active);
    if ((threadIdx.x & 31) == 0 && active) a.bitmap[i >> 5] = word;
    return;
  }
  if (OP == OP_CHALLENGE_DOUBLE) {  // in: -, -, -, R, R', m -> out0: c
    verify_double_hash_core(ldg_point(a.in[3], i, aff), ldg_point(a.in[4], i, aff), ldg_fq(a.in[5] + i * 8), c);
    if (active) stg8(a.out[0] + i * 8, c);
    return;
  }
  if (OP == OP_VERIFY_DOUBLE_EC) {  // in: pk, pk', u, R, R', - ; out0 (as input): c -> bitmap
    uint32_t u[8];
    ldg_scalar(a.in[2] + i * 8, u);
    ldg_scalar(a.out[0] + i * 8, c);
    bool ok = verify_double_ec(ldg_point(a.in[0], i, aff), ldg_point(a.in[1], i, aff), u, ldg_point(a.in[3], i, aff),
                               ldg_point(a.in[4], i, aff), c, a.combG, a.combGp);
    unsigned word = __ballot_sync(0xffffffffu, ok && active);
    if ((threadIdx.x & 31) == 0 && active) a.bitmap[i >> 5] = word;
    return;
  }

  if (OP == OP_SIGN || OP == OP_SIGN_DOUBLE || OP == OP_SIGN_VARGEN) {
    // out: u, R, (R'), c
    uint32_t sk[8], nonce[8], u[8];
    fq Ru, Rv;
    ldg_scalar(a.in[0] + i * 8, sk);
    if (OP == OP_SIGN) {  // in: sk, m, nonce
      ldg_scalar(a.in[2] + i * 8, nonce);
      sign_core(sk, nonce, ldg_fq(a.in[1] + i * 8), a.combG, u, Ru, Rv, c);
    } else if (OP == OP_SIGN_DOUBLE) {
      fq Rpu, Rpv;
      ldg_scalar(a.in[2] + i * 8, nonce);
      sign_double_core(sk, nonce, ldg_fq(a.in[1] + i * 8), a.combG, a.combGp, u, Ru, Rv, Rpu, Rpv, c);
      if (active) stg_point(a.out[2], i, Rpu, Rpv);
    } else {  // in: sk, gen, m, nonce
      ldg_scalar(a.in[3] + i * 8, nonce);
      sign_vargen_core(sk, ldg_point(a.in[1], i, aff), nonce, ldg_fq(a.in[2] + i * 8), u, Ru, Rv, c);
    }
    if (active) {
      stg8(a.out[0] + i * 8, u);
      stg_point(a.out[1], i, Ru, Rv);
      if (a.out[3]) stg8(a.out[3] + i * 8, c);
    }
    return;
  }

  if (OP == OP_KEYGEN || OP == OP_KEYGEN_DOUBLE || OP == OP_KEYGEN_VARGEN) {
    uint32_t sk[8];
    fq u, v;
    ldg_scalar(a.in[0] + i * 8, sk);
    if (OP == OP_KEYGEN_VARGEN) {
      ext_to_affine(var_base_mul(ldg_point(a.in[1], i, aff), sk), u, v);
    } else {
      ext_to_affine(fixed_base_mul(a.combG, sk), u, v);
    }
    if (active) stg_point(a.out[0], i, u, v);
    if (OP == OP_KEYGEN_DOUBLE) {
      ext_to_affine(fixed_base_mul(a.combGp, sk), u, v);
      if (active) stg_point(a.out[1], i, u, v);
    }
    return;
  }

  if (OP == OP_DBG_FQ) {
    fq x = ldg_fq(a.in[0] + i * 8), y = a.in[1] ? ldg_fq(a.in[1] + i * 8) : fq_zero(), r;
    switch (a.aux) {
      case 0: r = fq_mul(x, y); break;
      case 1: r = fq_add(x, y); break;
      case 2: r = fq_sub(x, y); break;
      case 3: r = fq_inv(x); break;
      case 4: r = fq_sqr(x); break;
      case 5: r = fq_to_mont(x); break;
      default: r = fq_from_mont(x); break;
    }
    if (active) stg8(a.out[0] + i * 8, r.v);
    return;
  }
  if (OP == OP_DBG_FR_MUL) {
    fr x, y;
    ldg_scalar(a.in[0] + i * 8, x.v);
    ldg_scalar(a.in[1] + i * 8, y.v);
    fr r = fr_mul(x, y);
    if (active) stg8(a.out[0] + i * 8, r.v);
    return;
  }
  if (OP == OP_DBG_HADES) {
    fq s[5];
#pragma unroll
    for (int k = 0; k < 5; k++) s[k] = ldg_fq(a.in[0] + i * 40 + k * 8);
    if (a.aux == 2) {  // FP64-pipe permutation (hades_fd.cuh); same Montgomery-2^256 words in and out
      fd t[5];
#pragma unroll 1
      for (int k = 0; k < 5; k++) t[k] = fd_from_fq(s[k]);
      hades_perm_fd(t);
#pragma unroll 1
      for (int k = 0; k < 5; k++) {
        fd_to_canonical(t[k], s[k].v);
        s[k] = fq_to_mont(s[k]);
      }
    } else if (a.aux) hades_perm_dense(s); else hades_perm(s);
    if (active) {
#pragma unroll
      for (int k = 0; k < 5; k++) stg8(a.out[0] + i * 40 + k * 8, s[k].v);
    }
    return;
  }
  if (OP == OP_DBG_SMUL) {  // in: points (base 2 only), k
    uint32_t k[8];
    fq u, v;
    ldg_scalar(a.in[1] + i * 8, k);
    if (a.aux == 2) ext_to_affine(var_base_mul(ldg_point(a.in[0], i, aff), k), u, v);
    else ext_to_affine(fixed_base_mul(a.aux == 0 ? a.combG : a.combGp, k), u, v);
    if (active) stg_point(a.out[0], i, u, v);
    return;
  }
  if (OP == OP_DECOMPRESS) {  // in: bytes32 -> out0: affine points, bitmap: ok
    uint32_t b[8];
    fq u, v;
    ldg_scalar(a.in[0] + i * 8, b);
    bool ok = point_decompress(b, u, v);
    unsigned word = __ballot_sync(0xffffffffu, ok && active);
    if ((threadIdx.x & 31) == 0 && active) a.bitmap[i >> 5] = word;
    if (active) stg_point(a.out[0], i, u, v);
    return;
  }
  if (OP == OP_COMPRESS) {  // in: points -> out0: bytes32
    point_in p = ldg_point(a.in[0], i, aff);
    fq u, v;
    point_to_affine(p, u, v);
    uint32_t b[8];
    point_compress(u, v, b);
    if (active) stg8(a.out[0] + i * 8, b);
    return;
  }
  if (OP == OP_FROM_WIDE) {  // in: 64-byte draws -> out0: scalar (aux 0: mod r canonical, aux 1: mod q Montgomery)
    uint32_t w[16], r[8];
    ldg_scalar(a.in[0] + i * 16, w);
    ldg_scalar(a.in[0] + i * 16 + 8, w + 8);
    if (a.aux == 0) {
      fr_from_wide(w, r);
    } else {
      fq x = fq_from_wide(w);
#pragma unroll
      for (int k = 0; k < 8; k++) r[k] = x.v[k];
    }
    if (active) stg8(a.out[0] + i * 8, r);
    return;
  }
  if (OP == OP_VERIFY_BYTES) {  // in: pk32, sig64, msg32 -> bitmap: verdicts, out0 (as bitmap): invalid
    uint32_t pk[8], sig[16], m[8];
    ldg_scalar(a.in[0] + i * 8, pk);
    ldg_scalar(a.in[1] + i * 16, sig);
    ldg_scalar(a.in[1] + i * 16 + 8, sig + 8);
    ldg_scalar(a.in[2] + i * 8, m);
    bool invalid;
    bool ok = verify_bytes_core(pk, sig, m, a.combG, invalid);
    unsigned word = __ballot_sync(0xffffffffu, ok && active), inv = __ballot_sync(0xffffffffu, invalid && active);
    if ((threadIdx.x & 31) == 0 && active) {
      a.bitmap[i >> 5] = word;
      if (a.out[0]) a.out[0][i >> 5] = inv;
    }
    return;
  }
  if (OP == OP_VERIFY_DOUBLE_BYTES || OP == OP_VERIFY_VARGEN_BYTES) {  // in: pk64, sig96|sig64, msg32 -> bitmap, out0: invalid
    constexpr int SW = OP == OP_VERIFY_DOUBLE_BYTES ? 24 : 16;
    uint32_t pk[16], sig[SW], m[8];
#pragma unroll
    for (int k = 0; k < 2; k++) ldg_scalar(a.in[0] + i * 16 + 8 * k, pk + 8 * k);
#pragma unroll
    for (int k = 0; k < SW / 8; k++) ldg_scalar(a.in[1] + i * SW + 8 * k, sig + 8 * k);
    ldg_scalar(a.in[2] + i * 8, m);
    bool invalid, ok;
    if (OP == OP_VERIFY_DOUBLE_BYTES) ok = verify_double_bytes_core(pk, sig, m, a.combG, a.combGp, invalid);
    else ok = verify_vargen_bytes_core(pk, sig, m, invalid);
    unsigned word = __ballot_sync(0xffffffffu, ok && active), inv = __ballot_sync(0xffffffffu, invalid && active);
    if ((threadIdx.x & 31) == 0 && active) {
      a.bitmap[i >> 5] = word;
      if (a.out[0]) a.out[0][i >> 5] = inv;
    }
    return;
  }
  if (OP == OP_SIGN_DOUBLE_BYTES) {  // in: sk32, msg32, nonce32 -> out0: sig96 = u || R || R'
    uint32_t sk[8], nonce[8], mb[8], sig[24];
    ldg_scalar(a.in[0] + i * 8, sk);
    ldg_scalar(a.in[1] + i * 8, mb);
    ldg_scalar(a.in[2] + i * 8, nonce);
    sign_double_bytes_core(sk, mb, nonce, a.combG, a.combGp, sig);
    if (active) {
#pragma unroll
      for (int k = 0; k < 3; k++) stg8(a.out[0] + i * 24 + 8 * k, sig + 8 * k);
    }
    return;
  }
  if (OP == OP_SIGN_VARGEN_BYTES) {  // in: sk64 (sk || generator), msg32, nonce32 -> out0: sig64, bitmap: generator decoded
    uint32_t sk[16], nonce[8], mb[8], sig[16];
    ldg_scalar(a.in[0] + i * 16, sk);
    ldg_scalar(a.in[0] + i * 16 + 8, sk + 8);
    ldg_scalar(a.in[1] + i * 8, mb);
    ldg_scalar(a.in[2] + i * 8, nonce);
    bool ok = sign_vargen_bytes_core(sk, mb, nonce, sig);
    unsigned word = __ballot_sync(0xffffffffu, ok && active);
    if ((threadIdx.x & 31) == 0 && active && a.bitmap) a.bitmap[i >> 5] = word;
    if (active) {
      stg8(a.out[0] + i * 16, sig);
      stg8(a.out[0] + i * 16 + 8, sig + 8);
    }
    return;
  }
  if (OP == OP_SIGN_BYTES) {  // in: sk32, msg32 (canonical), nonce32 -> out0: sig64 = u || compress(R)
    uint32_t sk[8], nonce[8], mb[8], u[8], sig[16];
    fq Ru, Rv, mc;
    ldg_scalar(a.in[0] + i * 8, sk);
    ldg_scalar(a.in[1] + i * 8, mb);
    ldg_scalar(a.in[2] + i * 8, nonce);
#pragma unroll
    for (int k = 0; k < 8; k++) mc.v[k] = mb[k];
    sign_core(sk, nonce, fq_to_mont(mc), a.combG, u, Ru, Rv, c);
#pragma unroll
    for (int k = 0; k < 8; k++) sig[k] = u[k];
    point_compress(Ru, Rv, sig + 8);
    if (active) {
      stg8(a.out[0] + i * 16, sig);
      stg8(a.out[0] + i * 16 + 8, sig + 8);
    }
    return;
  }
}

// ------------------------------------------------------------------------------------------------
// Warp-specialised single-key verification: one persistent 512-thread CTA per SM = 8 hash warps + 8 curve warps,
// i.e. on every SM sub-partition (warp w runs on sub-partition w % 4) 2 hash warps beside 2 curve warps.
//
// The curve half of a verification (c*PK + u*G == R) saturates the FMA-heavy pipe with IMAD.WIDE while using ~40 %
// of the issue slots; the Poseidon challenge on the FP64 pipe (hades_fd.cuh) needs issue slots and the FP64 / ALU
// pipes only.  Keeping both kinds of warp on every sub-partition at all times overlaps the two: the hash warps
// compute the challenges of the CTA's NEXT tile into shared memory while the curve warps verify the current tile.
// Two curve warps per sub-partition already drive the FMA-heavy pipe to 94 % of what four do (measured with the
// single-role kernel at 2 CTAs/SM), so the other two warp slots of the 128-register budget go to hash warps; a hash
// warp alone is latency-bound (IPC ~0.17 measured), two per sub-partition keep up with one challenge per
// verification.  Tile = 8 x 32 = 256 tuples per iteration, handed out by a global counter; verdict words never
// straddle warps.
// ------------------------------------------------------------------------------------------------
#ifndef SB_VERIFY_WS
#define SB_VERIFY_WS 1
#endif
#ifndef SB_VARGEN_SPLIT
#define SB_VARGEN_SPLIT 1
#endif
#ifndef SB_VERIFY_SPLIT
#define SB_VERIFY_SPLIT 1  // hash and curve halves as two launches: with the half-size curve path the single kernel's phases evict each other from the instruction cache (21.1 -> 22.9 M verifies/s)
#endif
#ifndef SB_WS_SETMAXNREG
#define SB_WS_SETMAXNREG 0
#endif
#ifndef SB_WS_HREGS
#define SB_WS_HREGS 56
#endif
#ifndef SB_WS_EREGS
#define SB_WS_EREGS 120
#endif
constexpr int WS_HW = 8, WS_EW = 8, WS_THREADS = (WS_HW + WS_EW) * 32, WS_EPER = 1;
constexpr int WS_TILE = WS_EW * 32 * WS_EPER;       // 256 tuples per iteration
constexpr int WS_HPER = WS_TILE / (WS_HW * 32);     // 1 hash per hash lane
constexpr size_t WS_SMEM = (size_t)2 * WS_TILE * 8 * sizeof(uint32_t);
static_assert(WS_HPER * WS_HW * 32 == WS_TILE, "hash and curve warps must cover the same tile");
__global__ void __launch_bounds__(WS_THREADS, 1) k_verify_ws(const KArgs a) {
  extern __shared__ __align__(16) uint32_t ws_smem[];
  uint32_t(*cbuf)[WS_TILE][8] = reinterpret_cast<uint32_t(*)[WS_TILE][8]>(ws_smem);
  __shared__ unsigned tiles[2];
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const bool aff = (a.flags & SB200_POINTS_AFFINE) != 0;
  const unsigned ntiles = (unsigned)((a.n + WS_TILE - 1) / WS_TILE);
  if (tid == 0) tiles[0] = atomicAdd(&a.ws->tile, 1u);
  __syncthreads();
  const bool is_hash = warp < WS_HW;
#if SB_WS_SETMAXNREG
  // register re-partitioning between warpgroups: the hash warps give up what the curve warps need
  if (is_hash) asm volatile("setmaxnreg.dec.sync.aligned.u32 %0;" ::"n"(SB_WS_HREGS));
  else asm volatile("setmaxnreg.inc.sync.aligned.u32 %0;" ::"n"(SB_WS_EREGS));
#endif

  auto hash_tile = [&](unsigned t, uint32_t (*buf)[8]) {
#pragma unroll 1
    for (int s = 0; s < WS_HPER; s++) {
      const int local = (s * WS_HW + warp) * 32 + lane;
      int64_t i = (int64_t)t * WS_TILE + local;
      const bool act = i < a.n;
      if (!act) i = a.n - 1;
      uint32_t c[8];
      fq ru, rv;
      point_to_affine(ldg_point(a.in[2], i, aff), ru, rv);
      challenge3_fd(ru, rv, ldg_fq(a.in[3] + i * 8), c);
      stg8(buf[local], c);
      if (a.out[0] && act) stg8(a.out[0] + i * 8, c);
    }
  };

  int cur = 0;
  if (is_hash && tiles[0] < ntiles) hash_tile(tiles[0], cbuf[0]);
  if (tid == 0) tiles[1] = atomicAdd(&a.ws->tile, 1u);
  __syncthreads();
  while (tiles[cur] < ntiles) {
    const unsigned t = tiles[cur], tn = tiles[cur ^ 1];
    if (is_hash) {
      if (tn < ntiles) hash_tile(tn, cbuf[cur ^ 1]);
    } else {
#pragma unroll 1
      for (int s = 0; s < WS_EPER; s++) {
        const int local0 = (s * WS_EW + (warp - WS_HW)) * 32;
        const int64_t i0 = (int64_t)t * WS_TILE + local0;
        if (i0 >= a.n) break;  // warp-uniform
        int64_t i = i0 + lane;
        const bool act = i < a.n;
        if (!act) i = a.n - 1;
        uint32_t u[8], c[8];
        ldg_scalar(a.in[1] + i * 8, u);
        const uint4* cp = reinterpret_cast<const uint4*>(cbuf[cur][act ? local0 + lane : local0]);
        uint4 c0 = cp[0], c1 = cp[1];
        c[0] = c0.x; c[1] = c0.y; c[2] = c0.z; c[3] = c0.w; c[4] = c1.x; c[5] = c1.y; c[6] = c1.z; c[7] = c1.w;
        bool ok = verify_ec_core(ldg_point(a.in[0], i, aff), u, ldg_point(a.in[2], i, aff), c, a.combG);
        unsigned word = __ballot_sync(0xffffffffu, ok && act);
        if (lane == 0) a.bitmap[i0 >> 5] = word;
      }
    }
    __syncthreads();
    if (tid == 0) tiles[cur] = atomicAdd(&a.ws->tile, 1u);
    cur ^= 1;
    __syncthreads();
  }
}

// Curve half of the single-key verification as a persistent kernel whose window tables live in a thread-major global
// scratch region (18 entries x 128 B per resident thread) instead of local memory: see verify_ec_half.
#ifndef SB_EC_GLOBAL_TABLES
#define SB_EC_GLOBAL_TABLES 0  // measured: DRAM traffic 188 -> 32 GB per 2^22 launch, but the curve kernel takes 149 ms instead of 125 (L1 serves local memory write-back, global stores go through to L2): 20.4 vs 23.0 M verifies/s
#endif
constexpr int EC_P_CTAS = 4;
__global__ void __launch_bounds__(TPB, EC_P_CTAS) k_verify_ec_p(const KArgs a, pniels* scratch) {
  const bool aff = (a.flags & SB200_POINTS_AFFINE) != 0;
  pniels* store = scratch + ((size_t)blockIdx.x * TPB + threadIdx.x) * 18;
  const int64_t ntiles = (a.n + TPB - 1) / TPB;
#pragma unroll 1
  for (int64_t tile = blockIdx.x; tile < ntiles; tile += gridDim.x) {
    int64_t i = tile * TPB + threadIdx.x;
    const bool active = i < a.n;
    if (!active) i = a.n - 1;
    uint32_t u[8], c[8];
    ldg_scalar(a.in[1] + i * 8, u);
    ldg_scalar(a.out[0] + i * 8, c);
    bool ok = verify_ec<true>(ldg_point(a.in[0], i, aff), u, ldg_point(a.in[2], i, aff), c, a.combG, store);
    unsigned word = __ballot_sync(0xffffffffu, ok && active);
    if ((threadIdx.x & 31) == 0 && active) a.bitmap[i >> 5] = word;
  }
}

// Fixed-base signing / key generation with several tuples per thread: the scalar multiples of G (and G') are
// computed first, all their Z coordinates are inverted together (one field inversion per thread instead of one
// per point: the inversion is a third of a single signature's work), then each tuple is finished.
// Thread t of a CTA handles tuples base + j * TPB + t (j < K): loads and stores stay coalesced.
// Tuples per thread: 4 when signing (the hash dominates, more only adds local-memory traffic), SB_KEYGEN_K for key
// generation, where the shared inversion is a quarter of the work at 4 (measured: 311 M keys/s at 4, 369 at 8, 351 at 16).
#ifndef SB_SIGN_K
#define SB_SIGN_K 4
#endif
#ifndef SB_KEYGEN_K
#define SB_KEYGEN_K 8
#endif
__host__ __device__ constexpr int fixed_k(int op) { return (op == OP_KEYGEN || op == OP_KEYGEN_DOUBLE) ? SB_KEYGEN_K : SB_SIGN_K; }
template <int OP>
__global__ void __launch_bounds__(TPB, 4) k_fixed_batch(const KArgs a) {
  constexpr int SIGN_K = fixed_k(OP);
  constexpr bool DOUBLE = (OP == OP_SIGN_DOUBLE || OP == OP_KEYGEN_DOUBLE);
  constexpr bool SIGN = (OP == OP_SIGN || OP == OP_SIGN_DOUBLE || OP == OP_SIGN_BYTES);
  constexpr int NP = DOUBLE ? 2 * SIGN_K : SIGN_K;
  const int64_t base = (int64_t)blockIdx.x * TPB * SIGN_K + threadIdx.x;
  fq X[NP], Y[NP], Z[NP], pre[NP];
  // scalar whose multiple is needed: the nonce when signing, the secret key for keygen
  const uint32_t* kin = SIGN ? a.in[2] : a.in[0];
#pragma unroll 1
  for (int j = 0; j < SIGN_K; j++) {
    int64_t i = base + (int64_t)j * TPB;
    if (i >= a.n) i = a.n - 1;
    uint32_t k[8];
    ldg_scalar(kin + i * 8, k);
    ext p = fixed_base_mul(a.combG, k);
    X[j] = p.X; Y[j] = p.Y; Z[j] = p.Z;
    if (DOUBLE) {
      ext q = fixed_base_mul(a.combGp, k);
      X[SIGN_K + j] = q.X; Y[SIGN_K + j] = q.Y; Z[SIGN_K + j] = q.Z;
    }
  }
  batch_inverse(Z, pre, NP);
#pragma unroll 1
  for (int j = 0; j < SIGN_K; j++) {
    int64_t i = base + (int64_t)j * TPB;
    const bool active = i < a.n;
    if (!active) i = a.n - 1;
    fq Ru = fq_mul(X[j], Z[j]), Rv = fq_mul(Y[j], Z[j]), Rpu, Rpv;
    if (DOUBLE) {
      Rpu = fq_mul(X[SIGN_K + j], Z[SIGN_K + j]);
      Rpv = fq_mul(Y[SIGN_K + j], Z[SIGN_K + j]);
    }
    if (!SIGN) {
      if (active) {
        stg_point(a.out[0], i, Ru, Rv);
        if (DOUBLE) stg_point(a.out[1], i, Rpu, Rpv);
      }
      continue;
    }
    uint32_t sk[8], nonce[8], u[8], c[8];
    ldg_scalar(a.in[0] + i * 8, sk);
    ldg_scalar(a.in[2] + i * 8, nonce);
    fq m = ldg_fq(a.in[1] + i * 8);
    if (OP == OP_SIGN_BYTES) m = fq_to_mont(m);
    if (DOUBLE) chal5(Ru, Rv, Rpu, Rpv, m, c); else chal3(Ru, Rv, m, c);
    sign_finish(nonce, c, sk, u);
    if (!active) continue;
    if (OP == OP_SIGN_BYTES) {
      uint32_t rb[8];
      point_compress(Ru, Rv, rb);
      stg8(a.out[0] + i * 16, u);
      stg8(a.out[0] + i * 16 + 8, rb);
    } else {
      stg8(a.out[0] + i * 8, u);
      stg_point(a.out[1], i, Ru, Rv);
      if (DOUBLE) stg_point(a.out[2], i, Rpu, Rpv);
      if (a.out[3]) stg8(a.out[3] + i * 8, c);
    }
  }
}

__global__ void __launch_bounds__(TPB) k_comb_build(uint32_t* table, int which) {
  int t = blockIdx.x * TPB + threadIdx.x;
  if (t >= COMB_WINDOWS * COMB_ENTRIES) return;  // 524 304 entries at 16-bit windows
  const fq gu = {SB200_G_U_INIT}, gv = {SB200_G_V_INIT}, hu = {SB200_GP_U_INIT}, hv = {SB200_GP_V_INIT};
  comb_build_entry(which ? hu : gu, which ? hv : gv, t / COMB_ENTRIES, t % COMB_ENTRIES, table + (size_t)t * 24);
}

// ------------------------------------------------------------------------------------------------
// host side
// ------------------------------------------------------------------------------------------------
struct DevCtx {
  int dev = 0;
  cudaStream_t stream[2] = {nullptr, nullptr};
  uint32_t* combG = nullptr;
  uint32_t* combGp = nullptr;
  uint8_t* arena[2] = {nullptr, nullptr};
  size_t arena_cap[2] = {0, 0};
  WsState* ws[3] = {nullptr, nullptr, nullptr};  // per pipeline stream, [2] = caller's stream (SB200_DEVICE_PTRS)
  pniels* ec_scratch[3] = {nullptr, nullptr, nullptr};  // k_verify_ec_p window tables, per stream like ws
  uint32_t* cscratch = nullptr;                  // challenges of a SB200_DEVICE_PTRS verify call without c_out
  size_t cscratch_cap = 0;
  int nsm = 0;
};

struct Desc {
  Op op;
  uint32_t flags;
  int aux = 0;
  int nin = 0, nout = 0;
  const uint32_t* in[MAX_IN] = {};
  int in_words[MAX_IN] = {};
  uint32_t* out[MAX_OUT] = {};
  int out_words[MAX_OUT] = {};
  uint32_t* bitmap = nullptr;
};

}  // namespace

struct sb200_ctx {
  std::vector<DevCtx> devs;
  std::mutex mu;
  std::string err;
  cudaStream_t user_stream = nullptr;
  std::atomic<uint64_t> launches{0};
};

namespace {

#define CU(call)                                                                              \
  do {                                                                                        \
    cudaError_t _e = (call);                                                                  \
    if (_e != cudaSuccess) {                                                                  \
      ctx->err = std::string(#call) + ": " + cudaGetErrorString(_e);                          \
      return SB200_ERR_CUDA;                                                                  \
    }                                                                                         \
  } while (0)

int launch(sb200_ctx* ctx, Op op, const KArgs& a, cudaStream_t st) {
  if (a.n <= 0) return SB200_OK;
  unsigned grid = (unsigned)((a.n + TPB - 1) / TPB);
  auto gridk = [&](int op_) { return (unsigned)((a.n + TPB * fixed_k(op_) - 1) / (TPB * fixed_k(op_))); };
#if SB_VERIFY_WS
  if (op == OP_VERIFY && (a.flags & SB200_VERIFY_DUAL_PIPE)) {  // persistent warp-specialised kernel, one CTA per SM
    CU(cudaMemsetAsync(&a.ws->tile, 0, sizeof(unsigned), st));
    unsigned ntiles = (unsigned)((a.n + WS_TILE - 1) / WS_TILE);
    unsigned g = std::min<unsigned>(ntiles, (unsigned)a.nsm);
    k_verify_ws<<<g, WS_THREADS, WS_SMEM, st>>>(a);
    ctx->launches.fetch_add(1, std::memory_order_relaxed);
    CU(cudaGetLastError());
    return SB200_OK;
  }
#endif
#if SB_VERIFY_SPLIT
  if (op == OP_VERIFY) {  // a.out[0] is always set here: the caller's c_out or scratch
    k_run<OP_CHALLENGE><<<grid, TPB, 0, st>>>(a);
#if SB_EC_GLOBAL_TABLES
    k_verify_ec_p<<<std::min<unsigned>(grid, (unsigned)(EC_P_CTAS * a.nsm)), TPB, 0, st>>>(a, a.ec_scratch);
#else
    k_run<OP_VERIFY_EC><<<grid, TPB, 0, st>>>(a);
#endif
    ctx->launches.fetch_add(2, std::memory_order_relaxed);
    CU(cudaGetLastError());
    return SB200_OK;
  }
#if SB_VARGEN_SPLIT
  if (op == OP_VERIFY_VARGEN) {
    k_run<OP_CHALLENGE_VARGEN><<<grid, TPB, 0, st>>>(a);
    k_run<OP_VERIFY_VARGEN_EC><<<grid, TPB, 0, st>>>(a);
    ctx->launches.fetch_add(2, std::memory_order_relaxed);
    CU(cudaGetLastError());
    return SB200_OK;
  }
#endif
  if (op == OP_VERIFY_DOUBLE) {
    k_run<OP_CHALLENGE_DOUBLE><<<grid, TPB, 0, st>>>(a);
    k_run<OP_VERIFY_DOUBLE_EC><<<grid, TPB, 0, st>>>(a);
    ctx->launches.fetch_add(2, std::memory_order_relaxed);
    CU(cudaGetLastError());
    return SB200_OK;
  }
#endif
  switch (op) {
#define CASEK(O) case O: k_fixed_batch<O><<<gridk(O), TPB, 0, st>>>(a); break;
    CASEK(OP_SIGN) CASEK(OP_SIGN_DOUBLE) CASEK(OP_KEYGEN) CASEK(OP_KEYGEN_DOUBLE) CASEK(OP_SIGN_BYTES)
#undef CASEK
#define CASE(O) case O: k_run<O><<<grid, TPB, 0, st>>>(a); break;
    CASE(OP_VERIFY) CASE(OP_VERIFY_DOUBLE) CASE(OP_VERIFY_VARGEN) CASE(OP_SIGN_VARGEN)
    CASE(OP_KEYGEN_VARGEN) CASE(OP_DBG_FQ) CASE(OP_DBG_FR_MUL) CASE(OP_DBG_HADES)
    CASE(OP_DBG_SMUL) CASE(OP_DECOMPRESS) CASE(OP_COMPRESS) CASE(OP_FROM_WIDE) CASE(OP_VERIFY_BYTES)
    CASE(OP_VERIFY_DOUBLE_BYTES) CASE(OP_VERIFY_VARGEN_BYTES) CASE(OP_SIGN_DOUBLE_BYTES) CASE(OP_SIGN_VARGEN_BYTES)
#undef CASE
  }
  ctx->launches.fetch_add(1, std::memory_order_relaxed);
  CU(cudaGetLastError());
  return SB200_OK;
}

int run(sb200_ctx* ctx, int64_t n, const Desc& d) {
  if (!ctx || n < 0) return SB200_ERR_ARG;
  if (d.flags & ~(SB200_POINTS_AFFINE | SB200_DEVICE_PTRS | (d.op == OP_VERIFY ? SB200_VERIFY_DUAL_PIPE : 0u))) return SB200_ERR_ARG;
  for (int k = 0; k < d.nin; k++)
    if (d.in_words[k] && (!d.in[k] || ((uintptr_t)d.in[k] & 15))) return SB200_ERR_ARG;
  for (int k = 0; k < d.nout; k++)
    if (d.out[k] && ((uintptr_t)d.out[k] & 15)) return SB200_ERR_ARG;
  if (n == 0) return SB200_OK;
  std::lock_guard<std::mutex> lock(ctx->mu);

  if (d.flags & SB200_DEVICE_PTRS) {
    if (ctx->devs.size() != 1) return SB200_ERR_ARG;
    DevCtx& dc = ctx->devs[0];
    CU(cudaSetDevice(dc.dev));
    KArgs a{};
    a.n = n; a.flags = d.flags; a.aux = d.aux; a.combG = dc.combG; a.combGp = dc.combGp; a.bitmap = d.bitmap;
    a.ws = dc.ws[2]; a.nsm = dc.nsm; a.ec_scratch = dc.ec_scratch[2];
    for (int k = 0; k < d.nin; k++) a.in[k] = d.in[k];
    for (int k = 0; k < d.nout; k++) a.out[k] = d.out[k];
#if SB_VERIFY_SPLIT
    if (((d.op == OP_VERIFY && !(d.flags & SB200_VERIFY_DUAL_PIPE)) || d.op == OP_VERIFY_DOUBLE || (SB_VARGEN_SPLIT && d.op == OP_VERIFY_VARGEN)) && !a.out[0]) {  // hash and curve kernels hand c over in memory
      size_t need = (size_t)n * 32;
      if (dc.cscratch_cap < need) {
        CU(cudaStreamSynchronize(ctx->user_stream));
        if (dc.cscratch) CU(cudaFree(dc.cscratch));
        dc.cscratch = nullptr; dc.cscratch_cap = 0;
        if (cudaMalloc(&dc.cscratch, need) != cudaSuccess) { ctx->err = "cudaMalloc challenge scratch"; return SB200_ERR_NOMEM; }
        dc.cscratch_cap = need;
      }
      a.out[0] = dc.cscratch;
    }
#endif
    return launch(ctx, d.op, a, ctx->user_stream);
  }

  // bytes per tuple on the device arena (every array padded to 16 B; bitmap = 4 B per 32 tuples)
  size_t per_tuple = 0;
  for (int k = 0; k < d.nin; k++) per_tuple += (size_t)d.in_words[k] * 4;
  for (int k = 0; k < d.nout; k++) per_tuple += (d.out[k] && d.out_words[k] > 0) ? (size_t)d.out_words[k] * 4 : 0;
  const bool c_scratch = SB_VERIFY_SPLIT && (d.op == OP_VERIFY || d.op == OP_VERIFY_DOUBLE || (SB_VARGEN_SPLIT && d.op == OP_VERIFY_VARGEN)) && !d.out[0];  // device-only challenge rows between the two kernels
  if (c_scratch) per_tuple += 32;

  const int ndev = (int)ctx->devs.size();
  int64_t per_dev = ((n + ndev - 1) / ndev + 31) & ~(int64_t)31;
  for (int di = 0; di < ndev; di++) {
    DevCtx& dc = ctx->devs[di];
    int64_t lo = std::min<int64_t>(n, di * per_dev), hi = std::min<int64_t>(n, lo + per_dev);
    if (lo >= hi) continue;
    CU(cudaSetDevice(dc.dev));
    int slot = 0;
    for (int64_t c0 = lo; c0 < hi; c0 += CHUNK, slot ^= 1) {
      int64_t cn = std::min<int64_t>(CHUNK, hi - c0);
      size_t need = (size_t)cn * per_tuple + (size_t)((cn + 31) / 32) * 4 * (1 + MAX_OUT) + 64 * (MAX_IN + MAX_OUT + 1);
      if (dc.arena_cap[slot] < need) {
        CU(cudaStreamSynchronize(dc.stream[slot]));
        if (dc.arena[slot]) CU(cudaFree(dc.arena[slot]));
        dc.arena[slot] = nullptr; dc.arena_cap[slot] = 0;
        size_t cap = std::max(need, (size_t)std::min<int64_t>(CHUNK, per_dev) * per_tuple + (1 << 20));
        if (cudaMalloc(&dc.arena[slot], cap) != cudaSuccess) { ctx->err = "cudaMalloc arena"; return SB200_ERR_NOMEM; }
        dc.arena_cap[slot] = cap;
      }
      cudaStream_t st = dc.stream[slot];
      uint8_t* p = dc.arena[slot];
      auto carve = [&](size_t bytes) { uint8_t* r = p; p += (bytes + 63) & ~(size_t)63; return r; };
      KArgs a{};
      a.n = cn; a.flags = d.flags; a.aux = d.aux; a.combG = dc.combG; a.combGp = dc.combGp;
      a.ws = dc.ws[slot]; a.nsm = dc.nsm; a.ec_scratch = dc.ec_scratch[slot];
      for (int k = 0; k < d.nin; k++) {
        if (!d.in_words[k]) continue;
        size_t bytes = (size_t)cn * d.in_words[k] * 4;
        uint32_t* dp = (uint32_t*)carve(bytes);
        CU(cudaMemcpyAsync(dp, d.in[k] + (size_t)c0 * d.in_words[k], bytes, cudaMemcpyHostToDevice, st));
        a.in[k] = dp;
      }
      uint32_t* dout[MAX_OUT] = {};
      auto out_bytes = [&](int k) {  // out_words == -1: bitmap-shaped output (one word per 32 tuples)
        return d.out_words[k] < 0 ? (size_t)((cn + 31) / 32) * 4 : (size_t)cn * d.out_words[k] * 4;
      };
      for (int k = 0; k < d.nout; k++)
        if (d.out[k]) a.out[k] = dout[k] = (uint32_t*)carve(out_bytes(k));
      if (c_scratch) a.out[0] = (uint32_t*)carve((size_t)cn * 32);
      uint32_t* dbm = nullptr;
      if (d.bitmap) a.bitmap = dbm = (uint32_t*)carve((size_t)((cn + 31) / 32) * 4);
      int rc = launch(ctx, d.op, a, st);
      if (rc) return rc;
      for (int k = 0; k < d.nout; k++)
        if (d.out[k])
          CU(cudaMemcpyAsync(d.out_words[k] < 0 ? d.out[k] + c0 / 32 : d.out[k] + (size_t)c0 * d.out_words[k], dout[k],
                             out_bytes(k), cudaMemcpyDeviceToHost, st));
      if (d.bitmap)
        CU(cudaMemcpyAsync(d.bitmap + c0 / 32, dbm, (size_t)((cn + 31) / 32) * 4, cudaMemcpyDeviceToHost, st));
    }
  }
  for (auto& dc : ctx->devs) {
    CU(cudaSetDevice(dc.dev));
    CU(cudaStreamSynchronize(dc.stream[0]));
    CU(cudaStreamSynchronize(dc.stream[1]));
  }
  return SB200_OK;
}

int pt_words(uint32_t flags) { return (flags & SB200_POINTS_AFFINE) ? 16 : 24; }

}  // namespace

extern "C" {

int sb200_init(const int* devices, int n_devices, sb200_ctx** out) {
  if (!out || n_devices < 0 || (n_devices > 0 && !devices)) return SB200_ERR_ARG;
  *out = nullptr;
  int count = 0;
  if (cudaGetDeviceCount(&count) != cudaSuccess || count == 0) return SB200_ERR_NODEV;
  int dflt = 0;
  if (n_devices == 0) { devices = &dflt; n_devices = 1; }
  sb200_ctx* ctx = new sb200_ctx();
  auto fail = [&](int code) { sb200_destroy(ctx); return code; };
  for (int i = 0; i < n_devices; i++) {
    if (devices[i] < 0 || devices[i] >= count) return fail(SB200_ERR_ARG);
    cudaDeviceProp prop;
    if (cudaGetDeviceProperties(&prop, devices[i]) != cudaSuccess || prop.major < 10) return fail(SB200_ERR_NODEV);
    DevCtx dc;
    dc.dev = devices[i];
    if (cudaSetDevice(dc.dev) != cudaSuccess) return fail(SB200_ERR_CUDA);
    for (int s = 0; s < 2; s++)
      if (cudaStreamCreateWithFlags(&dc.stream[s], cudaStreamNonBlocking) != cudaSuccess) return fail(SB200_ERR_CUDA);
    size_t tb = (size_t)COMB_WINDOWS * COMB_ENTRIES * 24 * 4;
    if (cudaMalloc(&dc.combG, tb) != cudaSuccess || cudaMalloc(&dc.combGp, tb) != cudaSuccess) return fail(SB200_ERR_NOMEM);
    dc.nsm = prop.multiProcessorCount;
    if (cudaFuncSetAttribute(k_verify_ws, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)WS_SMEM) != cudaSuccess) return fail(SB200_ERR_CUDA);
    for (int s = 0; s < 3; s++)
      if (cudaMalloc(&dc.ws[s], sizeof(WsState)) != cudaSuccess || cudaMemset(dc.ws[s], 0, sizeof(WsState)) != cudaSuccess)
        return fail(SB200_ERR_NOMEM);
#if SB_EC_GLOBAL_TABLES
    for (int s = 0; s < 3; s++)  // 4 x 148 x 128 threads x 2 304 B = 175 MB per stream
      if (cudaMalloc(&dc.ec_scratch[s], (size_t)EC_P_CTAS * dc.nsm * TPB * 18 * sizeof(pniels)) != cudaSuccess) return fail(SB200_ERR_NOMEM);
#endif
    ctx->devs.push_back(dc);
    int grid = (COMB_WINDOWS * COMB_ENTRIES + TPB - 1) / TPB;
    k_comb_build<<<grid, TPB, 0, dc.stream[0]>>>(dc.combG, 0);
    k_comb_build<<<grid, TPB, 0, dc.stream[0]>>>(dc.combGp, 1);
    ctx->launches += 2;
    if (cudaStreamSynchronize(dc.stream[0]) != cudaSuccess) return fail(SB200_ERR_CUDA);
  }
  *out = ctx;
  return SB200_OK;
}

void sb200_destroy(sb200_ctx* ctx) {
  if (!ctx) return;
  for (auto& dc : ctx->devs) {
    cudaSetDevice(dc.dev);
    for (int s = 0; s < 2; s++) {
      if (dc.stream[s]) { cudaStreamSynchronize(dc.stream[s]); cudaStreamDestroy(dc.stream[s]); }
      if (dc.arena[s]) cudaFree(dc.arena[s]);
    }
    if (dc.combG) cudaFree(dc.combG);
    if (dc.combGp) cudaFree(dc.combGp);
    for (int s = 0; s < 3; s++)
      if (dc.ws[s]) cudaFree(dc.ws[s]);
    if (dc.cscratch) cudaFree(dc.cscratch);
    for (int s = 0; s < 3; s++)
      if (dc.ec_scratch[s]) cudaFree(dc.ec_scratch[s]);
  }
  delete ctx;
}

const char* sb200_strerror(int code) {
  switch (code) {
    case SB200_OK: return "ok";
    case SB200_ERR_ARG: return "invalid argument";
    case SB200_ERR_CUDA: return "CUDA error";
    case SB200_ERR_NODEV: return "no usable sm_100 CUDA device";
    case SB200_ERR_NOMEM: return "out of device memory";
    default: return "unknown error";
  }
}
const char* sb200_last_error(const sb200_ctx* ctx) { return ctx ? ctx->err.c_str() : ""; }
int sb200_device_count(const sb200_ctx* ctx) { return ctx ? (int)ctx->devs.size() : 0; }
int sb200_set_stream(sb200_ctx* ctx, void* s) {
  if (!ctx) return SB200_ERR_ARG;
  ctx->user_stream = (cudaStream_t)s;
  return SB200_OK;
}
uint64_t sb200_launch_count(const sb200_ctx* ctx) { return ctx ? ctx->launches.load() : 0; }
int sb200_host_alloc(size_t bytes, void** out) {
  if (!out) return SB200_ERR_ARG;
  return cudaHostAlloc(out, bytes, cudaHostAllocPortable) == cudaSuccess ? SB200_OK : SB200_ERR_NOMEM;
}
void sb200_host_free(void* p) { if (p) cudaFreeHost(p); }

#define IN(k, ptr, words) d.in[k] = (ptr); d.in_words[k] = (words)
#define OUT(k, ptr, words) d.out[k] = (ptr); d.out_words[k] = (words)

int sb200_verify(sb200_ctx* ctx, int64_t n, uint32_t flags, const uint32_t* pk, const uint32_t* sig_u, const uint32_t* sig_R,
                 const uint32_t* msg, uint32_t* verdicts, uint32_t* c_out) {
  if (!verdicts) return SB200_ERR_ARG;
  Desc d; d.op = OP_VERIFY; d.flags = flags; d.nin = 4; d.nout = 1; d.bitmap = verdicts;
  IN(0, pk, pt_words(flags)); IN(1, sig_u, 8); IN(2, sig_R, pt_words(flags)); IN(3, msg, 8); OUT(0, c_out, 8);
  return run(ctx, n, d);
}
int sb200_verify_double(sb200_ctx* ctx, int64_t n, uint32_t flags, const uint32_t* pk, const uint32_t* pkp, const uint32_t* sig_u,
                        const uint32_t* sig_R, const uint32_t* sig_Rp, const uint32_t* msg, uint32_t* verdicts, uint32_t* c_out) {
  if (!verdicts) return SB200_ERR_ARG;
  Desc d; d.op = OP_VERIFY_DOUBLE; d.flags = flags; d.nin = 6; d.nout = 1; d.bitmap = verdicts;
  int pw = pt_words(flags);
  IN(0, pk, pw); IN(1, pkp, pw); IN(2, sig_u, 8); IN(3, sig_R, pw); IN(4, sig_Rp, pw); IN(5, msg, 8); OUT(0, c_out, 8);
  return run(ctx, n, d);
}
int sb200_verify_vargen(sb200_ctx* ctx, int64_t n, uint32_t flags, const uint32_t* pk, const uint32_t* gen, const uint32_t* sig_u,
                        const uint32_t* sig_R, const uint32_t* msg, uint32_t* verdicts, uint32_t* c_out) {
  if (!verdicts) return SB200_ERR_ARG;
  Desc d; d.op = OP_VERIFY_VARGEN; d.flags = flags; d.nin = 5; d.nout = 1; d.bitmap = verdicts;
  int pw = pt_words(flags);
  IN(0, pk, pw); IN(1, gen, pw); IN(2, sig_u, 8); IN(3, sig_R, pw); IN(4, msg, 8); OUT(0, c_out, 8);
  return run(ctx, n, d);
}
int sb200_sign(sb200_ctx* ctx, int64_t n, uint32_t flags, const uint32_t* sk, const uint32_t* msg, const uint32_t* nonce,
               uint32_t* u_out, uint32_t* R_out, uint32_t* c_out) {
  if (!u_out || !R_out) return SB200_ERR_ARG;
  Desc d; d.op = OP_SIGN; d.flags = flags; d.nin = 3; d.nout = 4;
  IN(0, sk, 8); IN(1, msg, 8); IN(2, nonce, 8); OUT(0, u_out, 8); OUT(1, R_out, 16); OUT(3, c_out, 8);
  return run(ctx, n, d);
}
int sb200_sign_double(sb200_ctx* ctx, int64_t n, uint32_t flags, const uint32_t* sk, const uint32_t* msg, const uint32_t* nonce,
                      uint32_t* u_out, uint32_t* R_out, uint32_t* Rp_out, uint32_t* c_out) {
  if (!u_out || !R_out || !Rp_out) return SB200_ERR_ARG;
  Desc d; d.op = OP_SIGN_DOUBLE; d.flags = flags; d.nin = 3; d.nout = 4;
  IN(0, sk, 8); IN(1, msg, 8); IN(2, nonce, 8); OUT(0, u_out, 8); OUT(1, R_out, 16); OUT(2, Rp_out, 16); OUT(3, c_out, 8);
  return run(ctx, n, d);
}
int sb200_sign_vargen(sb200_ctx* ctx, int64_t n, uint32_t flags, const uint32_t* sk, const uint32_t* gen, const uint32_t* msg,
                      const uint32_t* nonce, uint32_t* u_out, uint32_t* R_out, uint32_t* c_out) {
  if (!u_out || !R_out) return SB200_ERR_ARG;
  Desc d; d.op = OP_SIGN_VARGEN; d.flags = flags; d.nin = 4; d.nout = 4;
  IN(0, sk, 8); IN(1, gen, pt_words(flags)); IN(2, msg, 8); IN(3, nonce, 8); OUT(0, u_out, 8); OUT(1, R_out, 16); OUT(3, c_out, 8);
  return run(ctx, n, d);
}
int sb200_keygen(sb200_ctx* ctx, int64_t n, uint32_t flags, const uint32_t* sk, uint32_t* pk_out) {
  if (!pk_out) return SB200_ERR_ARG;
  Desc d; d.op = OP_KEYGEN; d.flags = flags; d.nin = 1; d.nout = 1;
  IN(0, sk, 8); OUT(0, pk_out, 16);
  return run(ctx, n, d);
}
int sb200_keygen_double(sb200_ctx* ctx, int64_t n, uint32_t flags, const uint32_t* sk, uint32_t* pk_out, uint32_t* pkp_out) {
  if (!pk_out || !pkp_out) return SB200_ERR_ARG;
  Desc d; d.op = OP_KEYGEN_DOUBLE; d.flags = flags; d.nin = 1; d.nout = 2;
  IN(0, sk, 8); OUT(0, pk_out, 16); OUT(1, pkp_out, 16);
  return run(ctx, n, d);
}
int sb200_keygen_vargen(sb200_ctx* ctx, int64_t n, uint32_t flags, const uint32_t* sk, const uint32_t* gen, uint32_t* pk_out) {
  if (!pk_out) return SB200_ERR_ARG;
  Desc d; d.op = OP_KEYGEN_VARGEN; d.flags = flags; d.nin = 2; d.nout = 1;
  IN(0, sk, 8); IN(1, gen, pt_words(flags)); OUT(0, pk_out, 16);
  return run(ctx, n, d);
}
int sb200_points_decompress(sb200_ctx* ctx, int64_t n, uint32_t flags, const uint8_t* bytes, uint32_t* points_out, uint32_t* ok_bitmap) {
  if (!points_out || !ok_bitmap || (flags & SB200_POINTS_AFFINE)) return SB200_ERR_ARG;
  Desc d; d.op = OP_DECOMPRESS; d.flags = flags; d.nin = 1; d.nout = 1; d.bitmap = ok_bitmap;
  IN(0, (const uint32_t*)bytes, 8); OUT(0, points_out, 16);
  return run(ctx, n, d);
}
int sb200_points_compress(sb200_ctx* ctx, int64_t n, uint32_t flags, const uint32_t* points, uint8_t* bytes_out) {
  if (!bytes_out) return SB200_ERR_ARG;
  Desc d; d.op = OP_COMPRESS; d.flags = flags; d.nin = 1; d.nout = 1;
  IN(0, points, pt_words(flags)); OUT(0, (uint32_t*)bytes_out, 8);
  return run(ctx, n, d);
}
int sb200_scalars_from_wide(sb200_ctx* ctx, int64_t n, uint32_t flags, int field, const uint8_t* wide, uint32_t* out) {
  if (!out || field < 0 || field > 1 || (flags & SB200_POINTS_AFFINE)) return SB200_ERR_ARG;
  Desc d; d.op = OP_FROM_WIDE; d.flags = flags; d.aux = field; d.nin = 1; d.nout = 1;
  IN(0, (const uint32_t*)wide, 16); OUT(0, out, 8);
  return run(ctx, n, d);
}
int sb200_fq_to_mont(sb200_ctx* ctx, int64_t n, uint32_t flags, const uint32_t* in, uint32_t* out) {
  if (!out || (flags & SB200_POINTS_AFFINE)) return SB200_ERR_ARG;
  Desc d; d.op = OP_DBG_FQ; d.flags = flags; d.aux = 5; d.nin = 2; d.nout = 1;
  IN(0, in, 8); IN(1, nullptr, 0); OUT(0, out, 8);
  return run(ctx, n, d);
}
int sb200_fq_from_mont(sb200_ctx* ctx, int64_t n, uint32_t flags, const uint32_t* in, uint32_t* out) {
  if (!out || (flags & SB200_POINTS_AFFINE)) return SB200_ERR_ARG;
  Desc d; d.op = OP_DBG_FQ; d.flags = flags; d.aux = 6; d.nin = 2; d.nout = 1;
  IN(0, in, 8); IN(1, nullptr, 0); OUT(0, out, 8);
  return run(ctx, n, d);
}
int sb200_verify_bytes(sb200_ctx* ctx, int64_t n, uint32_t flags, const uint8_t* pk, const uint8_t* sig, const uint8_t* msg,
                       uint32_t* verdicts, uint32_t* invalid) {
  if (!verdicts || (flags & SB200_POINTS_AFFINE)) return SB200_ERR_ARG;
  Desc d; d.op = OP_VERIFY_BYTES; d.flags = flags | SB200_POINTS_AFFINE; d.nin = 3; d.nout = 1; d.bitmap = verdicts;
  IN(0, (const uint32_t*)pk, 8); IN(1, (const uint32_t*)sig, 16); IN(2, (const uint32_t*)msg, 8);
  d.out[0] = invalid; d.out_words[0] = -1;  // bitmap-shaped output: one word per 32 tuples
  return run(ctx, n, d);
}
int sb200_sign_bytes(sb200_ctx* ctx, int64_t n, uint32_t flags, const uint8_t* sk, const uint8_t* msg, const uint8_t* nonce,
                     uint8_t* sig_out) {
  if (!sig_out || (flags & SB200_POINTS_AFFINE)) return SB200_ERR_ARG;
  Desc d; d.op = OP_SIGN_BYTES; d.flags = flags; d.nin = 3; d.nout = 1;
  IN(0, (const uint32_t*)sk, 8); IN(1, (const uint32_t*)msg, 8); IN(2, (const uint32_t*)nonce, 8); OUT(0, (uint32_t*)sig_out, 16);
  return run(ctx, n, d);
}
int sb200_verify_double_bytes(sb200_ctx* ctx, int64_t n, uint32_t flags, const uint8_t* pk, const uint8_t* sig, const uint8_t* msg,
                              uint32_t* verdicts, uint32_t* invalid) {
  if (!verdicts || (flags & SB200_POINTS_AFFINE)) return SB200_ERR_ARG;
  Desc d; d.op = OP_VERIFY_DOUBLE_BYTES; d.flags = flags | SB200_POINTS_AFFINE; d.nin = 3; d.nout = 1; d.bitmap = verdicts;
  IN(0, (const uint32_t*)pk, 16); IN(1, (const uint32_t*)sig, 24); IN(2, (const uint32_t*)msg, 8);
  d.out[0] = invalid; d.out_words[0] = -1;
  return run(ctx, n, d);
}
int sb200_verify_vargen_bytes(sb200_ctx* ctx, int64_t n, uint32_t flags, const uint8_t* pk, const uint8_t* sig, const uint8_t* msg,
                              uint32_t* verdicts, uint32_t* invalid) {
  if (!verdicts || (flags & SB200_POINTS_AFFINE)) return SB200_ERR_ARG;
  Desc d; d.op = OP_VERIFY_VARGEN_BYTES; d.flags = flags | SB200_POINTS_AFFINE; d.nin = 3; d.nout = 1; d.bitmap = verdicts;
  IN(0, (const uint32_t*)pk, 16); IN(1, (const uint32_t*)sig, 16); IN(2, (const uint32_t*)msg, 8);
  d.out[0] = invalid; d.out_words[0] = -1;
  return run(ctx, n, d);
}
int sb200_sign_double_bytes(sb200_ctx* ctx, int64_t n, uint32_t flags, const uint8_t* sk, const uint8_t* msg, const uint8_t* nonce,
                            uint8_t* sig_out) {
  if (!sig_out || (flags & SB200_POINTS_AFFINE)) return SB200_ERR_ARG;
  Desc d; d.op = OP_SIGN_DOUBLE_BYTES; d.flags = flags; d.nin = 3; d.nout = 1;
  IN(0, (const uint32_t*)sk, 8); IN(1, (const uint32_t*)msg, 8); IN(2, (const uint32_t*)nonce, 8); OUT(0, (uint32_t*)sig_out, 24);
  return run(ctx, n, d);
}
int sb200_sign_vargen_bytes(sb200_ctx* ctx, int64_t n, uint32_t flags, const uint8_t* sk, const uint8_t* msg, const uint8_t* nonce,
                            uint8_t* sig_out, uint32_t* ok_bitmap) {
  if (!sig_out || !ok_bitmap || (flags & SB200_POINTS_AFFINE)) return SB200_ERR_ARG;
  Desc d; d.op = OP_SIGN_VARGEN_BYTES; d.flags = flags | SB200_POINTS_AFFINE; d.nin = 3; d.nout = 1; d.bitmap = ok_bitmap;
  IN(0, (const uint32_t*)sk, 16); IN(1, (const uint32_t*)msg, 8); IN(2, (const uint32_t*)nonce, 8); OUT(0, (uint32_t*)sig_out, 16);
  return run(ctx, n, d);
}
int sb200_dbg_fq(sb200_ctx* ctx, int64_t n, int op, const uint32_t* a, const uint32_t* b, uint32_t* out) {
  if (!out || op < 0 || op > 6) return SB200_ERR_ARG;
  Desc d; d.op = OP_DBG_FQ; d.flags = 0; d.aux = op; d.nin = 2; d.nout = 1;
  IN(0, a, 8); IN(1, b, b ? 8 : 0); OUT(0, out, 8);
  return run(ctx, n, d);
}
int sb200_dbg_fr_mul(sb200_ctx* ctx, int64_t n, const uint32_t* a, const uint32_t* b, uint32_t* out) {
  if (!out) return SB200_ERR_ARG;
  Desc d; d.op = OP_DBG_FR_MUL; d.flags = 0; d.nin = 2; d.nout = 1;
  IN(0, a, 8); IN(1, b, 8); OUT(0, out, 8);
  return run(ctx, n, d);
}
int sb200_dbg_hades(sb200_ctx* ctx, int64_t n, int dense, uint32_t* states) {
  if (!states) return SB200_ERR_ARG;
  Desc d; d.op = OP_DBG_HADES; d.flags = 0; d.aux = dense; d.nin = 1; d.nout = 1;
  IN(0, states, 40); OUT(0, states, 40);
  return run(ctx, n, d);
}
int sb200_dbg_scalar_mul(sb200_ctx* ctx, int64_t n, uint32_t flags, int base, const uint32_t* points, const uint32_t* k, uint32_t* out) {
  if (!out || base < 0 || base > 2 || (base == 2 && !points)) return SB200_ERR_ARG;
  Desc d; d.op = OP_DBG_SMUL; d.flags = flags; d.aux = base; d.nin = 2; d.nout = 1;
  IN(0, points, base == 2 ? pt_words(flags) : 0); IN(1, k, 8); OUT(0, out, 16);
  return run(ctx, n, d);
}

}  // extern "C"
