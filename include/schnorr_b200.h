/* schnorr_b200 -- C ABI of the B200 batch engine for dusk-schnorr's sign / verify hot path.
 *
 * The reference (dusk-schnorr 0.18.0, pure Rust) has no FFI; its boundary for this path is the
 * crate's public API.  Each entry point below is the batch form of one reference method and is what
 * a `#[link(name = "schnorr_b200")] extern "C"` block in the crate would bind:
 *
 *   sb200_verify          PublicKey::verify            /root/reference/src/keys/public.rs:121-130
 *   sb200_verify_double   PublicKeyDouble::verify      /root/reference/src/keys/public.rs:222-244
 *   sb200_verify_vargen   PublicKeyVarGen::verify      /root/reference/src/keys/public.rs:401-415
 *   sb200_sign            SecretKey::sign              /root/reference/src/keys/secret.rs:150-168
 *   sb200_sign_double     SecretKey::sign_double       /root/reference/src/keys/secret.rs:217-240
 *   sb200_sign_vargen     SecretKeyVarGen::sign        /root/reference/src/keys/secret.rs:433-451
 *   sb200_keygen          PublicKey::from(&SecretKey)          /root/reference/src/keys/public.rs:61-67
 *   sb200_keygen_double   PublicKeyDouble::from(&SecretKey)    /root/reference/src/keys/public.rs:265-272
 *   sb200_keygen_vargen   PublicKeyVarGen::from(&SecretKeyVarGen) /root/reference/src/keys/public.rs:337-344
 *
 * Data layout (all arrays are caller-owned, tuple-major, tightly packed, 16-byte aligned):
 *   field element (BlsScalar: message, point coordinates) = 8 x u32 little-endian limbs in
 *       MONTGOMERY form (x * 2^256 mod q) -- bit-identical to `BlsScalar.0: [u64; 4]` on a
 *       little-endian host, so Rust hands over its internal limbs without conversion.  Must be < q.
 *   scalar (JubJubScalar: sk, nonce, u, c) = 8 x u32 little-endian limbs of the CANONICAL integer
 *       (= `JubJubScalar::to_bytes()`), must be < r.  A `u` >= r makes that tuple's verdict 0.
 *   point  = SB200_POINTS_AFFINE:      (u, v)      2 field elements (64 B);
 *            SB200_POINTS_PROJECTIVE:  (U, V, Z)   3 field elements (96 B), Z != 0 -- the
 *            (u, v, z) of a JubJubExtended; t1/t2 are not needed.  Points must be on the curve
 *            (the reference can only produce on-curve points outside `from_raw_unchecked`).
 *       One flag applies to every point array of the call.  Points written by the library are
 *       always affine (u, v), Montgomery form.
 *   verdicts = bitmap, bit (i & 31) of word (i >> 5) is tuple i's `verify` result; ceil(n/32) words.
 *
 * Nonces are an INPUT (`nonce[i]` is the one `JubJubScalar::random(rng)` draw of signature i,
 * /root/reference/src/keys/secret.rs:155): the RNG stays on the host; the library never generates
 * randomness.
 *
 * Threading: one batch in flight per context (calls on one context are serialised by an internal
 * mutex); distinct contexts are independent.  Calls block until the outputs are in host memory,
 * except with SB200_DEVICE_PTRS (see below).  No call aborts; errors are return codes.
 * There is no CPU fallback: without a usable CUDA device sb200_init fails with SB200_ERR_NODEV.
 */
#ifndef SCHNORR_B200_H
#define SCHNORR_B200_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

typedef struct sb200_ctx sb200_ctx;

enum {
  SB200_OK = 0,
  SB200_ERR_ARG = -1,   /* null pointer, n < 0, unknown flag, misaligned buffer */
  SB200_ERR_CUDA = -2,  /* a CUDA call failed; sb200_last_error() has the text */
  SB200_ERR_NODEV = -3, /* no usable sm_100 device */
  SB200_ERR_NOMEM = -4
};

#define SB200_POINTS_PROJECTIVE 0u
#define SB200_POINTS_AFFINE 1u
/* All buffers are device pointers on the context's (single) device; the call enqueues its kernel on
 * the stream set with sb200_set_stream and returns without synchronising. */
#define SB200_DEVICE_PTRS 2u
/* sb200_verify only: run the warp-specialised kernel (hash warps on the FP64 pipe beside curve warps on the
 * FMA-heavy pipe).  Same verdicts and challenges; measured at parity with the default kernel. */
#define SB200_VERIFY_DUAL_PIPE 4u

/* devices: CUDA ordinals to shard over (tuples are split into contiguous blocks, multiples of 32);
 * n_devices = 0 means "device 0".  Builds the comb tables of G and G' on every device. */
int sb200_init(const int* devices, int n_devices, sb200_ctx** out);
void sb200_destroy(sb200_ctx* ctx);
const char* sb200_strerror(int code);
const char* sb200_last_error(const sb200_ctx* ctx);
int sb200_device_count(const sb200_ctx* ctx);
/* stream (a cudaStream_t) used by SB200_DEVICE_PTRS calls; NULL = the legacy default stream */
int sb200_set_stream(sb200_ctx* ctx, void* cuda_stream);
/* number of kernels this context has launched since creation (bench.py's gpu_launches) */
uint64_t sb200_launch_count(const sb200_ctx* ctx);
/* pinned host memory for full-speed host<->device copies (pageable memory also works, slower) */
int sb200_host_alloc(size_t bytes, void** out);
void sb200_host_free(void* p);

/* c_out (n x 8 u32, canonical challenge scalars) may be NULL in every call below. */
int sb200_verify(sb200_ctx* ctx, int64_t n, uint32_t flags, const uint32_t* pk, const uint32_t* sig_u,
                 const uint32_t* sig_R, const uint32_t* msg, uint32_t* verdicts, uint32_t* c_out);
int sb200_verify_double(sb200_ctx* ctx, int64_t n, uint32_t flags, const uint32_t* pk, const uint32_t* pk_prime,
                        const uint32_t* sig_u, const uint32_t* sig_R, const uint32_t* sig_R_prime, const uint32_t* msg,
                        uint32_t* verdicts, uint32_t* c_out);
int sb200_verify_vargen(sb200_ctx* ctx, int64_t n, uint32_t flags, const uint32_t* pk, const uint32_t* generator,
                        const uint32_t* sig_u, const uint32_t* sig_R, const uint32_t* msg, uint32_t* verdicts,
                        uint32_t* c_out);

/* u_out: n x 8 u32 canonical; R_out (and R_prime_out): n x 16 u32 affine (u, v) Montgomery. */
int sb200_sign(sb200_ctx* ctx, int64_t n, uint32_t flags, const uint32_t* sk, const uint32_t* msg, const uint32_t* nonce,
               uint32_t* u_out, uint32_t* R_out, uint32_t* c_out);
int sb200_sign_double(sb200_ctx* ctx, int64_t n, uint32_t flags, const uint32_t* sk, const uint32_t* msg,
                      const uint32_t* nonce, uint32_t* u_out, uint32_t* R_out, uint32_t* R_prime_out, uint32_t* c_out);
int sb200_sign_vargen(sb200_ctx* ctx, int64_t n, uint32_t flags, const uint32_t* sk, const uint32_t* generator,
                      const uint32_t* msg, const uint32_t* nonce, uint32_t* u_out, uint32_t* R_out, uint32_t* c_out);

/* pk_out: n x 16 u32 affine (u, v) Montgomery */
int sb200_keygen(sb200_ctx* ctx, int64_t n, uint32_t flags, const uint32_t* sk, uint32_t* pk_out);
int sb200_keygen_double(sb200_ctx* ctx, int64_t n, uint32_t flags, const uint32_t* sk, uint32_t* pk_out,
                        uint32_t* pk_prime_out);
int sb200_keygen_vargen(sb200_ctx* ctx, int64_t n, uint32_t flags, const uint32_t* sk, const uint32_t* generator,
                        uint32_t* pk_out);

/* ---- wire formats on the device (the reference's serialised forms) ---------------------
 * bytes32 point  = JubJubAffine::to_bytes (v little-endian, bit 255 = low bit of u);
 * sig64          = Signature::to_bytes = u (32 B canonical) || compressed R   /root/reference/src/signatures.rs:106-123
 * msg32 / sk32   = BlsScalar::to_bytes / SecretKey::to_bytes (32 B canonical little-endian)
 * `flags` here may only carry SB200_DEVICE_PTRS. */
/* JubJubAffine::from_bytes for n points; ok bit i = 0 where the reference returns None (v >= q or not on the
 * curve).  No subgroup check, as in the reference (/root/reference/src/keys/public.rs:94-100). */
int sb200_points_decompress(sb200_ctx* ctx, int64_t n, uint32_t flags, const uint8_t* bytes32, uint32_t* points_out,
                            uint32_t* ok_bitmap);
/* JubJubAffine::from(point).to_bytes(); points per SB200_POINTS_AFFINE / PROJECTIVE in flags */
int sb200_points_compress(sb200_ctx* ctx, int64_t n, uint32_t flags, const uint32_t* points, uint8_t* bytes32_out);
/* Field::random's from_bytes_wide on n 64-byte draws: field 0 -> JubJubScalar (canonical limbs),
 * field 1 -> BlsScalar (Montgomery limbs).  /root/reference/src/keys/secret.rs:83,155 */
int sb200_scalars_from_wide(sb200_ctx* ctx, int64_t n, uint32_t flags, int field, const uint8_t* wide64, uint32_t* out);
/* canonical <-> Montgomery limbs of n field elements (BlsScalar::from_bytes / to_bytes minus the range check) */
int sb200_fq_to_mont(sb200_ctx* ctx, int64_t n, uint32_t flags, const uint32_t* in, uint32_t* out);
int sb200_fq_from_mont(sb200_ctx* ctx, int64_t n, uint32_t flags, const uint32_t* in, uint32_t* out);
/* PublicKey::from_bytes(pk)?.verify(&Signature::from_bytes(sig)?, BlsScalar::from_bytes(msg)?) for n tuples.
 * invalid bit i = 1 where any of the three from_bytes would return Err(InvalidData); its verdict bit is 0.
 * `invalid` may be NULL. */
int sb200_verify_bytes(sb200_ctx* ctx, int64_t n, uint32_t flags, const uint8_t* pk32, const uint8_t* sig64,
                       const uint8_t* msg32, uint32_t* verdicts, uint32_t* invalid);
/* SecretKey::from_bytes(sk).sign(nonce, BlsScalar::from_bytes(msg)).to_bytes() for n tuples */
int sb200_sign_bytes(sb200_ctx* ctx, int64_t n, uint32_t flags, const uint8_t* sk32, const uint8_t* msg32,
                     const uint8_t* nonce32, uint8_t* sig64_out);

/* The same for the other two schemes:
 * pk64  = PublicKeyDouble::to_bytes = pk || pk'           /root/reference/src/keys/public.rs:282-299
 *       | PublicKeyVarGen::to_bytes = pk || generator     /root/reference/src/keys/public.rs:347-372
 * sig96 = SignatureDouble::to_bytes = u || R || R'        /root/reference/src/signatures.rs:245-270
 * sig64 = SignatureVarGen::to_bytes = u || R              /root/reference/src/signatures.rs:387-404
 * sk64  = SecretKeyVarGen::to_bytes = sk || generator     /root/reference/src/keys/secret.rs:313-336 */
int sb200_verify_double_bytes(sb200_ctx* ctx, int64_t n, uint32_t flags, const uint8_t* pk64, const uint8_t* sig96,
                              const uint8_t* msg32, uint32_t* verdicts, uint32_t* invalid);
int sb200_verify_vargen_bytes(sb200_ctx* ctx, int64_t n, uint32_t flags, const uint8_t* pk64, const uint8_t* sig64,
                              const uint8_t* msg32, uint32_t* verdicts, uint32_t* invalid);
/* SecretKey::from_bytes(sk).sign_double(nonce, msg).to_bytes() */
int sb200_sign_double_bytes(sb200_ctx* ctx, int64_t n, uint32_t flags, const uint8_t* sk32, const uint8_t* msg32,
                            const uint8_t* nonce32, uint8_t* sig96_out);
/* SecretKeyVarGen::from_bytes(sk64)?.sign(nonce, msg).to_bytes(); ok bit i = 0 where the generator does not decode */
int sb200_sign_vargen_bytes(sb200_ctx* ctx, int64_t n, uint32_t flags, const uint8_t* sk64, const uint8_t* msg32,
                            const uint8_t* nonce32, uint8_t* sig64_out, uint32_t* ok_bitmap);

/* Building-block probes for the parity tests (same kernels' device functions, one element per thread).
 * op: 0 mul (Montgomery product), 1 add, 2 sub, 3 inverse (b ignored), 4 square, 5 to_mont, 6 from_mont */
int sb200_dbg_fq(sb200_ctx* ctx, int64_t n, int op, const uint32_t* a, const uint32_t* b, uint32_t* out);
/* canonical product a*b mod r */
int sb200_dbg_fr_mul(sb200_ctx* ctx, int64_t n, const uint32_t* a, const uint32_t* b, uint32_t* out);
/* Hades252 permutation of n states (5 field elements each, in place); dense != 0 runs the
 * reference-shaped dense partial rounds instead of the sparse factorisation. */
int sb200_dbg_hades(sb200_ctx* ctx, int64_t n, int dense, uint32_t* states);  /* dense: 0 = sparse IMAD, 1 = dense IMAD (reference-shaped), 2 = FP64 pipe */
/* out[i] = k[i] * P[i] (affine); base: 0 = fixed G, 1 = fixed G', 2 = variable (points given) */
int sb200_dbg_scalar_mul(sb200_ctx* ctx, int64_t n, uint32_t flags, int base, const uint32_t* points,
                         const uint32_t* k, uint32_t* out);

#ifdef __cplusplus
}
#endif
#endif /* SCHNORR_B200_H */
