//! Batch sign / verify on B200 for dusk-schnorr (`--features cuda`).
//!
//! NOT compiled in this image (no Rust toolchain); this is the reference-side half of the C ABI in
//! `include/schnorr_b200.h`.  Marshalling copies internal limbs and nothing else:
//!   * `BlsScalar.0: [u64; 4]` is Montgomery (R = 2^256) little-endian  = the ABI's field element;
//!   * `JubJubExtended::{get_u, get_v, get_z}` give the projective triple = the ABI's PROJECTIVE point;
//!   * `JubJubScalar::to_bytes()` is the canonical scalar                = the ABI's scalar.
//! Verdicts / signatures are bit-exact with the single-tuple methods of the crate (src/keys/public.rs:121-130,
//! src/keys/secret.rs:150-168) at the level the crate defines equality (projective points, canonical scalars).
#![cfg(feature = "cuda")]
#![allow(non_snake_case)]

use core::ffi::{c_char, c_int, c_void};
use dusk_bls12_381::BlsScalar;
use dusk_bytes::Serializable;
use dusk_jubjub::{JubJubAffine, JubJubExtended, JubJubScalar};
use ff::Field;
use rand_core::{CryptoRng, RngCore};

use crate::{PublicKey, SecretKey, Signature};
#[cfg(feature = "double")]
use crate::{PublicKeyDouble, SignatureDouble};
#[cfg(feature = "var_generator")]
use crate::{PublicKeyVarGen, SecretKeyVarGen, SignatureVarGen};

#[repr(C)]
pub struct Sb200Ctx {
    _private: [u8; 0],
}
pub const SB200_POINTS_PROJECTIVE: u32 = 0;
pub const SB200_POINTS_AFFINE: u32 = 1;
pub const SB200_DEVICE_PTRS: u32 = 2;
/// `sb200_verify` only: warp-specialised kernel (hash warps on the FP64 pipe beside curve warps); same results.
pub const SB200_VERIFY_DUAL_PIPE: u32 = 4;

#[link(name = "schnorr_b200")]
extern "C" {
    fn sb200_init(devices: *const c_int, n_devices: c_int, out: *mut *mut Sb200Ctx) -> c_int;
    fn sb200_destroy(ctx: *mut Sb200Ctx);
    fn sb200_strerror(code: c_int) -> *const c_char;
    fn sb200_last_error(ctx: *const Sb200Ctx) -> *const c_char;
    fn sb200_device_count(ctx: *const Sb200Ctx) -> c_int;
    fn sb200_set_stream(ctx: *mut Sb200Ctx, stream: *mut c_void) -> c_int;
    fn sb200_launch_count(ctx: *const Sb200Ctx) -> u64;
    fn sb200_host_alloc(bytes: usize, out: *mut *mut c_void) -> c_int;
    fn sb200_host_free(p: *mut c_void);
    fn sb200_verify(ctx: *mut Sb200Ctx, n: i64, flags: u32, pk: *const u32, sig_u: *const u32, sig_R: *const u32,
                    msg: *const u32, verdicts: *mut u32, c_out: *mut u32) -> c_int;
    fn sb200_verify_double(ctx: *mut Sb200Ctx, n: i64, flags: u32, pk: *const u32, pk_prime: *const u32,
                           sig_u: *const u32, sig_R: *const u32, sig_R_prime: *const u32, msg: *const u32,
                           verdicts: *mut u32, c_out: *mut u32) -> c_int;
    fn sb200_verify_vargen(ctx: *mut Sb200Ctx, n: i64, flags: u32, pk: *const u32, generator: *const u32,
                           sig_u: *const u32, sig_R: *const u32, msg: *const u32, verdicts: *mut u32,
                           c_out: *mut u32) -> c_int;
    fn sb200_sign(ctx: *mut Sb200Ctx, n: i64, flags: u32, sk: *const u32, msg: *const u32, nonce: *const u32,
                  u_out: *mut u32, R_out: *mut u32, c_out: *mut u32) -> c_int;
    fn sb200_sign_double(ctx: *mut Sb200Ctx, n: i64, flags: u32, sk: *const u32, msg: *const u32, nonce: *const u32,
                         u_out: *mut u32, R_out: *mut u32, R_prime_out: *mut u32, c_out: *mut u32) -> c_int;
    fn sb200_sign_vargen(ctx: *mut Sb200Ctx, n: i64, flags: u32, sk: *const u32, generator: *const u32,
                         msg: *const u32, nonce: *const u32, u_out: *mut u32, R_out: *mut u32, c_out: *mut u32) -> c_int;
    fn sb200_keygen(ctx: *mut Sb200Ctx, n: i64, flags: u32, sk: *const u32, pk_out: *mut u32) -> c_int;
    fn sb200_keygen_double(ctx: *mut Sb200Ctx, n: i64, flags: u32, sk: *const u32, pk_out: *mut u32,
                           pk_prime_out: *mut u32) -> c_int;
    fn sb200_keygen_vargen(ctx: *mut Sb200Ctx, n: i64, flags: u32, sk: *const u32, generator: *const u32,
                           pk_out: *mut u32) -> c_int;
    fn sb200_points_decompress(ctx: *mut Sb200Ctx, n: i64, flags: u32, bytes32: *const u8, points_out: *mut u32,
                               ok_bitmap: *mut u32) -> c_int;
    fn sb200_points_compress(ctx: *mut Sb200Ctx, n: i64, flags: u32, points: *const u32, bytes32_out: *mut u8) -> c_int;
    fn sb200_scalars_from_wide(ctx: *mut Sb200Ctx, n: i64, flags: u32, field: c_int, wide64: *const u8,
                               out: *mut u32) -> c_int;
    fn sb200_fq_to_mont(ctx: *mut Sb200Ctx, n: i64, flags: u32, input: *const u32, out: *mut u32) -> c_int;
    fn sb200_fq_from_mont(ctx: *mut Sb200Ctx, n: i64, flags: u32, input: *const u32, out: *mut u32) -> c_int;
    fn sb200_verify_bytes(ctx: *mut Sb200Ctx, n: i64, flags: u32, pk32: *const u8, sig64: *const u8, msg32: *const u8,
                          verdicts: *mut u32, invalid: *mut u32) -> c_int;
    fn sb200_sign_bytes(ctx: *mut Sb200Ctx, n: i64, flags: u32, sk32: *const u8, msg32: *const u8, nonce32: *const u8,
                        sig64_out: *mut u8) -> c_int;
    // the double-key and variable-generator schemes in their `Serializable` forms
    // (pk64 = pk || pk' | pk || generator; sig96 = u || R || R'; sk64 = sk || generator)
    fn sb200_verify_double_bytes(ctx: *mut Sb200Ctx, n: i64, flags: u32, pk64: *const u8, sig96: *const u8,
                                 msg32: *const u8, verdicts: *mut u32, invalid: *mut u32) -> c_int;
    fn sb200_verify_vargen_bytes(ctx: *mut Sb200Ctx, n: i64, flags: u32, pk64: *const u8, sig64: *const u8,
                                 msg32: *const u8, verdicts: *mut u32, invalid: *mut u32) -> c_int;
    fn sb200_sign_double_bytes(ctx: *mut Sb200Ctx, n: i64, flags: u32, sk32: *const u8, msg32: *const u8,
                               nonce32: *const u8, sig96_out: *mut u8) -> c_int;
    fn sb200_sign_vargen_bytes(ctx: *mut Sb200Ctx, n: i64, flags: u32, sk64: *const u8, msg32: *const u8,
                               nonce32: *const u8, sig64_out: *mut u8, ok_bitmap: *mut u32) -> c_int;
}

/// Owns one `sb200_ctx` (comb tables of G and G' resident on each listed device).  There is no CPU
/// fallback: `new` fails without a B200.
pub struct CudaCtx {
    raw: *mut Sb200Ctx,
}
unsafe impl Send for CudaCtx {}
unsafe impl Sync for CudaCtx {} // calls on one context are serialised inside the library

#[derive(Debug)]
pub struct CudaError(pub i32, pub String);

impl CudaCtx {
    pub fn new(devices: &[i32]) -> Result<Self, CudaError> {
        let mut raw = core::ptr::null_mut();
        let rc = unsafe { sb200_init(devices.as_ptr(), devices.len() as c_int, &mut raw) };
        if rc != 0 {
            return Err(CudaError(rc, cstr(unsafe { sb200_strerror(rc) })));
        }
        Ok(Self { raw })
    }
    fn check(&self, rc: c_int) -> Result<(), CudaError> {
        if rc == 0 { Ok(()) } else { Err(CudaError(rc, cstr(unsafe { sb200_last_error(self.raw) }))) }
    }
}
impl Drop for CudaCtx {
    fn drop(&mut self) {
        unsafe { sb200_destroy(self.raw) }
    }
}
fn cstr(p: *const c_char) -> String {
    unsafe { std::ffi::CStr::from_ptr(p) }.to_string_lossy().into_owned()
}

// ---- marshalling: copies of internal limbs ---------------------------------------------------------
fn push_fq(v: &mut Vec<u32>, x: &BlsScalar) {
    for w in x.0 {
        v.push(w as u32);
        v.push((w >> 32) as u32);
    }
}
fn push_point(v: &mut Vec<u32>, p: &JubJubExtended) {
    push_fq(v, &p.get_u());
    push_fq(v, &p.get_v());
    push_fq(v, &p.get_z());
}
fn push_scalar(v: &mut Vec<u32>, s: &JubJubScalar) {
    for c in s.to_bytes().chunks_exact(4) {
        v.push(u32::from_le_bytes([c[0], c[1], c[2], c[3]]));
    }
}
fn fq_at(a: &[u32], i: usize) -> BlsScalar {
    let w = &a[8 * i..8 * i + 8];
    BlsScalar([
        w[0] as u64 | (w[1] as u64) << 32,
        w[2] as u64 | (w[3] as u64) << 32,
        w[4] as u64 | (w[5] as u64) << 32,
        w[6] as u64 | (w[7] as u64) << 32,
    ])
}
fn point_at(a: &[u32], i: usize) -> JubJubExtended {
    JubJubAffine::from_raw_unchecked(fq_at(a, 2 * i), fq_at(a, 2 * i + 1)).into()
}
fn scalar_at(a: &[u32], i: usize) -> JubJubScalar {
    let mut b = [0u8; 32];
    for (k, w) in a[8 * i..8 * i + 8].iter().enumerate() {
        b[4 * k..4 * k + 4].copy_from_slice(&w.to_le_bytes());
    }
    JubJubScalar::from_bytes(&b).expect("the library returns canonical scalars")
}
fn bits(words: &[u32], n: usize) -> Vec<bool> {
    (0..n).map(|i| (words[i >> 5] >> (i & 31)) & 1 == 1).collect()
}

impl PublicKey {
    /// `out[i] == pks[i].verify(&sigs[i], msgs[i])` (src/keys/public.rs:121-130) for every i.
    pub fn verify_batch(ctx: &CudaCtx, pks: &[PublicKey], sigs: &[Signature], msgs: &[BlsScalar]) -> Result<Vec<bool>, CudaError> {
        let n = pks.len();
        assert!(sigs.len() == n && msgs.len() == n);
        let (mut pk, mut u, mut r, mut m) = (Vec::with_capacity(24 * n), Vec::with_capacity(8 * n), Vec::with_capacity(24 * n), Vec::with_capacity(8 * n));
        for i in 0..n {
            push_point(&mut pk, pks[i].as_ref());
            push_scalar(&mut u, sigs[i].u());
            push_point(&mut r, sigs[i].R());
            push_fq(&mut m, &msgs[i]);
        }
        let mut words = vec![0u32; (n + 31) / 32];
        ctx.check(unsafe {
            sb200_verify(ctx.raw, n as i64, SB200_POINTS_PROJECTIVE, pk.as_ptr(), u.as_ptr(), r.as_ptr(), m.as_ptr(),
                         words.as_mut_ptr(), core::ptr::null_mut())
        })?;
        Ok(bits(&words, n))
    }

    /// Byte-level form: `PublicKey::from_bytes(pk)?.verify(&Signature::from_bytes(sig)?, BlsScalar::from_bytes(msg)?)`;
    /// returns (verdicts, invalid) where `invalid[i]` marks tuples whose decoding would be `Err(InvalidData)`.
    pub fn verify_bytes_batch(ctx: &CudaCtx, pk32: &[[u8; 32]], sig64: &[[u8; 64]], msg32: &[[u8; 32]]) -> Result<(Vec<bool>, Vec<bool>), CudaError> {
        let n = pk32.len();
        let (mut v, mut inv) = (vec![0u32; (n + 31) / 32], vec![0u32; (n + 31) / 32]);
        ctx.check(unsafe {
            sb200_verify_bytes(ctx.raw, n as i64, 0, pk32.as_ptr() as *const u8, sig64.as_ptr() as *const u8,
                               msg32.as_ptr() as *const u8, v.as_mut_ptr(), inv.as_mut_ptr())
        })?;
        Ok((bits(&v, n), bits(&inv, n)))
    }

    /// `PublicKey::from(&sk)` for every key (src/keys/public.rs:61-67).
    pub fn from_secret_keys(ctx: &CudaCtx, sks: &[SecretKey]) -> Result<Vec<PublicKey>, CudaError> {
        let n = sks.len();
        let mut k = Vec::with_capacity(8 * n);
        sks.iter().for_each(|s| push_scalar(&mut k, s.as_ref()));
        let mut out = vec![0u32; 16 * n];
        ctx.check(unsafe { sb200_keygen(ctx.raw, n as i64, 0, k.as_ptr(), out.as_mut_ptr()) })?;
        Ok((0..n).map(|i| PublicKey::from(point_at(&out, i))).collect())
    }
}

impl SecretKey {
    /// Batch form of `sign` (src/keys/secret.rs:150-168).  Signature i consumes the rng's i-th
    /// `JubJubScalar::random` draw, in order, exactly like n consecutive `sign` calls.
    pub fn sign_batch<R: RngCore + CryptoRng>(ctx: &CudaCtx, sks: &[SecretKey], rng: &mut R, msgs: &[BlsScalar]) -> Result<Vec<Signature>, CudaError> {
        let n = sks.len();
        let (mut sk, mut m, mut nonce) = (Vec::with_capacity(8 * n), Vec::with_capacity(8 * n), Vec::with_capacity(8 * n));
        for i in 0..n {
            push_scalar(&mut sk, sks[i].as_ref());
            push_fq(&mut m, &msgs[i]);
            push_scalar(&mut nonce, &JubJubScalar::random(&mut *rng));
        }
        let (mut u, mut r) = (vec![0u32; 8 * n], vec![0u32; 16 * n]);
        ctx.check(unsafe {
            sb200_sign(ctx.raw, n as i64, 0, sk.as_ptr(), m.as_ptr(), nonce.as_ptr(), u.as_mut_ptr(), r.as_mut_ptr(),
                       core::ptr::null_mut())
        })?;
        Ok((0..n).map(|i| Signature::new(scalar_at(&u, i), point_at(&r, i))).collect())
    }

    #[cfg(feature = "double")]
    pub fn sign_double_batch<R: RngCore + CryptoRng>(ctx: &CudaCtx, sks: &[SecretKey], rng: &mut R, msgs: &[BlsScalar]) -> Result<Vec<SignatureDouble>, CudaError> {
        let n = sks.len();
        let (mut sk, mut m, mut nonce) = (Vec::with_capacity(8 * n), Vec::with_capacity(8 * n), Vec::with_capacity(8 * n));
        for i in 0..n {
            push_scalar(&mut sk, sks[i].as_ref());
            push_fq(&mut m, &msgs[i]);
            push_scalar(&mut nonce, &JubJubScalar::random(&mut *rng));
        }
        let (mut u, mut r, mut rp) = (vec![0u32; 8 * n], vec![0u32; 16 * n], vec![0u32; 16 * n]);
        ctx.check(unsafe {
            sb200_sign_double(ctx.raw, n as i64, 0, sk.as_ptr(), m.as_ptr(), nonce.as_ptr(), u.as_mut_ptr(),
                              r.as_mut_ptr(), rp.as_mut_ptr(), core::ptr::null_mut())
        })?;
        Ok((0..n).map(|i| SignatureDouble::new(scalar_at(&u, i), point_at(&r, i), point_at(&rp, i))).collect())
    }
}

#[cfg(feature = "double")]
impl PublicKeyDouble {
    /// Batch form of `verify` (src/keys/public.rs:222-244).
    pub fn verify_batch(ctx: &CudaCtx, pks: &[PublicKeyDouble], sigs: &[SignatureDouble], msgs: &[BlsScalar]) -> Result<Vec<bool>, CudaError> {
        let n = pks.len();
        let (mut pk, mut pkp, mut u, mut r, mut rp, mut m) = (vec![], vec![], vec![], vec![], vec![], vec![]);
        for i in 0..n {
            push_point(&mut pk, pks[i].pk());
            push_point(&mut pkp, pks[i].pk_prime());
            push_scalar(&mut u, sigs[i].u());
            push_point(&mut r, sigs[i].R());
            push_point(&mut rp, sigs[i].R_prime());
            push_fq(&mut m, &msgs[i]);
        }
        let mut words = vec![0u32; (n + 31) / 32];
        ctx.check(unsafe {
            sb200_verify_double(ctx.raw, n as i64, SB200_POINTS_PROJECTIVE, pk.as_ptr(), pkp.as_ptr(), u.as_ptr(),
                                r.as_ptr(), rp.as_ptr(), m.as_ptr(), words.as_mut_ptr(), core::ptr::null_mut())
        })?;
        Ok(bits(&words, n))
    }
}

#[cfg(feature = "var_generator")]
impl PublicKeyVarGen {
    /// Batch form of `verify` (src/keys/public.rs:401-415).
    pub fn verify_batch(ctx: &CudaCtx, pks: &[PublicKeyVarGen], sigs: &[SignatureVarGen], msgs: &[BlsScalar]) -> Result<Vec<bool>, CudaError> {
        let n = pks.len();
        let (mut pk, mut g, mut u, mut r, mut m) = (vec![], vec![], vec![], vec![], vec![]);
        for i in 0..n {
            push_point(&mut pk, pks[i].public_key());
            push_point(&mut g, pks[i].generator());
            push_scalar(&mut u, sigs[i].u());
            push_point(&mut r, sigs[i].R());
            push_fq(&mut m, &msgs[i]);
        }
        let mut words = vec![0u32; (n + 31) / 32];
        ctx.check(unsafe {
            sb200_verify_vargen(ctx.raw, n as i64, SB200_POINTS_PROJECTIVE, pk.as_ptr(), g.as_ptr(), u.as_ptr(),
                                r.as_ptr(), m.as_ptr(), words.as_mut_ptr(), core::ptr::null_mut())
        })?;
        Ok(bits(&words, n))
    }
}

#[cfg(feature = "var_generator")]
impl SecretKeyVarGen {
    /// Batch form of `sign` (src/keys/secret.rs:433-451).
    pub fn sign_batch<R: RngCore + CryptoRng>(ctx: &CudaCtx, sks: &[SecretKeyVarGen], rng: &mut R, msgs: &[BlsScalar]) -> Result<Vec<SignatureVarGen>, CudaError> {
        let n = sks.len();
        let (mut sk, mut g, mut m, mut nonce) = (vec![], vec![], vec![], vec![]);
        for i in 0..n {
            push_scalar(&mut sk, sks[i].secret_key());
            push_point(&mut g, sks[i].generator());
            push_fq(&mut m, &msgs[i]);
            push_scalar(&mut nonce, &JubJubScalar::random(&mut *rng));
        }
        let (mut u, mut r) = (vec![0u32; 8 * n], vec![0u32; 16 * n]);
        ctx.check(unsafe {
            sb200_sign_vargen(ctx.raw, n as i64, SB200_POINTS_PROJECTIVE, sk.as_ptr(), g.as_ptr(), m.as_ptr(),
                              nonce.as_ptr(), u.as_mut_ptr(), r.as_mut_ptr(), core::ptr::null_mut())
        })?;
        Ok((0..n).map(|i| SignatureVarGen::new(scalar_at(&u, i), point_at(&r, i))).collect())
    }
}
