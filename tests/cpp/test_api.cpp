// The reference's own tests, re-written against the C++ mirror of its API (include/schnorr_b200.hpp):
//   /root/reference/tests/schnorr.rs, schnorr_double.rs, schnorr_var_generator.rs, keys.rs
// Prints `name: ok` lines and, for the parity check done by pytest against the oracle, the hex of the
// seeded signature.  Exit code != 0 on any failed assertion.
#include <cstdio>
#include <cstdlib>

#include "../../include/schnorr_b200.hpp"
using namespace dusk_schnorr;

#define CHECK(cond)                                                   \
  do {                                                                \
    if (!(cond)) {                                                    \
      std::fprintf(stderr, "FAILED %s:%d: %s\n", __FILE__, __LINE__, #cond); \
      std::exit(1);                                                   \
    }                                                                 \
  } while (0)

template <size_t N>
static void hex(const char* tag, const std::array<uint8_t, N>& b) {
  std::printf("%s=", tag);
  for (uint8_t x : b) std::printf("%02x", x);
  std::printf("\n");
}

static void schnorr_sign_verify() {
  StdRng rng = StdRng::seed_from_u64(2321);
  SecretKey sk = SecretKey::random(rng);
  BlsScalar message = BlsScalar::random(rng);
  PublicKey pk = PublicKey::from(sk);
  Signature sig = sk.sign(rng, message);
  CHECK(pk.verify(sig, message));
  hex("sk", sk.to_bytes());
  hex("msg", message.to_bytes());
  hex("pk", pk.to_bytes());
  hex("sig", sig.to_bytes());
  std::printf("sign_verify: ok\n");
}
static void schnorr_wrong_keys() {
  StdRng rng = StdRng::seed_from_u64(2321);
  SecretKey sk = SecretKey::random(rng);
  BlsScalar message = BlsScalar::random(rng);
  Signature sig = sk.sign(rng, message);
  PublicKey pk = PublicKey::from(SecretKey::random(rng));
  CHECK(!pk.verify(sig, message));
  std::printf("test_wrong_keys: ok\n");
}
static void schnorr_to_from_bytes() {
  StdRng rng = StdRng::seed_from_u64(2321);
  SecretKey sk = SecretKey::random(rng);
  BlsScalar message = BlsScalar::random(rng);
  Signature sig = sk.sign(rng, message);
  auto b = sig.to_bytes();
  CHECK(sig == Signature::from_bytes(b.data(), b.size()));
  CHECK(SecretKey::from_bytes(sk.to_bytes()) == sk);
  CHECK(PublicKey::from_bytes(PublicKey::from(sk).to_bytes()) == PublicKey::from(sk));
  CHECK(BlsScalar::from_bytes(message.to_bytes()) == message);
  // error behaviour: non-canonical scalar, bytes that are not a point, wrong length
  Bytes32 r_bytes = {0xb7, 0x2c, 0xf7, 0xd6, 0x5e, 0x0e, 0x97, 0xd0, 0x82, 0x10, 0xc8, 0xcc, 0x93, 0x20, 0x68, 0xa6,
                     0x00, 0x3b, 0x34, 0x01, 0x01, 0x3b, 0x67, 0x06, 0xa9, 0xaf, 0x33, 0x65, 0xea, 0xb4, 0x7d, 0x0e};
  bool threw = false;
  try { SecretKey::from_bytes(r_bytes); } catch (const BytesError& e) { threw = e.kind == BytesError::InvalidData; }
  CHECK(threw);
  threw = false;
  Bytes32 not_point{};  // v = 2 is not on the curve
  not_point[0] = 2;
  try { PublicKey::from_bytes(not_point); } catch (const BytesError& e) { threw = e.kind == BytesError::InvalidData; }
  CHECK(threw);
  threw = false;
  try { Signature::from_bytes(b.data(), 63); } catch (const BytesError& e) { threw = e.kind == BytesError::BadLength; }
  CHECK(threw);
  std::printf("to_from_bytes: ok\n");
}
static void schnorr_double() {
  StdRng rng = StdRng::seed_from_u64(2321);
  SecretKey sk = SecretKey::random(rng);
  BlsScalar message = BlsScalar::random(rng);
  PublicKeyDouble pk = PublicKeyDouble::from(sk);
  SignatureDouble sig = sk.sign_double(rng, message);
  CHECK(pk.verify(sig, message));
  CHECK(!PublicKeyDouble::from(SecretKey::random(rng)).verify(sig, message));
  auto b = sig.to_bytes();
  CHECK(sig == SignatureDouble::from_bytes(b.data(), b.size()));
  auto pb = pk.to_bytes();
  CHECK(PublicKeyDouble::from_bytes(pb.data(), pb.size()) == pk);
  hex("sig_double", b);
  std::printf("double: ok\n");
}
static void schnorr_var_generator() {
  StdRng rng = StdRng::seed_from_u64(2321);
  SecretKeyVarGen sk = SecretKeyVarGen::random(rng);
  BlsScalar message = BlsScalar::random(rng);
  PublicKeyVarGen pk = PublicKeyVarGen::from(sk);
  SignatureVarGen sig = sk.sign(rng, message);
  CHECK(pk.verify(sig, message));
  CHECK(!PublicKeyVarGen::from(SecretKeyVarGen::random(rng)).verify(sig, message));
  auto b = sig.to_bytes();
  Signature back = Signature::from_bytes(b.data(), b.size());
  CHECK(back == static_cast<const Signature&>(sig));
  auto kb = sk.to_bytes();
  CHECK(SecretKeyVarGen::from_bytes(kb.data(), kb.size()) == sk);
  auto pb = pk.to_bytes();
  CHECK(PublicKeyVarGen::from_bytes(pb.data(), pb.size()) == pk);
  hex("sig_vargen", b);
  hex("vargen_generator", sk.generator().to_bytes());
  std::printf("var_generator: ok\n");
}
static void keys_partial_eq() {
  // keys.rs:19-60: equality is projective -- the same point reached along different paths, and in different
  // (U, V, Z) representatives, compares equal; a different point does not.
  auto scalar = [](uint64_t k) { JubJubScalar s; s.l[0] = (uint32_t)k; s.l[1] = (uint32_t)(k >> 32); return SecretKey(s); };
  PublicKey p9 = PublicKey::from(scalar(9));
  // 9G as PublicKeyVarGen with generator 3G and secret 3  (another path to the same point)
  PublicKey g3 = PublicKey::from(scalar(3));
  SecretKeyVarGen vg(scalar(3).as_ref(), g3.as_ref());
  PublicKeyVarGen pv = PublicKeyVarGen::from(vg);
  CHECK(PublicKey(pv.public_key()) == p9);
  CHECK(PublicKey::from(scalar(567758789)) != p9);
  // from_bytes gives Z = 1, keygen gives whatever the kernel produced: equal as points
  CHECK(PublicKey::from_bytes(p9.to_bytes()) == p9);
  std::printf("partial_eq: ok\n");
}
static void batch_matches_single() {
  StdRng a = StdRng::seed_from_u64(77), b = StdRng::seed_from_u64(77);
  std::vector<SecretKey> sks;
  std::vector<BlsScalar> msgs;
  for (int i = 0; i < 5; i++) { sks.push_back(SecretKey::random(a)); SecretKey::random(b); }
  for (int i = 0; i < 5; i++) { msgs.push_back(BlsScalar::random(a)); BlsScalar::random(b); }
  auto batch = SecretKey::sign_batch(sks, a, msgs);
  for (int i = 0; i < 5; i++) CHECK(batch[i] == sks[i].sign(b, msgs[i]));
  auto pks = PublicKey::from_batch(sks);
  auto ok = PublicKey::verify_batch(pks, batch, msgs);
  for (int i = 0; i < 5; i++) CHECK(ok[i]);
  std::vector<PublicKey> rot(pks.begin() + 1, pks.end());
  rot.push_back(pks[0]);
  auto bad = PublicKey::verify_batch(rot, batch, msgs);
  for (int i = 0; i < 5; i++) CHECK(!bad[i]);
  std::printf("batch: ok\n");
}

int main() {
  StdRng t = StdRng(std::array<uint8_t, 32>{1, 0, 0, 0, 23, 0, 0, 0, 200, 1, 0, 0, 210, 30, 0, 0});
  CHECK(t.next_u64() == 10719222850664546238ULL);  // rand 0.8 test_stdrng_construction
  schnorr_sign_verify();
  schnorr_wrong_keys();
  schnorr_to_from_bytes();
  schnorr_double();
  schnorr_var_generator();
  keys_partial_eq();
  batch_matches_single();
  std::printf("ALL OK\n");
  return 0;
}
