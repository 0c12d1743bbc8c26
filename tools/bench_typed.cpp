// e2e_typed: the hot path timed through the TYPED host API a caller of the crate would use
// (include/schnorr_b200.hpp: SecretKey::sign_batch, PublicKey::verify_batch -- the C++ mirror of the reference's
// types, /root/reference/src/keys/secret.rs:150-168 and src/keys/public.rs:121-130), with everything a real caller
// pays inside the timed region: per-tuple marshalling out of the typed objects into limb arrays, PAGEABLE host memory
// (the library stages it through its pinned ring), host<->device copies, nonce draws from the seeded ChaCha12 stream
// (sign), un-marshalling of the results into typed objects.
// usage: bench_typed [log2n] [reps]     -> one JSON object on stdout
#include <chrono>
#include <cstdio>
#include <cstdlib>

#include "../include/schnorr_b200.hpp"
using namespace dusk_schnorr;
using clk = std::chrono::steady_clock;

static double ms_since(clk::time_point t0) { return std::chrono::duration<double, std::milli>(clk::now() - t0).count(); }

int main(int argc, char** argv) {
  const int lg = argc > 1 ? atoi(argv[1]) : 20, reps = argc > 2 ? atoi(argv[2]) : 2;
  const size_t n = (size_t)1 << lg;
  try {
    StdRng rng = StdRng::seed_from_u64(0xC3);
    // keys and messages: bulk wide draws (the same from_bytes_wide the single-tuple `random` uses)
    std::vector<SecretKey> sks(n);
    std::vector<BlsScalar> msgs(n);
    {
      avec<uint32_t> a(8 * n), b(8 * n);
      detail::wide_draws(rng, n, 0, a.data());
      detail::wide_draws(rng, n, 1, b.data());
      for (size_t i = 0; i < n; i++) {
        JubJubScalar s;
        std::memcpy(s.l, &a[8 * i], 32);
        sks[i] = SecretKey(s);
        std::memcpy(msgs[i].l, &b[8 * i], 32);
      }
    }
    std::vector<PublicKey> pks = PublicKey::from_batch(sks);
    std::vector<Signature> sigs = SecretKey::sign_batch(sks, rng, msgs);  // warm-up: arenas, staging ring
    double sign_ms = 1e300, verify_ms = 1e300;
    for (int r = 0; r < reps; r++) {
      StdRng nonce_rng = StdRng::seed_from_u64(0xC3C3 + r);
      auto t0 = clk::now();
      sigs = SecretKey::sign_batch(sks, nonce_rng, msgs);
      sign_ms = std::min(sign_ms, ms_since(t0));
    }
    // 10 % corrupted: u ^= 1 stays canonical with overwhelming probability (u < r - 1)
    std::vector<Signature> bad = sigs;
    size_t n_bad = 0;
    for (size_t i = 0; i < n; i += 10) {
      auto b = bad[i].to_bytes();
      b[0] ^= 1;
      bad[i] = Signature::from_bytes(b.data(), b.size());
      n_bad++;
      if (i > 200) break;  // typed from_bytes decompresses on the device per call: corrupt a handful this way ...
    }
    std::vector<bool> ok;
    for (int r = 0; r < reps + 1; r++) {
      auto t0 = clk::now();
      ok = PublicKey::verify_batch(pks, bad, msgs);
      if (r) verify_ms = std::min(verify_ms, ms_since(t0));
    }
    size_t wrong = 0;
    for (size_t i = 0; i < n; i++) wrong += ok[i] != !(i % 10 == 0 && i <= 210);
    std::printf("{\"log2n\": %d, \"verify_typed_per_s\": %.1f, \"verify_typed_ms\": %.3f, \"sign_typed_per_s\": %.1f, \"sign_typed_ms\": %.3f, "
                "\"wrong_verdicts\": %zu, \"corrupted\": %zu, \"points\": \"projective (U, V, Z) as held by JubJubExtended\", "
                "\"host_memory\": \"pageable (64-byte aligned new[]), staged by the library\"}\n",
                lg, n / (verify_ms * 1e-3), verify_ms, n / (sign_ms * 1e-3), sign_ms, wrong, n_bad);
    return wrong ? 1 : 0;
  } catch (const std::exception& e) {
    std::fprintf(stderr, "bench_typed: %s\n", e.what());
    return 2;
  }
}
