"""The reference's own behavioural tests, re-run against the oracle with the reference's recipe
(`StdRng::seed_from_u64(2321)`, draw order sk -> message -> nonce):
    /root/reference/tests/schnorr.rs:15-51, schnorr_double.rs:15-52, schnorr_var_generator.rs:15-51,
    /root/reference/tests/keys.rs:19-127 (projective equality).
The GPU twins of these tests are in tests/test_gpu_api.py."""
import schnorr_oracle as o

Q, R = o.Q, o.R


def _setup():
    rng = o.StdRng.seed_from_u64(2321)
    sk = rng.random_fr()
    message = rng.random_fq()
    return rng, sk, message


def test_sign_verify():
    rng, sk, message = _setup()
    pk = o.keygen(sk)
    u, Rp, _ = o.sign(sk, rng.random_fr(), message)
    assert o.verify(pk, u, Rp, message)


def test_wrong_keys():
    rng, sk, message = _setup()
    u, Rp, _ = o.sign(sk, rng.random_fr(), message)
    pk = o.keygen(rng.random_fr())
    assert not o.verify(pk, u, Rp, message)


def test_to_from_bytes():
    rng, sk, message = _setup()
    u, Rp, _ = o.sign(sk, rng.random_fr(), message)
    b = u.to_bytes(32, "little") + o.affine_to_bytes(Rp)
    assert len(b) == 64
    assert (o.scalar_from_bytes(b[:32]), o.affine_from_bytes(b[32:])) == (u, Rp)


def test_double_sign_verify_wrong_keys():
    rng, sk, message = _setup()
    pk, pkp = o.keygen_double(sk)
    u, Rp, Rpp, _ = o.sign_double(sk, rng.random_fr(), message, mul=o.pt_mul_fast)
    assert o.verify_double(pk, pkp, u, Rp, Rpp, message, mul=o.pt_mul_fast)
    wrong = o.keygen_double(rng.random_fr())
    assert not o.verify_double(wrong[0], wrong[1], u, Rp, Rpp, message, mul=o.pt_mul_fast)


def test_var_generator_sign_verify_wrong_keys():
    rng = o.StdRng.seed_from_u64(2321)
    sk, s = rng.random_fr(), rng.random_fr()  # SecretKeyVarGen::random draws sk, then the generator scalar
    gen = o.pt_mul_fast(o.G, s)
    message = rng.random_fq()
    pk = o.keygen_vargen(sk, gen)
    u, Rp, _ = o.sign_vargen(sk, gen, rng.random_fr(), message, mul=o.pt_mul_fast)
    assert o.verify_vargen(pk, gen, u, Rp, message, mul=o.pt_mul_fast)
    sk2, s2 = rng.random_fr(), rng.random_fr()
    gen2 = o.pt_mul_fast(o.G, s2)
    assert not o.verify_vargen(o.keygen_vargen(sk2, gen2), gen2, u, Rp, message, mul=o.pt_mul_fast)


def test_partial_eq_is_projective():
    """keys.rs:33-58: 2G+7G == 4G+5G != 4G+567758785G; different (u, v, z), same affine image => equal"""
    g = lambda k: o.pt_mul_fast(o.G, k)
    left, right, wrong = o.pt_add(g(2), g(7)), o.pt_add(g(4), g(5)), o.pt_add(g(4), g(567758785))
    assert left == right and left != wrong
    z1, z2 = 12345, 67890
    p1 = (left[0] * z1 % Q, left[1] * z1 % Q, z1)
    p2 = (right[0] * z2 % Q, right[1] * z2 % Q, z2)
    assert p1[0] != p2[0] and p1[1] != p2[1] and p1[2] != p2[2]
    assert o.proj_eq(p1, p2) and not o.proj_eq(p1, (wrong[0], wrong[1], 1))
