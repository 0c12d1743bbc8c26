// Regenerates tests/golden/schnorr_golden_crate.json FROM THE REAL CRATE, in the schema of the two oracle-generated
// files next to it (tests/golden/make_golden.py), so that `python -m pytest tests/test_golden.py` can say which of the
// two recalled round-constant rules dusk-schnorr actually uses -- the step that turns "parity unpinned" into "pinned".
//
// NOT compiled in the build image of this repository (no Rust toolchain there).  To use it, on any machine with cargo:
//
//     cp <repo>/rust/tests/dump_golden.rs <dusk-schnorr 0.18>/tests/dump_golden.rs
//     cd <dusk-schnorr 0.18>
//     SB200_GOLDEN_DIR=<repo>/tests/golden cargo test --features double,var_generator --test dump_golden -- --nocapture
//     cd <repo> && python -m pytest tests/test_golden.py -k crate -s
//
// (add to Cargo.toml:  [[test]] name = "dump_golden"  required-features = ["double", "var_generator"])
//
// The recipe is the reference's own tests (tests/schnorr.rs:15-25, tests/schnorr_double.rs:15-25,
// tests/schnorr_var_generator.rs:15-25): StdRng::seed_from_u64(2321); sk = random; message = random; sign (one
// JubJubScalar::random draw) -- followed by 7 more tuples per scheme from one continuing stream (seeds 0xC1, 0xC2,
// 0xC4), 4 edge tuples (seed 0xED6E: msg = 0, msg = q - 1, sk = 1, sk = r - 1), and the crate's verdicts on the
// negative / small-order cases whose INPUTS are read from the two committed files.  Every byte string is the crate's
// own `to_bytes()`.  The nonce of a signature is recovered by cloning the rng before `sign` and drawing the same
// `JubJubScalar::random` (src/keys/secret.rs:155).
use dusk_bls12_381::BlsScalar;
use dusk_bytes::Serializable;
use dusk_jubjub::JubJubScalar;
use dusk_schnorr::{
    PublicKey, PublicKeyDouble, PublicKeyVarGen, SecretKey, SecretKeyVarGen, Signature,
};
use ff::Field;
use rand::rngs::StdRng;
use rand::SeedableRng;
use std::fmt::Write as _;

fn hx(b: &[u8]) -> String {
    let mut s = String::with_capacity(2 * b.len());
    for x in b {
        write!(s, "{:02x}", x).unwrap();
    }
    s
}
fn unhex(s: &str) -> Vec<u8> {
    (0..s.len() / 2).map(|i| u8::from_str_radix(&s[2 * i..2 * i + 2], 16).unwrap()).collect()
}

/// c = challenge of a signature, recomputed through the public API is not possible (`challenge_hash` is
/// pub(crate), src/signatures.rs:127); it follows from the scheme equation instead: u = r - c * sk  =>  c = (r - u) / sk.
fn challenge_from(sk: &JubJubScalar, nonce: &JubJubScalar, u: &JubJubScalar) -> JubJubScalar {
    (*nonce - *u) * sk.invert().unwrap()
}

fn single_case(rng: &mut StdRng, msg: Option<BlsScalar>, sk: Option<JubJubScalar>) -> String {
    let sk = match sk {
        Some(s) => SecretKey::from(s),
        None => SecretKey::random(rng),
    };
    let m = msg.unwrap_or_else(|| BlsScalar::random(&mut *rng));
    let nonce = JubJubScalar::random(&mut rng.clone());
    let sig = sk.sign(rng, m);
    let pk = PublicKey::from(&sk);
    assert!(pk.verify(&sig, m));
    let c = challenge_from(sk.as_ref(), &nonce, sig.u());
    format!(
        "  {{\n   \"sk\": \"{}\",\n   \"msg\": \"{}\",\n   \"nonce\": \"{}\",\n   \"pk\": \"{}\",\n   \"sig\": \"{}\",\n   \"c\": \"{}\",\n   \"valid\": true\n  }}",
        hx(&sk.to_bytes()), hx(&m.to_bytes()), hx(&nonce.to_bytes()), hx(&pk.to_bytes()), hx(&sig.to_bytes()), hx(&c.to_bytes())
    )
}

fn double_case(rng: &mut StdRng) -> String {
    let sk = SecretKey::random(rng);
    let m = BlsScalar::random(&mut *rng);
    let nonce = JubJubScalar::random(&mut rng.clone());
    let sig = sk.sign_double(rng, m);
    let pk = PublicKeyDouble::from(&sk);
    assert!(pk.verify(&sig, m));
    let c = challenge_from(sk.as_ref(), &nonce, sig.u());
    format!(
        "  {{\n   \"sk\": \"{}\",\n   \"msg\": \"{}\",\n   \"nonce\": \"{}\",\n   \"pk\": \"{}\",\n   \"sig\": \"{}\",\n   \"c\": \"{}\",\n   \"valid\": true\n  }}",
        hx(&sk.to_bytes()), hx(&m.to_bytes()), hx(&nonce.to_bytes()), hx(&pk.to_bytes()), hx(&sig.to_bytes()), hx(&c.to_bytes())
    )
}

fn vargen_case(rng: &mut StdRng) -> String {
    let sk = SecretKeyVarGen::random(rng); // two draws: sk, then the generator scalar (src/keys/secret.rs:371-373)
    let m = BlsScalar::random(&mut *rng);
    let nonce = JubJubScalar::random(&mut rng.clone());
    let sig = sk.sign(rng, m);
    let pk = PublicKeyVarGen::from(&sk);
    assert!(pk.verify(&sig, m));
    let c = challenge_from(sk.secret_key(), &nonce, sig.u());
    format!(
        "  {{\n   \"sk\": \"{}\",\n   \"msg\": \"{}\",\n   \"nonce\": \"{}\",\n   \"pk\": \"{}\",\n   \"sig\": \"{}\",\n   \"c\": \"{}\",\n   \"valid\": true\n  }}",
        hx(&sk.to_bytes()), hx(&m.to_bytes()), hx(&nonce.to_bytes()), hx(&pk.to_bytes()), hx(&sig.to_bytes()), hx(&c.to_bytes())
    )
}

/// value of `"key": "…"` inside one JSON object (the committed files are written by json.dump: no escapes in these fields)
fn field<'a>(obj: &'a str, key: &str) -> &'a str {
    let pat = format!("\"{}\": \"", key);
    let a = obj.find(&pat).unwrap() + pat.len();
    let b = obj[a..].find('"').unwrap();
    &obj[a..a + b]
}

/// the crate's verdicts on the `single_verify_cases` inputs of one committed file
fn verify_cases(path: &std::path::Path) -> Option<String> {
    let text = std::fs::read_to_string(path).ok()?;
    let start = text.find("\"single_verify_cases\": [")?;
    let body = &text[start..];
    let end = body.find(']')?;
    let mut out = Vec::new();
    for obj in body[..end].split('{').skip(1) {
        let (why, pk, sig, msg) = (field(obj, "why"), field(obj, "pk"), field(obj, "sig"), field(obj, "msg"));
        let mut b32 = [0u8; 32];
        let mut b64 = [0u8; 64];
        b32.copy_from_slice(&unhex(pk));
        let pkv = PublicKey::from_bytes(&b32).expect("committed cases decode");
        b64.copy_from_slice(&unhex(sig));
        let sigv = Signature::from_bytes(&b64).expect("committed cases decode");
        b32.copy_from_slice(&unhex(msg));
        let m = BlsScalar::from_bytes(&b32).expect("committed cases decode");
        out.push(format!(
            "  {{\n   \"why\": \"{}\",\n   \"pk\": \"{}\",\n   \"sig\": \"{}\",\n   \"msg\": \"{}\",\n   \"valid\": {}\n  }}",
            why, pk, sig, msg, pkv.verify(&sigv, m)
        ));
    }
    Some(out.join(",\n"))
}

#[test]
fn dump_golden() {
    let dir = std::path::PathBuf::from(std::env::var("SB200_GOLDEN_DIR").unwrap_or_else(|_| "../tests/golden".into()));
    let mut single = vec![single_case(&mut StdRng::seed_from_u64(2321), None, None)];
    let mut double = vec![double_case(&mut StdRng::seed_from_u64(2321))];
    let mut vargen = vec![vargen_case(&mut StdRng::seed_from_u64(2321))];
    let mut rng = StdRng::seed_from_u64(0xC1);
    for _ in 0..7 {
        single.push(single_case(&mut rng, None, None));
    }
    let mut rng = StdRng::seed_from_u64(0xC2);
    for _ in 0..7 {
        double.push(double_case(&mut rng));
    }
    let mut rng = StdRng::seed_from_u64(0xC4);
    for _ in 0..7 {
        vargen.push(vargen_case(&mut rng));
    }
    let mut rng = StdRng::seed_from_u64(0xED6E);
    single.push(single_case(&mut rng, Some(BlsScalar::zero()), None));
    single.push(single_case(&mut rng, Some(-BlsScalar::one()), None));
    single.push(single_case(&mut rng, None, Some(JubJubScalar::one())));
    single.push(single_case(&mut rng, None, Some(-JubJubScalar::one())));

    let mut json = String::from("{\n \"about\": \"generated by rust/tests/dump_golden.rs from the real dusk-schnorr crate\",\n \"ark\": \"crate\",\n");
    for (name, v) in [("single", &single), ("double", &double), ("vargen", &vargen)] {
        write!(json, " \"{}\": [\n{}\n ],\n", name, v.join(",\n")).unwrap();
    }
    for rule in ["cumsum", "plain"] {
        let cases = verify_cases(&dir.join(format!("schnorr_golden_{}.json", rule))).unwrap_or_default();
        write!(json, " \"single_verify_cases_{}\": [\n{}\n ],\n", rule, cases).unwrap();
    }
    json.push_str(" \"crate_version\": \"");
    json.push_str(env!("CARGO_PKG_VERSION"));
    json.push_str("\"\n}\n");
    let path = dir.join("schnorr_golden_crate.json");
    std::fs::write(&path, json).expect("cannot write the golden file (set SB200_GOLDEN_DIR)");
    println!("wrote {}", path.display());
}
