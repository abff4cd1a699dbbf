//! `cargo run --release > upstream_dump.json`, then `python tools/gen_params.py --from-dump upstream_dump.json`
//! regenerates `params/params.json` + `include/cheetah_params.h` (every tag becomes REF) and `tests/golden/`
//! gains the upstream known answers; the oracle pins them in `tests/test_oracle_pins.py`.
//!
//! What is dumped = exactly the list DESIGN.md §3 calls "not pinned":
//!   * the generator G (affine limbs + compressed bytes, so the y-sign flag convention shows);
//!   * BASEPOINT_TABLE spot values: 2G, 3G, (2^13)G;
//!   * the Rescue instance observed as a black box: digests of a few inputs chosen to separate padding rules
//!     (empty, 1, 7, 8, 9 elements; all-zero and all-one states);
//!   * seeded signatures (ChaCha20 seed 0..4) over messages of 0, 1, 7, 8, 80, 160 bytes with their public keys,
//!     `hash_message` digests are implied by the signatures but the challenge scalar is dumped too;
//!   * `from_compressed` on both flag values of one point.

use cheetah::{AffinePoint, CompressedPoint, Fp, Scalar, BASEPOINT_TABLE};
use hash::{rescue_64_12_8::RescueHash, traits::Hasher};
use rand_chacha::ChaCha20Rng;
use rand_core::SeedableRng;
use schnorr_sig::KeyPair;

fn hex(bytes: &[u8]) -> String {
    bytes.iter().map(|b| format!("{b:02x}")).collect()
}

fn point_json(p: &AffinePoint) -> String {
    format!("{{\"x\": \"{}\", \"y\": \"{}\", \"compressed\": \"{}\"}}", hex(&p.get_x().to_bytes()), hex(&p.get_y().to_bytes()),
            hex(&p.to_compressed().to_bytes()))
}

fn main() {
    let g = AffinePoint::generator();
    println!("{{");
    println!(" \"generator\": {},", point_json(&g));
    for (name, k) in [("2G", 2u64), ("3G", 3), ("8192G", 8192)] {
        println!(" \"{name}\": {},", point_json(&AffinePoint::from(&BASEPOINT_TABLE * Scalar::from(k))));
    }
    // Rescue as a black box
    println!(" \"rescue\": [");
    let cases: Vec<Vec<Fp>> = vec![vec![], vec![Fp::one()], vec![Fp::one(); 7], vec![Fp::one(); 8], vec![Fp::one(); 9],
                                   vec![Fp::zero(); 8], (1..=13u64).map(Fp::new).collect()];
    for (i, c) in cases.iter().enumerate() {
        let d = RescueHash::hash_field(c);
        println!("  {{\"input_len\": {}, \"digest\": \"{}\"}}{}", c.len(), hex(&d.to_bytes()), if i + 1 < cases.len() { "," } else { "" });
    }
    println!(" ],");
    // seeded signatures
    println!(" \"signatures\": [");
    let lens = [0usize, 1, 7, 8, 80, 160];
    for (i, len) in lens.iter().enumerate() {
        let mut rng = ChaCha20Rng::seed_from_u64(i as u64);
        let kp = KeyPair::new(&mut rng);
        let msg: Vec<u8> = (0..*len).map(|k| (k * 7 + i) as u8).collect();
        let sig = kp.sign(&msg, &mut rng);
        assert!(sig.verify(&msg, &kp.public_key).is_ok());
        println!("  {{\"msg\": \"{}\", \"private_key\": \"{}\", \"public_key\": {}, \"signature\": \"{}\"}}{}", hex(&msg),
                 hex(&kp.private_key.to_bytes()), point_json(&kp.public_key.0), hex(&sig.to_bytes()),
                 if i + 1 < lens.len() { "," } else { "" });
    }
    println!(" ],");
    // y-sign flag: both decompressions of the generator's x
    let mut c = g.to_compressed().to_bytes();
    let a = AffinePoint::from_compressed(&CompressedPoint::from_bytes(&c)).unwrap();
    c[48] ^= 0x40;
    let b = AffinePoint::from_compressed(&CompressedPoint::from_bytes(&c));
    println!(" \"flag_flip_decodes\": {}, \"flag_flip_is_negation\": {}", bool::from(b.is_some()),
             bool::from(b.is_some()) && b.unwrap() == -a);
    println!("}}");
}
