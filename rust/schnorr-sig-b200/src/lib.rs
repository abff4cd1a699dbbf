//! Safe Rust over `schnorr-sig-sys`, written against the reference's own types: the functions below are what the bodies
//! of `Signature::verify` (src/signature.rs:181-205), `PublicKey::verify_signature` (:170-176),
//! `KeyedSignature::verify` (:232-234) and `verify_batch` (src/batch.rs:31-50) become.  Signatures, argument meaning,
//! `Result` / panic behaviour are the reference's; only the arithmetic moves to the GPU.
//!
//! In-tree use: move `verify` / `verify_many` into `impl Signature`, `verify_batch` into `src/batch.rs`, and keep the
//! engine in a `thread_local!` (the C context is single-owner).

pub mod batch;
pub mod signature;

use cheetah::AffinePoint;
use schnorr_sig::PublicKey;
use schnorr_sig_sys::Engine;

thread_local! {
    /// One context per thread.  `SCHNORR_B200_DEVICES` is read by the APPLICATION (e.g. "0,1,2,3,4,5,6,7"), never by
    /// the C library.
    pub static ENGINE: Engine = {
        let devices: Vec<i32> = std::env::var("SCHNORR_B200_DEVICES").ok()
            .map(|s| s.split(',').filter_map(|d| d.trim().parse().ok()).collect())
            .filter(|v: &Vec<i32>| !v.is_empty())
            .unwrap_or_else(|| vec![0]);
        Engine::new(&devices).expect("schnorr_b200_create failed: a CUDA device is required (there is no CPU fallback)")
    };
}

/// `PublicKey(AffinePoint)` -> the 96-byte `x || y` record + identity flag of the C ABI (a copy of limbs, no arithmetic).
pub fn public_key_record(pkey: &PublicKey) -> ([u8; 96], u8) {
    let p: AffinePoint = pkey.0;
    let mut rec = [0u8; 96];
    rec[..48].copy_from_slice(&p.get_x().to_bytes());
    rec[48..].copy_from_slice(&p.get_y().to_bytes());
    (rec, bool::from(p.is_identity()) as u8)
}
