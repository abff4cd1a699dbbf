//! Replacement body of `Signature::verify` and a slice form for throughput callers.

use schnorr_sig::{PublicKey, Signature, SignatureError};

use crate::{public_key_record, ENGINE};

/// Maps a verdict byte of the C ABI onto the reference's `Result`.  Verdict 3 reproduces the panic of
/// `Fp6::from_bytes(..).unwrap()` at src/signature.rs:186.
fn verdict_to_result(v: u8) -> Result<(), SignatureError> {
    match v {
        0 => Ok(()),
        1 => Err(SignatureError::InvalidPublicKey),
        2 => Err(SignatureError::InvalidSignature),
        _ => panic!("called `Option::unwrap()` on a `None` value"),
    }
}

/// `Signature::verify(self, message, pkey)`.
pub fn verify(signature: Signature, message: &[u8], pkey: &PublicKey) -> Result<(), SignatureError> {
    let (pk, inf) = public_key_record(pkey);
    let v = ENGINE.with(|e| e.verify_many(&[signature.to_bytes()], &[pk], &[inf], &[message]));
    verdict_to_result(v[0])
}

/// Many independent verifications in one call (what a node does with a block of transactions): one result per triple,
/// in order.  A single-signature call pays a PCIe round trip and cannot fill 148 SMs; hand over slices.
pub fn verify_many(signatures: &[Signature], public_keys: &[PublicKey], messages: &[&[u8]]) -> Vec<Result<(), SignatureError>> {
    assert!(signatures.len() == public_keys.len(), "We should have the same number of signatures than public keys");
    assert!(messages.len() == public_keys.len(), "We should have the same number of messages than public keys");
    let sigs: Vec<[u8; 81]> = signatures.iter().map(|s| s.to_bytes()).collect();
    let (pks, infs): (Vec<[u8; 96]>, Vec<u8>) = public_keys.iter().map(public_key_record).unzip();
    ENGINE.with(|e| e.verify_many(&sigs, &pks, &infs, messages)).into_iter().map(verdict_to_result).collect()
}
