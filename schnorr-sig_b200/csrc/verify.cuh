// Per-signature building blocks of the verification path, written once and instantiated by the
// CUDA kernels (schnorr_b200.cu) and, for CPU-side formula tests only, by tests/hostsim.
//   Signature::verify            src/signature.rs:181-205
//   KeyPair::sign                src/signature.rs:114-129   (device signer: synthetic inputs, §8 f1)
//   PublicKey::from(&PrivateKey) src/public.rs:26-32
#pragma once
#include "affine.cuh"
#include "curve.cuh"
#include "rescue.cuh"

namespace sb {

enum verdict_t : uint8_t {
    VERDICT_OK = 0,
    VERDICT_INVALID_PUBLIC_KEY = 1,  // SignatureError::InvalidPublicKey  src/error.rs:15
    VERDICT_INVALID_SIGNATURE = 2,   // SignatureError::InvalidSignature  src/error.rs:17
    VERDICT_MALFORMED = 3,           // inputs on which the reference panics / that its types cannot hold
    VERDICT_NEEDS_EXACT = 0xff,      // internal: the fast path met an exceptional case (never returned to callers)
};

// Everything after the challenge hash: subgroup check, h*P + e*G, x-only comparison.
// `x_ok` = sig.x limbs canonical (the reference unwraps Fp6::from_bytes AFTER the subgroup check,
// src/signature.rs:182-186, so an off-subgroup key wins over a malformed x).
SB_DEV uint8_t verify_points(const fp6& sig_x, bool x_ok, const scalar& e, const fp6& pk_x, const fp6& pk_y, bool pk_inf,
                             const scalar& h, const uint64_t* __restrict__ gtab, jac_pt* d_storage) {
    // [q]P and h*P share one doubling chain (curve.cuh: torsion_check_and_mul); then + e*G from the
    // fixed-base table (multiply_double_with_basepoint_vartime, src/signature.rs:196-198)
    jac_pt r;
    if (!torsion_check_and_mul(jac_from_affine(pk_x, pk_y, pk_inf), h, &r, d_storage)) return VERDICT_INVALID_PUBLIC_KEY;
    if (!x_ok) return VERDICT_MALFORMED;
    fixed_base_accumulate(&r, e, gtab);
    return jac_x_equals(r, sig_x) ? VERDICT_OK : VERDICT_INVALID_SIGNATURE;
}

// Same verdicts through the fast path in (X, Y, w) coordinates (affine.cuh).  VERDICT_NEEDS_EXACT asks the caller to run
// verify_points on this item: identity key, or an exceptional case of the chord-and-tangent formulas.  `d_storage` /
// `bh_storage`: the caller's slots for the running point and the challenge buckets (affine.cuh: verify_core_fast).
SB_DEV uint8_t verify_points_fast(const fp6& sig_x, bool x_ok, const scalar& e, const fp6& pk_x, const fp6& pk_y, bool pk_inf,
                                  const scalar& h, const uint64_t* __restrict__ gtab, jf_pt* d_storage, jf_pt* bh_storage,
                                  int bh_stride) {
    if (pk_inf) return VERDICT_NEEDS_EXACT;
    jf_pt r;
    int fr = verify_core_fast(pk_x, pk_y, h, e, gtab, &r, d_storage, bh_storage, bh_stride);
    if (fr == FAST_EXCEPTIONAL) return VERDICT_NEEDS_EXACT;
    if (fr == FAST_NOT_TORSION_FREE) return VERDICT_INVALID_PUBLIC_KEY;
    if (!x_ok) return VERDICT_MALFORMED;
    // x(R) == sig.x  <=>  X == sig.x w^2   (src/signature.rs:200)
    return fp6_eq(r.X, fp6_scale(sig_x, fp_sqr_nc(r.w))) ? VERDICT_OK : VERDICT_INVALID_SIGNATURE;
}

// Challenge: h = Scalar::from_bits_vartime(hash_message(R.x, P, m))  (src/signature.rs:188-192).
// The identity public key hashes as x = y = 0 (its in-memory coordinates).
SB_DEV scalar challenge_scalar(const fp6& sig_x, const fp6& pk_x, const fp6& pk_y, bool pk_inf, const uint8_t* msg,
                               uint64_t len, bool sync = false) {
    fp6 px = pk_inf ? fp6_zero() : pk_x;
    fp_t py0 = pk_inf ? 0 : pk_y.c[0];
    fp_t d[4];
    hash_message(sig_x, px, py0, msg, len, d, sync);
    return digest_to_scalar(d);
}

}  // namespace sb
