/* schnorr_b200 -- C ABI of the B200-native Schnorr verification engine (Cheetah curve / Rescue hash).
 *
 * This is the drop-in boundary for the hot path of toposware/schnorr-sig.  The reference has no FFI
 * (`#![deny(unsafe_code)]`, src/lib.rs:152); the entry points below are what a `schnorr-sig-sys`
 * layer under the reference's public API would bind (INTEGRATION.md shows the Rust side).  Each
 * function cites the reference interface it replaces.
 *
 * Conventions
 *  - plain pointers and sizes, no CUDA/torch types; `void *stream` is a cudaStream_t.
 *  - record layouts are the reference's own byte encodings:
 *      signature  81 B  = CompressedPoint (48 B little-endian limbs c0..c5 of R.x, 1 flag byte:
 *                         bit 7 infinity, bit 6 y-sign) || Scalar e (32 B LE)    src/signature.rs:208-214
 *      public key 96 B  = affine x (48 B) || y (48 B), the in-memory `PublicKey(AffinePoint)`
 *                         (src/public.rs:24) + one byte per key: 1 = identity
 *      scalar     32 B  little-endian                                            src/constants.rs:12
 *      messages   one byte blob + uint64 offsets[n+1]  (the reference's `&[&[u8]]`)
 *  - verdict bytes: 0 = Ok(()), 1 = Err(InvalidPublicKey), 2 = Err(InvalidSignature)
 *    (src/error.rs:13-18), 3 = malformed input on which the reference panics or which its types
 *    cannot represent (non-canonical limb / scalar >= q; src/signature.rs:186, src/batch.rs:67,104).
 *  - return value: 0 on success, negative SCHNORR_B200_E* on argument / CUDA failure.  An invalid
 *    signature is data (a verdict), never an error code.
 *  - `*_dev` variants take DEVICE pointers for every buffer, enqueue on the context stream and do
 *    not synchronise; the host variants copy in, run the same kernels, copy out and synchronise.
 *  - a context is single-owner: one in-flight call per context; create several for concurrency.
 *  - there is NO CPU fallback: without a CUDA device `schnorr_b200_create` fails.
 */
#ifndef SCHNORR_B200_H
#define SCHNORR_B200_H
#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define SCHNORR_B200_OK 0
#define SCHNORR_B200_EARG (-1)    /* bad argument */
#define SCHNORR_B200_ECUDA (-2)   /* CUDA runtime failure, see schnorr_b200_last_error */
#define SCHNORR_B200_ENODEV (-3)  /* no usable CUDA device */

#define SCHNORR_B200_SIGNATURE_BYTES 81   /* SIGNATURE_LENGTH   src/constants.rs:30 */
#define SCHNORR_B200_POINT_BYTES 96       /* affine x||y */
#define SCHNORR_B200_COMPRESSED_BYTES 49  /* PUBLIC_KEY_LENGTH  src/constants.rs:24 */
#define SCHNORR_B200_SCALAR_BYTES 32      /* SCALAR_LENGTH      src/constants.rs:12 */
#define SCHNORR_B200_DIGEST_BYTES 32
#define SCHNORR_B200_PARTIAL_BYTES 192    /* Jacobian point 144 B || partial scalar 32 B || flags 16 B */

typedef struct schnorr_b200_ctx schnorr_b200_ctx;

/* Value-level provenance of the curve / hash constants compiled into the library (include/cheetah_params.h).
 * 0 = NOT pinned to the upstream `cheetah` / `hash` crates: the reference's tests hold no generator, digest or signature
 * known answers and its dependencies are un-vendored git crates, so the generator is a placeholder and the Rescue
 * instance is recalled / spec-derived; every result is bit-exact against the restated oracle (oracle/), and keys /
 * signatures interoperate with the real crate only after rust/dump_params has been run on a machine with cargo and the
 * header regenerated.  schnorr_b200_create prints this once per process while it is 0. */
int schnorr_b200_params_pinned(void);
const char *schnorr_b200_params_provenance(void);

/* Creates a context on CUDA device `device`: stream, scratch arena and the fixed-base table of G
 * (the reference's `cheetah::BASEPOINT_TABLE`, src/signature.rs:19) built on the device. */
int schnorr_b200_create(int device, schnorr_b200_ctx **out);
/* Multi-device context: one stream, scratch arena and fixed-base table per listed CUDA device.
 * The HOST entry points (verify_many, verify_keyed_many, hash_messages, verify_batch, keygen, sign_many) then shard
 * every call over the devices in contiguous slices -- independent verification without any exchange, batch
 * verification with ONE 192-byte peer copy per device (partial MSM point + partial scalar) to the first device, where
 * the batch is finished.  This is what `verify_batch(&[Signature], &[PublicKey], &[&[u8]], rng)` (src/batch.rs:31-36)
 * and a loop over `Signature::verify` bind to on an 8-GPU box.  The *_dev entry points and set_stream need a
 * single-device context (device buffers live on one GPU) and return SCHNORR_B200_EARG here.  A device may be listed
 * more than once (independent contexts on the same GPU).  n_devices == 1 is schnorr_b200_create. */
int schnorr_b200_create_multi(const int *devices, int n_devices, schnorr_b200_ctx **out);
/* number of device contexts behind `ctx` (1 for schnorr_b200_create) */
int schnorr_b200_device_count(const schnorr_b200_ctx *ctx);
void schnorr_b200_destroy(schnorr_b200_ctx *ctx);
const char *schnorr_b200_last_error(const schnorr_b200_ctx *ctx);
/* Use an external stream (e.g. the caller's current stream) for all subsequent work. NULL = own stream. */
int schnorr_b200_set_stream(schnorr_b200_ctx *ctx, void *stream);
int schnorr_b200_synchronize(schnorr_b200_ctx *ctx);
/* Number of kernel launches issued by this context since creation (bench `gpu_launches`). */
uint64_t schnorr_b200_launch_count(const schnorr_b200_ctx *ctx);
/* Device time (CUDA events on the context stream) of the dominant kernel of the last *_dev call:
 * k_verify (verify_many), k_hash (hash_messages), k_msm_segment_sum (batch).  Synchronises on it. */
int schnorr_b200_last_kernel_ms(schnorr_b200_ctx *ctx, float *ms);
/* Single verification runs a fast path (inversion-free affine formulas, denominators kept in Fp) and re-runs the items
 * that hit one of its exceptional cases (identity / small-order keys, colliding partial sums) through the exact Jacobian kernel.
 * exact_only = 1 sends everything through the exact kernel (A/B measurements, tests).
 * last_exact_count: how many items of the last verify_many* call took the exact kernel (synchronises). */
int schnorr_b200_set_exact_only(schnorr_b200_ctx *ctx, int exact_only);
int schnorr_b200_last_exact_count(schnorr_b200_ctx *ctx, uint64_t *count);
/* Calls of at most `max_signatures` signatures (per pipeline chunk) run the warp-cooperative kernel (one signature per
 * six lanes: ~4x lower latency, 6x more parallelism per signature, ~1.5x the work); larger calls the one-signature-per-
 * thread kernel.  0 disables it, SIZE_MAX forces it (tests).  Default 10240 (measured crossover ~12 k). */
int schnorr_b200_set_dist_threshold(schnorr_b200_ctx *ctx, size_t max_signatures);
/* Calls of at most `max_signatures` signatures (per pipeline chunk) run one thread BLOCK per signature: the doubling
 * chain, the challenge hash, e*G and then the sixteen bucket accumulations of one verification proceed side by side on
 * the block's warps (lowest latency for the reference's own Criterion case, a single Signature::verify,
 * benches/schnorr.rs:22-66).  0 disables it.  Default 512 (measured: 0.50-0.65 ms for 1..512 signatures against
 * 0.98-1.08 ms on six lanes; beyond ~1000 blocks the six-lane kernel is faster). */
int schnorr_b200_set_one_threshold(schnorr_b200_ctx *ctx, size_t max_signatures);
/* Batches of at most `max_signatures` compute their challenges with one signature per six lanes ahead of the (then
 * hash-free) prepare kernel; larger ones hash one signature per thread.  Default 2^18 (measured crossover). */
int schnorr_b200_set_batch_dist_threshold(schnorr_b200_ctx *ctx, size_t max_signatures);
/* Batches (per device) of at most `max_signatures` signatures run one thread block per signature: block i evaluates its
 * own term s_i R_i - s_i h_i P_i of the batch equation (src/batch.rs:102-123) with the two doubling chains, the challenge
 * hash and the bucket accumulations side by side on its warps, and one more block adds the terms -- the latency form for
 * the reference's Criterion cases of 4 ... 128 signatures (benches/schnorr.rs:67-96).  Larger batches take the Pippenger
 * pipeline.  0 disables it.  Default 256. */
int schnorr_b200_set_batch_small_threshold(schnorr_b200_ctx *ctx, size_t max_signatures);
/* Test hook: force the Pippenger window width (4..16, 0 = planner's choice) and the segment length of the bucket
 * accumulation (>= 8, 0 = automatic) of the batch path, to exercise the skewed-bucket code paths. */
int schnorr_b200_set_msm_geometry(schnorr_b200_ctx *ctx, int window_bits, unsigned segment_len);
/* Pippenger geometry of the last batch call on this context (window width c, number of windows K, entries per
 * accumulation segment): the unit count of the batch roofline is computed from the plan actually used (cost_model.py). */
int schnorr_b200_last_batch_plan(const schnorr_b200_ctx *ctx, int *window_bits, int *windows, unsigned *segment_len);

/* hash_message(&Fp6, &PublicKey, &[u8]) -> [u8; 32]            src/signature.rs:274-306
 * rx48: n x 48 B (R.x limbs), pk96: n x 96 B, digests: n x 32 B.  */
int schnorr_b200_hash_messages(schnorr_b200_ctx *ctx, size_t n, const uint8_t *rx48, const uint8_t *pk96,
                               const uint8_t *msgs, const uint64_t *msg_off, uint8_t *digests);
int schnorr_b200_hash_messages_dev(schnorr_b200_ctx *ctx, size_t n, const uint8_t *rx48, const uint8_t *pk96,
                                   const uint8_t *msgs, const uint64_t *msg_off, uint8_t *digests);

/* Signature::verify(self, message, &PublicKey) for n independent triples  src/signature.rs:181-205
 * (also PublicKey::verify_signature :170-176, KeyPair::verify_signature :159-165,
 *  KeyedSignature::verify :232-234).  pk_inf may be NULL (no identity keys). */
int schnorr_b200_verify_many(schnorr_b200_ctx *ctx, size_t n, const uint8_t *sigs81, const uint8_t *pk96,
                             const uint8_t *pk_inf, const uint8_t *msgs, const uint64_t *msg_off,
                             uint8_t *verdicts);
int schnorr_b200_verify_many_dev(schnorr_b200_ctx *ctx, size_t n, const uint8_t *sigs81, const uint8_t *pk96,
                                 const uint8_t *pk_inf, const uint8_t *msgs, const uint64_t *msg_off,
                                 uint8_t *verdicts);

/* KeyedSignature::from_bytes + KeyedSignature::verify over n 130-byte wire records
 * (49-byte compressed public key || 81-byte signature)   src/signature.rs:232-271, src/constants.rs:33
 * The key is decompressed on the device; a record whose key does not decode gets verdict 3. */
int schnorr_b200_verify_keyed_many(schnorr_b200_ctx *ctx, size_t n, const uint8_t *keyed130, const uint8_t *msgs,
                                   const uint64_t *msg_off, uint8_t *verdicts);
int schnorr_b200_verify_keyed_many_dev(schnorr_b200_ctx *ctx, size_t n, const uint8_t *keyed130, const uint8_t *msgs,
                                       const uint64_t *msg_off, uint8_t *verdicts);

/* verify_batch(&[Signature], &[PublicKey], &[&[u8]], rng)                  src/batch.rs:31-50
 * = generate_batch_coefficients (:56-81) + verify_prepared_batch (:84-130).
 * rand32: n x 32 B caller-supplied randomisers s_i (reduced mod q) -- the seam the reference has at
 * verify_prepared_batch.  *verdict: 0 / 2 / 3.  lhs97 / rhs97 (optional): affine x||y||inf of
 * sum s_i R_i - sum s_i h_i P_i  and of (sum s_i e_i) G  for bit-exact comparison. */
int schnorr_b200_verify_batch(schnorr_b200_ctx *ctx, size_t n, const uint8_t *sigs81, const uint8_t *pk96,
                              const uint8_t *pk_inf, const uint8_t *msgs, const uint64_t *msg_off,
                              const uint8_t *rand32, int *verdict, uint8_t *lhs97, uint8_t *rhs97);
/* Same on DEVICE buffers (single-device context): enqueues the whole batch -- partial MSM, (sum s_i e_i) G on a second
 * stream beside it, finish -- and does not synchronise.  result216 (device): verdict(1) pad(7) lhs97 pad(7) rhs97 pad(7). */
int schnorr_b200_verify_batch_dev(schnorr_b200_ctx *ctx, size_t n, const uint8_t *sigs81, const uint8_t *pk96,
                                  const uint8_t *pk_inf, const uint8_t *msgs, const uint64_t *msg_off,
                                  const uint8_t *rand32, uint8_t *result216);
/* Multi-GPU form: each rank reduces its slice to one 192-byte partial (device buffer), the caller
 * gathers the partials (one small NCCL gather) and any rank finishes.  */
int schnorr_b200_batch_partial_dev(schnorr_b200_ctx *ctx, size_t n, const uint8_t *sigs81, const uint8_t *pk96,
                                   const uint8_t *pk_inf, const uint8_t *msgs, const uint64_t *msg_off,
                                   const uint8_t *rand32, uint8_t *partial192);
int schnorr_b200_batch_finish_dev(schnorr_b200_ctx *ctx, size_t n_partials, const uint8_t *partials192,
                                  uint8_t *result216 /* device, 216 B: verdict(1) pad(7) lhs97 pad(7) rhs97 pad(7) */);
int schnorr_b200_batch_finish(schnorr_b200_ctx *ctx, size_t n_partials, const uint8_t *partials192_host,
                              int *verdict, uint8_t *lhs97, uint8_t *rhs97);

/* Failed-batch localisation.  The reference answers a bad batch with ONE Err for the lot
 * (src/batch.rs:125-129); this names the culprits, in BATCH semantics: item i is bad when its own term of the batch
 * equation, R_i - h_i P_i - e_i G with R_i = from_compressed(sig_i.x) INCLUDING the flag byte and no subgroup check on
 * P_i (src/batch.rs:102-106), is not the identity -- e.g. a signature whose y-sign flag is flipped fails the batch although
 * Signature::verify (x-only, src/signature.rs:186,200) accepts it.  Bisection over partial MSMs (a range whose own random
 * linear combination holds is clean), exact per-item check on the failing slices.  flags[i] = 0 clean, 2 bad, 3 malformed
 * (the reference panics on it); *n_bad (optional) = number of non-zero flags.  rand32 as in verify_batch. */
int schnorr_b200_locate_invalid(schnorr_b200_ctx *ctx, size_t n, const uint8_t *sigs81, const uint8_t *pk96,
                                const uint8_t *pk_inf, const uint8_t *msgs, const uint64_t *msg_off,
                                const uint8_t *rand32, uint8_t *flags, uint64_t *n_bad);
int schnorr_b200_locate_invalid_dev(schnorr_b200_ctx *ctx, size_t n, const uint8_t *sigs81, const uint8_t *pk96,
                                    const uint8_t *pk_inf, const uint8_t *msgs, const uint64_t *msg_off,
                                    const uint8_t *rand32, uint8_t *flags);

/* PublicKey::from(&PrivateKey) = BASEPOINT_TABLE * sk                      src/public.rs:26-32 */
int schnorr_b200_keygen(schnorr_b200_ctx *ctx, size_t n, const uint8_t *sk32, uint8_t *pk96, uint8_t *pk_inf);
int schnorr_b200_keygen_dev(schnorr_b200_ctx *ctx, size_t n, const uint8_t *sk32, uint8_t *pk96, uint8_t *pk_inf);
/* KeyPair::sign(&self, message, rng) with the nonce r supplied by the caller  src/signature.rs:114-129 */
int schnorr_b200_sign_many(schnorr_b200_ctx *ctx, size_t n, const uint8_t *sk32, const uint8_t *pk96,
                           const uint8_t *pk_inf, const uint8_t *msgs, const uint64_t *msg_off,
                           const uint8_t *nonce32, uint8_t *sigs81);
int schnorr_b200_sign_many_dev(schnorr_b200_ctx *ctx, size_t n, const uint8_t *sk32, const uint8_t *pk96,
                               const uint8_t *pk_inf, const uint8_t *msgs, const uint64_t *msg_off,
                               const uint8_t *nonce32, uint8_t *sigs81);

/* Hierarchical deterministic key derivation (src/derivation.rs), n children per call, one child per thread:
 *   master keys      ExtendedPrivateKey::generate_master_key(seed)                   :66-84
 *                    seeds32: n x 32 B -> xsk64: n x (private key 32 B LE || chain code 32 B)
 *   private children ExtendedPrivateKey::derive_private(i) of ONE parent, hardened (i >= 2^31) or normal by index
 *                                                                                     :90-153
 *   public children  ExtendedPublicKey::derive_normal_public(i) of ONE parent        :249-277
 *                    xpk81 = PublicKey::to_bytes() 49 B || chain code 32 B (= ExtendedPublicKey::to_bytes, :280-286)
 * indices: n little-endian u32 (the reference's `&[u8; 4]`).  ok[i] = 1 where the reference returns Some: non-zero child
 * key / non-identity tweak point, and for public children a non-hardened index.  Records of children whose ok is 0 are
 * what the formulas give and must not be used.  HMAC-SHA512 and the fixed-base multiplication run on the device. */
int schnorr_b200_derive_master_keys(schnorr_b200_ctx *ctx, size_t n, const uint8_t *seeds32, uint8_t *xsk64, uint8_t *ok);
int schnorr_b200_derive_private_children(schnorr_b200_ctx *ctx, size_t n, const uint8_t *parent_xsk64,
                                         const uint32_t *indices, uint8_t *children_xsk64, uint8_t *ok);
int schnorr_b200_derive_public_children(schnorr_b200_ctx *ctx, size_t n, const uint8_t *parent_xpk81,
                                        const uint32_t *indices, uint8_t *children_xpk81, uint8_t *ok);

/* PublicKey::from_bytes / AffinePoint::from_compressed for n 49-byte records  src/public.rs:54-56
 * ok[i] = 1 when the record decodes (CtOption is_some). */
int schnorr_b200_decompress(schnorr_b200_ctx *ctx, size_t n, const uint8_t *in49, uint8_t *pk96, uint8_t *pk_inf,
                            uint8_t *ok);
/* AffinePoint::to_compressed / PublicKey::to_bytes                           src/public.rs:49-51 */
int schnorr_b200_compress(schnorr_b200_ctx *ctx, size_t n, const uint8_t *pk96, const uint8_t *pk_inf,
                          uint8_t *out49);

/* Test hook (no reference counterpart): raw device field operations so that the PTX arithmetic can be
 * checked against the oracle on edge values.  a6, b6: n x 6 canonical limbs (host); out48: n x 48 u64 =
 * fp6_mul(a,b) | fp6_sqr(a) | a+b | a-b | limb-wise a_k*b_k | limb-wise a_k^2 | fp6_inv(a) | limb-wise a_k^(1/7). */
int schnorr_b200_debug_field_ops(schnorr_b200_ctx *ctx, size_t n, const uint64_t *a6, const uint64_t *b6,
                                 uint64_t *out48);

/* Test hook (no reference counterpart): the lazily reduced building blocks of the fast path (csrc/debug_ops.cuh).
 * a6: n x 6 ARBITRARY 64-bit limbs (non-canonical representatives allowed), b6: n x 6 canonical limbs; out48: n x 48
 * u64 = a*rot(a) | a^2-2b | a^2-b-rot(b)*a3 | a*rot(a)-b*a5 | cofactor(b) | norm(b) | b*a0-rot(b)*a1 (nc form) |
 * b*a0-rot(b)*a1, all canonicalised. */
int schnorr_b200_debug_lazy_ops(schnorr_b200_ctx *ctx, size_t n, const uint64_t *a6, const uint64_t *b6,
                                uint64_t *out48);

/* Integer-multiply roofline calibration: runs a register-resident chain of `iters` dependent-free
 * 32x32->64 multiply-adds per thread on a full grid and returns wide multiplies per second. */
int schnorr_b200_imad_peak(schnorr_b200_ctx *ctx, int iters, double *wide_mul_per_s, double *elapsed_ms);

#ifdef __cplusplus
}
#endif
#endif /* SCHNORR_B200_H */
