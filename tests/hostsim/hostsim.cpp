// TEST-ONLY host build of the device headers (schnorr-sig_b200/csrc/*.cuh) with g++.
// The PTX blocks fall back to portable C (fp.cuh), everything else is the exact code the CUDA
// kernels instantiate.  It lets tests/ check the field/curve/hash FORMULAS against the oracle on a
// CPU-only box.  It is not part of the product: the product library has no CPU path, and nothing
// outside tests/ builds or loads this file.
#define SB_COUNT_WIDE 1
#include <cstdint>
#include <cstring>
#include <vector>

#include "../../include/cheetah_params.h"
#include "../../schnorr-sig_b200/csrc/verify.cuh"
#include "../../schnorr-sig_b200/csrc/debug_ops.cuh"
#include "../../schnorr-sig_b200/csrc/derive.cuh"

using namespace sb;

#define API extern "C" __attribute__((visibility("default")))

static fp6 ld6(const uint64_t* p) {
    fp6 r;
    memcpy(r.c, p, 48);
    return r;
}
static void st6(uint64_t* p, const fp6& a) { memcpy(p, a.c, 48); }
static scalar ldsc(const uint8_t* p) { return sc_load_le(p); }
static void stsc(uint8_t* p, const scalar& s) { memcpy(p, s.l, 32); }

API uint64_t hs_fp_mul(uint64_t a, uint64_t b) { return fp_mul(a, b); }
API uint64_t hs_fp_sqr(uint64_t a) { return fp_sqr(a); }
API uint64_t hs_fp_add(uint64_t a, uint64_t b) { return fp_add(a, b); }
API uint64_t hs_fp_sub(uint64_t a, uint64_t b) { return fp_sub(a, b); }
API uint64_t hs_fp_inv(uint64_t a) { return fp_inv(a); }
API uint64_t hs_fp_mul_small(uint64_t a, uint32_t k) { return fp_mul_small(a, k); }
API int hs_fp_sqrt(uint64_t a, uint64_t* r) { fp_t o = 0; bool ok = fp_sqrt(a, o); *r = o; return ok; }
API uint64_t hs_fp_reduce160(uint32_t x0, uint32_t x1, uint32_t x2, uint32_t x3, uint32_t x4) {
    return fp_reduce160(x0, x1, x2, x3, x4);
}
API void hs_fp6_mul(const uint64_t* a, const uint64_t* b, uint64_t* r) { st6(r, fp6_mul(ld6(a), ld6(b))); }
API void hs_fp6_sqr(const uint64_t* a, uint64_t* r) { st6(r, fp6_sqr(ld6(a))); }
API void hs_fp6_inv(const uint64_t* a, uint64_t* r) { st6(r, fp6_inv(ld6(a))); }
API int hs_fp6_sqrt(const uint64_t* a, uint64_t* r) {
    fp6 o = fp6_zero();
    bool ok = fp6_sqrt(ld6(a), o);
    st6(r, o);
    return ok;
}
API int hs_fp6_lex_largest(const uint64_t* a) { return fp6_lex_largest(ld6(a)); }
API void hs_rescue_permutation(uint64_t* s) { rescue_permutation(s); }
API uint64_t hs_rescue_inv_sbox(uint64_t x) { return rescue_inv_sbox(x); }
API void hs_hash_message(const uint64_t* rx, const uint64_t* px, uint64_t py0, const uint8_t* msg, uint64_t len,
                         uint8_t* out32) {
    fp_t d[4];
    hash_message(ld6(rx), ld6(px), py0, msg, len, d);
    memcpy(out32, d, 32);
}
API void hs_sc_mul(const uint8_t* a, const uint8_t* b, uint8_t* r) { stsc(r, sc_mul(ldsc(a), ldsc(b))); }
API void hs_sc_add(const uint8_t* a, const uint8_t* b, uint8_t* r) { stsc(r, sc_add(ldsc(a), ldsc(b))); }
API void hs_sc_sub(const uint8_t* a, const uint8_t* b, uint8_t* r) { stsc(r, sc_sub(ldsc(a), ldsc(b))); }
API void hs_sc_from_u256(const uint8_t* a, uint8_t* r) { stsc(r, sc_from_u256(ldsc(a))); }
API int hs_sc_geq_q(const uint8_t* a) { return sc_geq_q(ldsc(a)); }
static std::vector<uint64_t> g_gtab;
static void build_gtab() {
    if (!g_gtab.empty()) return;
    g_gtab.assign((size_t)GTAB_WINDOWS * GTAB_ENTRIES * 12, 0);
    fp6 gx = ld6(CHEETAH_GX), gy = ld6(CHEETAH_GY);
    jac_pt base = jac_from_affine(gx, gy, false);
    for (int i = 0; i < GTAB_WINDOWS; i++) {
        jac_pt acc = jac_identity();
        for (int b = 1; b < GTAB_ENTRIES; b++) {
            acc = jac_add(acc, base);
            fp6 x, y;
            bool inf;
            jac_to_affine(acc, x, y, inf);
            uint64_t* o = &g_gtab[((size_t)i * GTAB_ENTRIES + b) * 12];
            memcpy(o, x.c, 48);
            memcpy(o + 6, y.c, 48);
        }
        for (int k = 0; k < GTAB_W; k++) base = jac_dbl(base);
    }
}
API const uint64_t* hs_gtab() {
    build_gtab();
    return g_gtab.data();
}

static void store_aff(const jac_pt& p, uint8_t* out96, int* out_inf) {
    fp6 x, y;
    bool inf;
    jac_to_affine(p, x, y, inf);
    memcpy(out96, x.c, 48);
    memcpy(out96 + 48, y.c, 48);
    *out_inf = inf;
}
static jac_pt load_pt(const uint8_t* p96, int inf) {
    fp6 x, y;
    memcpy(x.c, p96, 48);
    memcpy(y.c, p96 + 48, 48);
    return jac_from_affine(x, y, inf != 0);
}
API void hs_pt_add(const uint8_t* a96, int ainf, const uint8_t* b96, int binf, uint8_t* out96, int* out_inf) {
    store_aff(jac_add(load_pt(a96, ainf), load_pt(b96, binf)), out96, out_inf);
}
API void hs_pt_madd(const uint8_t* a96, int ainf, const uint8_t* b96, int binf, uint8_t* out96, int* out_inf) {
    fp6 x, y;
    memcpy(x.c, b96, 48);
    memcpy(y.c, b96 + 48, 48);
    // run the mixed addition from a non-trivial Z by first doubling-and-halving is not possible;
    // instead feed a Jacobian form with Z != 1: (X,Y,Z) = (x z^2, y z^3, z) with z = 3
    jac_pt p = load_pt(a96, ainf);
    if (!ainf) {
        fp6 z = fp6{{3, 1, 0, 0, 0, 0}};
        fp6 z2 = fp6_sqr(z);
        p.X = fp6_mul(p.X, z2);
        p.Y = fp6_mul(p.Y, fp6_mul(z2, z));
        p.Z = z;
    }
    store_aff(jac_madd(p, x, y, binf != 0), out96, out_inf);
}
API void hs_pt_dbl(const uint8_t* a96, int ainf, uint8_t* out96, int* out_inf) {
    store_aff(jac_dbl(load_pt(a96, ainf)), out96, out_inf);
}
API void hs_fixed_base_mul(const uint8_t* k32, uint8_t* out96, int* out_inf) {
    build_gtab();
    store_aff(fixed_base_mul(sc_from_u256(ldsc(k32)), g_gtab.data()), out96, out_inf);
}
API int hs_torsion_free(const uint8_t* p96, int inf) {
    jac_pt r;
    jac_pt d;
    return torsion_check_and_mul(load_pt(p96, inf), sc_zero(), &r, &d);
}
// h*P + e*G exactly as verify_points composes it
API void hs_double_base(const uint8_t* p96, int inf, const uint8_t* h32, const uint8_t* e32, uint8_t* out96, int* out_inf) {
    build_gtab();
    jac_pt r, d;
    torsion_check_and_mul(load_pt(p96, inf), ldsc(h32), &r, &d);
    fixed_base_accumulate(&r, ldsc(e32), g_gtab.data());
    store_aff(r, out96, out_inf);
}
// shared-doubling core: returns the subgroup verdict and h*P
API int hs_torsion_check_and_mul(const uint8_t* p96, int inf, const uint8_t* h32, uint8_t* out96, int* out_inf) {
    jac_pt r, d;
    bool tf = torsion_check_and_mul(load_pt(p96, inf), ldsc(h32), &r, &d);
    store_aff(r, out96, out_inf);
    return tf;
}
API void hs_recode_signed_w4(const uint8_t* k, int8_t* digits64) { recode_signed_w4(ldsc(k), digits64); }
API int hs_decompress(const uint8_t* in49, uint8_t* out96, int* out_inf) {
    fp6 x;
    memcpy(x.c, in49, 48);
    fp6 ox, oy;
    bool inf;
    bool ok = decompress_point(x, in49[48], ox, oy, inf);
    memcpy(out96, ox.c, 48);
    memcpy(out96 + 48, oy.c, 48);
    *out_inf = inf;
    return ok;
}
// full per-signature path exactly as k_ingest + k_verify compose it
API int hs_verify_one(const uint8_t* sig81, const uint8_t* pk96, int pk_inf, const uint8_t* msg, uint64_t len) {
    build_gtab();
    fp6 sx, px, py;
    memcpy(sx.c, sig81, 48);
    memcpy(px.c, pk96, 48);
    memcpy(py.c, pk96 + 48, 48);
    scalar e = ldsc(sig81 + 49);
    bool x_ok = fp6_is_canonical(sx);
    bool pk_ok = fp6_is_canonical(px) && fp6_is_canonical(py);
    if ((!pk_ok && !pk_inf) || sc_geq_q(e)) return VERDICT_MALFORMED;
    scalar h = sc_zero();
    if (x_ok) h = challenge_scalar(sx, px, py, pk_inf != 0, msg, len);
    jac_pt d;
    return verify_points(sx, x_ok, e, px, py, pk_inf != 0, h, g_gtab.data(), &d);
}

// ---- fast path in (X, Y, w) coordinates (affine.cuh) ----
// returns fast_result; out96 = h*P + e*G when FAST_TORSION_FREE / FAST_NOT_TORSION_FREE
API int hs_verify_core_fast(const uint8_t* p96, const uint8_t* h32, const uint8_t* e32, uint8_t* out96) {
    build_gtab();
    fp6 x, y;
    memcpy(x.c, p96, 48);
    memcpy(y.c, p96 + 48, 48);
    jf_pt r, d, bh[FAST_BH_SHARED];
    memset(&r, 0, sizeof r);
    r.w = 1;
    int fr = verify_core_fast(x, y, ldsc(h32), ldsc(e32), g_gtab.data(), &r, &d, bh, 1);
    // normalise (X, Y, w) for the comparison with the oracle
    fp_t wi = fp_inv(fp_canon(r.w)), wi2 = fp_sqr(wi);
    fp6 ax = fp6_scale(r.X, wi2), ay = fp6_scale(r.Y, fp_mul(wi2, wi));
    memcpy(out96, ax.c, 48);
    memcpy(out96 + 48, ay.c, 48);
    return fr;
}
// full per-signature path as k_ingest + k_verify_fast (+ the exact kernel for flagged items) compose it;
// *used_exact reports whether the fallback ran
API int hs_verify_one_fast(const uint8_t* sig81, const uint8_t* pk96, int pk_inf, const uint8_t* msg, uint64_t len, int* used_exact) {
    build_gtab();
    fp6 sx, px, py;
    memcpy(sx.c, sig81, 48);
    memcpy(px.c, pk96, 48);
    memcpy(py.c, pk96 + 48, 48);
    scalar e = ldsc(sig81 + 49);
    bool x_ok = fp6_is_canonical(sx);
    bool pk_ok = fp6_is_canonical(px) && fp6_is_canonical(py);
    *used_exact = 0;
    if ((!pk_ok && !pk_inf) || sc_geq_q(e)) return VERDICT_MALFORMED;
    scalar h = sc_zero();
    if (x_ok) h = challenge_scalar(sx, px, py, pk_inf != 0, msg, len);
    jf_pt da, bh[FAST_BH_SHARED];
    uint8_t v = verify_points_fast(sx, x_ok, e, px, py, pk_inf != 0, h, g_gtab.data(), &da, bh, 1);
    if (v != VERDICT_NEEDS_EXACT) return v;
    *used_exact = 1;
    jac_pt d;
    return verify_points(sx, x_ok, e, px, py, pk_inf != 0, h, g_gtab.data(), &d);
}

// one (X, Y, w) operation on affine inputs given a scrambling denominator: in/out as affine 96-byte points
static jf_pt jf_from_affine(const uint8_t* p96, uint64_t w) {
    jf_pt r;
    fp6 x, y;
    memcpy(x.c, p96, 48);
    memcpy(y.c, p96 + 48, 48);
    fp_t w2 = fp_sqr(w);
    r.X = fp6_scale(x, w2);
    r.Y = fp6_scale(y, fp_mul(w2, w));
    r.w = w;
    return r;
}
static void jf_to_affine(const jf_pt& r, uint8_t* out96) {
    fp_t wi = fp_inv(fp_canon(r.w)), wi2 = fp_sqr(wi);
    fp6 ax = fp6_scale(r.X, wi2), ay = fp6_scale(r.Y, fp_mul(wi2, wi));
    memcpy(out96, ax.c, 48);
    memcpy(out96 + 48, ay.c, 48);
}
API int hs_jf_add(const uint8_t* a96, uint64_t wa, const uint8_t* b96, uint64_t wb, int mode, uint8_t* out96) {
    jf_pt a = jf_from_affine(a96, wa), b = jf_from_affine(b96, wb);
    bool exc = jf_add(&a, &b, (uint8_t)mode);
    jf_to_affine(a, out96);
    return exc;
}
API int hs_jf_dbl(const uint8_t* a96, uint64_t wa, uint8_t* out96) {
    jf_pt a = jf_from_affine(a96, wa);
    bool exc = jf_dbl(&a);
    if (!exc) jf_to_affine(a, out96);
    return exc;
}
// the fused forms k_verify_fast instantiates (lazy accumulations, non-canonical intermediates)
API int hs_jf_add_fused(const uint8_t* a96, uint64_t wa, const uint8_t* b96, uint64_t wb, int mode, uint8_t* out96) {
    jf_pt a = jf_from_affine(a96, wa), b = jf_from_affine(b96, wb);
    bool exc = jf_add<true>(&a, &b, (uint8_t)mode);
    jf_to_affine(a, out96);
    return exc;
}
API int hs_jf_dbl_fused(const uint8_t* a96, uint64_t wa, uint8_t* out96) {
    jf_pt a = jf_from_affine(a96, wa);
    bool exc = jf_dbl<true>(&a);
    if (!exc) jf_to_affine(a, out96);
    return exc;
}
API uint64_t hs_fp6_cofactor_norm(const uint64_t* d, uint64_t* c) {
    fp6 dd = ld6(d), cc;
    fp_t n;
    fp6_cofactor_norm(&dd, &cc, &n);
    st6(c, cc);
    return n;
}

// lazily reduced building blocks on raw limbs (a may be non-canonical): out = 8 x 6 limbs (debug_ops.cuh)
API void hs_debug_lazy_ops(const uint64_t* a, const uint64_t* b, uint64_t* out48) {
    fp6 r[8];
    debug_lazy_ops(ld6(a), ld6(b), r);
    for (int j = 0; j < 8; j++) st6(out48 + 6 * j, r[j]);
}

// 32x32->64 multiplies executed since the last call (cost figures of DESIGN.md; test-only counter in fp.cuh)
API unsigned long long hs_wide_count_reset() {
    unsigned long long v = g_wide_count;
    g_wide_count = 0;
    return v;
}
// Pippenger bucket accumulation step: acc (affine, or identity when a_inf) scrambled by wa, += (+|-) b; returns the
// affine result (out_inf = identity)
API void hs_jf_madd_exact(const uint8_t* a96, int a_inf, uint64_t wa, const uint8_t* b96, int neg, uint8_t* out96, int* out_inf) {
    jf_pt a = jf_from_affine(a96, wa), t = jf_from_affine(b96, 1);
    if (a_inf) a.w = 0;
    jf_madd_exact(&a, &t, neg != 0);
    *out_inf = a.w == 0;
    if (a.w != 0) jf_to_affine(a, out96);
}
API void hs_jf_add_exact(const uint8_t* a96, int a_inf, uint64_t wa, const uint8_t* b96, int b_inf, uint64_t wb, uint8_t* out96, int* out_inf) {
    jf_pt a = jf_from_affine(a96, wa), b = jf_from_affine(b96, wb);
    if (a_inf) a.w = 0;
    if (b_inf) b.w = 0;
    jf_add_exact(&a, &b);
    *out_inf = a.w == 0;
    if (a.w != 0) jf_to_affine(a, out96);
}

// ---- HD derivation building blocks (derive.cuh) ----
API void hs_hmac_sha512(const uint8_t* key, int key_len, const uint8_t* data, int data_len, uint8_t* out64) {
    hmac_sha512_short(key, key_len, data, data_len, out64);
}
API void hs_sha512_short(const uint8_t* data, int len, uint8_t* out64) {
    sha512_state s;
    sha512_init(s);
    sha512_finish_short(s, 0, data, len, out64);
}
API int hs_derive_master(const uint8_t* seed32, uint8_t* xsk64) { return derive_master(seed32, xsk64); }
API int hs_derive_private_child(const uint8_t* xsk64, const uint8_t* pk49, uint32_t index, uint8_t* child64) {
    return derive_private_child(xsk64, pk49, index, child64);
}
