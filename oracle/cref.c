/* Oracle #2 -- plain-C CPU restatement of the schnorr-sig verification path.
 *
 * TEST INFRASTRUCTURE ONLY: built into oracle/libcref.so, loaded by tests/, by
 * __graft_entry__.smoke() and by bench.py's cpu_baseline / --impl reference legs as the
 * checker and the timed CPU baseline.  The product library (schnorr-sig_b200/) never links,
 * loads or calls it.
 *
 * PARITY STATUS: "parity unpinned" at value level -- see oracle/pyref.py header.  This file is
 * itself pinned against oracle/pyref.py (tests/test_oracle_*.py) and against every fact the
 * reference's own tests hold (KAT point, encodings, accept/reject behaviour).
 *
 * Protocol glue follows the reference:
 *   hash_message        /root/reference/src/signature.rs:274-306
 *   sign                /root/reference/src/signature.rs:114-129
 *   Signature::verify   /root/reference/src/signature.rs:181-205
 *   verify_batch        /root/reference/src/batch.rs:31-130
 * Arithmetic (the `cheetah` and `hash` crates are not on this box) follows the published
 * definitions; algorithms are deliberately the textbook ones (bitwise double-and-add, generic
 * Tonelli-Shanks) so that this checker shares no algorithmic choice with the CUDA kernels.
 */
#include <pthread.h>
#include <stdint.h>
#include <stdlib.h>
#include <string.h>

#include "../include/cheetah_params.h"

typedef unsigned __int128 u128;
typedef uint64_t u64;
typedef uint8_t u8;

#define PP CHEETAH_P
#define API __attribute__((visibility("default")))

/* ------------------------------------------------------------------ Fp (Goldilocks) */
static inline u64 fp_add(u64 a, u64 b) {
    u64 s = a + b;
    int c = s < a;
    if (c || s >= PP) s -= PP;
    return s;
}
static inline u64 fp_sub(u64 a, u64 b) { return a >= b ? a - b : a + (PP - b); }
static inline u64 fp_neg(u64 a) { return a ? PP - a : 0; }
/* x mod p for x < 2^128 using 2^64 = 2^32-1, 2^96 = -1 (mod p) */
static inline u64 fp_red(u128 x) {
    u64 lo = (u64)x, hi = (u64)(x >> 64);
    u64 hh = hi >> 32, hl = hi & 0xffffffffULL;
    u64 t = lo - hh;
    if (lo < hh) t -= 0xffffffffULL; /* borrowed 2^64 = 2^32-1 too much */
    u64 m = hl * 0xffffffffULL;
    u64 r = t + m;
    if (r < m) r += 0xffffffffULL;
    if (r >= PP) r -= PP;
    return r;
}
static inline u64 fp_mul(u64 a, u64 b) { return fp_red((u128)a * b); }
static u64 fp_pow(u64 a, u64 e) {
    u64 r = 1;
    while (e) {
        if (e & 1) r = fp_mul(r, a);
        a = fp_mul(a, a);
        e >>= 1;
    }
    return r;
}
static inline u64 fp_inv(u64 a) { return fp_pow(a, PP - 2); }

/* ------------------------------------------------------------------ Fp6 = Fp[u]/(u^6-7) */
typedef struct { u64 c[6]; } fp6;
static const fp6 F6_ZERO = {{0, 0, 0, 0, 0, 0}};
static const fp6 F6_ONE = {{1, 0, 0, 0, 0, 0}};

static inline fp6 f6_add(fp6 a, fp6 b) { fp6 r; for (int i = 0; i < 6; i++) r.c[i] = fp_add(a.c[i], b.c[i]); return r; }
static inline fp6 f6_sub(fp6 a, fp6 b) { fp6 r; for (int i = 0; i < 6; i++) r.c[i] = fp_sub(a.c[i], b.c[i]); return r; }
static inline fp6 f6_neg(fp6 a) { fp6 r; for (int i = 0; i < 6; i++) r.c[i] = fp_neg(a.c[i]); return r; }
static inline int f6_eq(fp6 a, fp6 b) { return memcmp(a.c, b.c, 48) == 0; }
static inline int f6_is_zero(fp6 a) { return f6_eq(a, F6_ZERO); }
static inline fp6 f6_dbl(fp6 a) { return f6_add(a, a); }

static fp6 f6_mul(fp6 a, fp6 b) {
    fp6 r;
    for (int k = 0; k < 6; k++) {
        u128 lo = 0, hi = 0;
        for (int i = 0; i < 6; i++) {
            int j = k - i;
            if (j >= 0) lo += fp_mul(a.c[i], b.c[j]);
            else hi += fp_mul(a.c[i], b.c[j + 6]);
        }
        r.c[k] = fp_red(lo + 7 * hi);
    }
    return r;
}
static inline fp6 f6_sqr(fp6 a) { return f6_mul(a, a); }

/* exponent given as little-endian 64-bit limbs */
static fp6 f6_pow(fp6 a, const u64 *e, int nl) {
    fp6 r = F6_ONE;
    for (int i = nl * 64 - 1; i >= 0; i--) {
        r = f6_sqr(r);
        if ((e[i >> 6] >> (i & 63)) & 1) r = f6_mul(r, a);
    }
    return r;
}

/* Fp3 = Fp[v]/(v^3-7) helpers for the tower inversion (v = u^2) */
typedef struct { u64 c[3]; } fp3;
static fp3 f3_mul(fp3 a, fp3 b) {
    fp3 r;
    u128 t;
    t = (u128)fp_mul(a.c[1], b.c[2]) + fp_mul(a.c[2], b.c[1]);
    r.c[0] = fp_red((u128)fp_mul(a.c[0], b.c[0]) + 7 * t);
    t = (u128)fp_mul(a.c[0], b.c[1]) + fp_mul(a.c[1], b.c[0]);
    r.c[1] = fp_red(t + 7 * (u128)fp_mul(a.c[2], b.c[2]));
    t = (u128)fp_mul(a.c[0], b.c[2]) + fp_mul(a.c[1], b.c[1]);
    r.c[2] = fp_red(t + fp_mul(a.c[2], b.c[0]));
    return r;
}
static fp3 f3_inv(fp3 d) {
    u64 d0 = d.c[0], d1 = d.c[1], d2 = d.c[2];
    u64 t0 = fp_sub(fp_mul(d0, d0), fp_mul(7, fp_mul(d1, d2)));
    u64 t1 = fp_sub(fp_mul(7, fp_mul(d2, d2)), fp_mul(d0, d1));
    u64 t2 = fp_sub(fp_mul(d1, d1), fp_mul(d0, d2));
    u64 n = fp_add(fp_mul(d0, t0), fp_mul(7, fp_add(fp_mul(d2, t1), fp_mul(d1, t2))));
    u64 ni = fp_inv(n);
    fp3 r = {{fp_mul(t0, ni), fp_mul(t1, ni), fp_mul(t2, ni)}};
    return r;
}
static fp6 f6_inv(fp6 a) {
    fp3 a0 = {{a.c[0], a.c[2], a.c[4]}}, a1 = {{a.c[1], a.c[3], a.c[5]}};
    fp3 s0 = f3_mul(a0, a0), s1 = f3_mul(a1, a1);
    fp3 vs1 = {{fp_mul(7, s1.c[2]), s1.c[0], s1.c[1]}};
    fp3 d = {{fp_sub(s0.c[0], vs1.c[0]), fp_sub(s0.c[1], vs1.c[1]), fp_sub(s0.c[2], vs1.c[2])}};
    fp3 di = f3_inv(d);
    fp3 r0 = f3_mul(a0, di);
    fp3 na1 = {{fp_neg(a1.c[0]), fp_neg(a1.c[1]), fp_neg(a1.c[2])}};
    fp3 r1 = f3_mul(na1, di);
    fp6 r = {{r0.c[0], r1.c[0], r0.c[1], r1.c[1], r0.c[2], r1.c[2]}};
    return r;
}

/* generic Tonelli-Shanks in Fp6: p^6-1 = 2^33 * T.  T, (T+1)/2 and z^T are computed once. */
static u64 TS_T[6], TS_T1H[6];
static fp6 TS_Z;
static int ts_ready = 0;
static pthread_mutex_t ts_lock = PTHREAD_MUTEX_INITIALIZER;

static void big_mul_small(u64 *a, int n, u64 m) { /* a *= m */
    u128 c = 0;
    for (int i = 0; i < n; i++) { c += (u128)a[i] * m; a[i] = (u64)c; c >>= 64; }
}
static void big_sub1(u64 *a, int n) { for (int i = 0; i < n; i++) { if (a[i]--) break; } }
static void big_add1(u64 *a, int n) { for (int i = 0; i < n; i++) { if (++a[i]) break; } }
static void big_shr(u64 *a, int n, int s) {
    for (int i = 0; i < n; i++) a[i] = (a[i] >> s) | (i + 1 < n ? a[i + 1] << (64 - s) : 0);
}
static int f6_is_square(fp6 a) {
    /* a^((p^6-1)/2) = (a^T)^(2^32) */
    fp6 t = f6_pow(a, TS_T, 6);
    for (int i = 0; i < 32; i++) t = f6_sqr(t);
    return f6_eq(t, F6_ONE);
}
static void ts_setup(void) {
    pthread_mutex_lock(&ts_lock);
    if (!ts_ready) {
        u64 n[6] = {1, 0, 0, 0, 0, 0};
        for (int i = 0; i < 6; i++) big_mul_small(n, 6, PP); /* p^6 */
        big_sub1(n, 6);
        big_shr(n, 6, 33);
        memcpy(TS_T, n, 48);
        big_add1(n, 6);
        big_shr(n, 6, 1);
        memcpy(TS_T1H, n, 48);
        for (u64 c = 0;; c++) { /* first non-square of the form c + u */
            fp6 z = {{c, 1, 0, 0, 0, 0}};
            if (!f6_is_square(z)) { TS_Z = f6_pow(z, TS_T, 6); break; }
        }
        ts_ready = 1;
    }
    pthread_mutex_unlock(&ts_lock);
}
/* returns 1 and a root in *out, or 0 */
static int f6_sqrt(fp6 a, fp6 *out) {
    if (f6_is_zero(a)) { *out = F6_ZERO; return 1; }
    if (!ts_ready) ts_setup();
    fp6 x = f6_pow(a, TS_T1H, 6), b = f6_pow(a, TS_T, 6), g = TS_Z;
    int r = 33;
    while (!f6_eq(b, F6_ONE)) {
        int m = 0;
        fp6 bb = b;
        while (!f6_eq(bb, F6_ONE)) { bb = f6_sqr(bb); if (++m == r) return 0; }
        fp6 gs = g;
        for (int i = 0; i < r - m - 1; i++) gs = f6_sqr(gs);
        g = f6_sqr(gs);
        x = f6_mul(x, gs);
        b = f6_mul(b, g);
        r = m;
    }
    *out = x;
    return 1;
}
static int f6_lex_largest(fp6 a) {
    for (int i = 5; i >= 0; i--) if (a.c[i]) return a.c[i] > (PP - 1) / 2;
    return 0;
}
static int f6_from_bytes(const u8 *b, fp6 *out) {
    for (int i = 0; i < 6; i++) {
        u64 v; memcpy(&v, b + 8 * i, 8);
        if (v >= PP) return 0;
        out->c[i] = v;
    }
    return 1;
}
static void f6_to_bytes(fp6 a, u8 *b) { memcpy(b, a.c, 48); }

/* ------------------------------------------------------------------ curve, Jacobian, a = 1 */
typedef struct { fp6 x, y; int inf; } aff;
typedef struct { fp6 X, Y, Z; } jac;
static const fp6 CURVE_B = {{CHEETAH_CURVE_B0, CHEETAH_CURVE_B1, 0, 0, 0, 0}};

static jac jac_inf(void) { jac r = {F6_ONE, F6_ONE, F6_ZERO}; return r; }
static jac to_jac(aff p) { if (p.inf) return jac_inf(); jac r = {p.x, p.y, F6_ONE}; return r; }
static aff to_aff(jac j) {
    aff r;
    if (f6_is_zero(j.Z)) { r.x = F6_ZERO; r.y = F6_ZERO; r.inf = 1; return r; }
    fp6 zi = f6_inv(j.Z), zi2 = f6_sqr(zi);
    r.x = f6_mul(j.X, zi2);
    r.y = f6_mul(j.Y, f6_mul(zi2, zi));
    r.inf = 0;
    return r;
}
static jac jac_dbl(jac p) {
    if (f6_is_zero(p.Z) || f6_is_zero(p.Y)) return jac_inf();
    fp6 YY = f6_sqr(p.Y);
    fp6 S = f6_dbl(f6_dbl(f6_mul(p.X, YY)));
    fp6 ZZ = f6_sqr(p.Z);
    fp6 XX = f6_sqr(p.X);
    fp6 M = f6_add(f6_add(f6_dbl(XX), XX), f6_sqr(ZZ));
    jac r;
    r.X = f6_sub(f6_sqr(M), f6_dbl(S));
    fp6 Y4 = f6_sqr(YY);
    r.Y = f6_sub(f6_mul(M, f6_sub(S, r.X)), f6_dbl(f6_dbl(f6_dbl(Y4))));
    r.Z = f6_dbl(f6_mul(p.Y, p.Z));
    return r;
}
static jac jac_add(jac p, jac q) {
    if (f6_is_zero(p.Z)) return q;
    if (f6_is_zero(q.Z)) return p;
    fp6 Z1Z1 = f6_sqr(p.Z), Z2Z2 = f6_sqr(q.Z);
    fp6 U1 = f6_mul(p.X, Z2Z2), U2 = f6_mul(q.X, Z1Z1);
    fp6 S1 = f6_mul(p.Y, f6_mul(q.Z, Z2Z2)), S2 = f6_mul(q.Y, f6_mul(p.Z, Z1Z1));
    if (f6_eq(U1, U2)) return f6_eq(S1, S2) ? jac_dbl(p) : jac_inf();
    fp6 H = f6_sub(U2, U1), R = f6_sub(S2, S1);
    fp6 HH = f6_sqr(H), HHH = f6_mul(H, HH), V = f6_mul(U1, HH);
    jac r;
    r.X = f6_sub(f6_sub(f6_sqr(R), HHH), f6_dbl(V));
    r.Y = f6_sub(f6_mul(R, f6_sub(V, r.X)), f6_mul(S1, HHH));
    r.Z = f6_mul(f6_mul(p.Z, q.Z), H);
    return r;
}
/* k: 32 little-endian bytes, any 256-bit value */
static jac jac_mul(aff p, const u8 *k) {
    jac acc = jac_inf(), base = to_jac(p);
    for (int i = 255; i >= 0; i--) {
        acc = jac_dbl(acc);
        if ((k[i >> 3] >> (i & 7)) & 1) acc = jac_add(acc, base);
    }
    return acc;
}
/* k1*p1 + k2*p2, bitwise Shamir */
static jac jac_mul2(aff p1, const u8 *k1, aff p2, const u8 *k2) {
    jac acc = jac_inf(), j1 = to_jac(p1), j2 = to_jac(p2), j12 = jac_add(j1, j2);
    for (int i = 255; i >= 0; i--) {
        acc = jac_dbl(acc);
        int b1 = (k1[i >> 3] >> (i & 7)) & 1, b2 = (k2[i >> 3] >> (i & 7)) & 1;
        if (b1 && b2) acc = jac_add(acc, j12);
        else if (b1) acc = jac_add(acc, j1);
        else if (b2) acc = jac_add(acc, j2);
    }
    return acc;
}
static aff generator(void) {
    aff g; memcpy(g.x.c, CHEETAH_GX, 48); memcpy(g.y.c, CHEETAH_GY, 48); g.inf = 0; return g;
}
static int is_torsion_free(aff p) { /* [q]P == O, call site src/signature.rs:182 */
    return f6_is_zero(jac_mul(p, (const u8 *)CHEETAH_Q64).Z);
}
static void compress(aff p, u8 *out49) {
    if (p.inf) { memset(out49, 0, 48); out49[48] = 0x80; return; }
    f6_to_bytes(p.x, out49);
    out49[48] = f6_lex_largest(p.y) ? 0x40 : 0x00;
}
static int decompress(const u8 *in49, aff *out) {
    u8 flags = in49[48];
    out->x = F6_ZERO; out->y = F6_ZERO; out->inf = 1;
    if (flags & 0x3f) return 0;
    int inf = flags >> 7, sign = (flags >> 6) & 1;
    fp6 x;
    if (!f6_from_bytes(in49, &x)) return 0;
    if (inf) return f6_is_zero(x) && !sign;
    fp6 rhs = f6_add(f6_add(f6_mul(f6_sqr(x), x), x), CURVE_B), y;
    if (!f6_sqrt(rhs, &y)) return 0;
    if (f6_lex_largest(y) != sign) y = f6_neg(y);
    out->x = x; out->y = y; out->inf = 0;
    return 1;
}

/* ------------------------------------------------------------------ scalars mod q (4 x u64) */
typedef struct { u64 l[4]; } sc;
static int sc_geq_q(const u64 *a) {
    for (int i = 3; i >= 0; i--) { if (a[i] != CHEETAH_Q64[i]) return a[i] > CHEETAH_Q64[i]; }
    return 1;
}
static void sc_sub_q(u64 *a) {
    u128 b = 0;
    for (int i = 0; i < 4; i++) { u128 d = (u128)a[i] - CHEETAH_Q64[i] - b; a[i] = (u64)d; b = (d >> 64) & 1; }
}
/* reduce an n-limb little-endian integer mod q by bitwise long division (obviously correct) */
static sc sc_reduce(const u64 *x, int n) {
    u64 r[5] = {0, 0, 0, 0, 0};
    for (int i = n * 64 - 1; i >= 0; i--) {
        r[4] = r[3] >> 63; r[3] = (r[3] << 1) | (r[2] >> 63); r[2] = (r[2] << 1) | (r[1] >> 63);
        r[1] = (r[1] << 1) | (r[0] >> 63); r[0] = (r[0] << 1) | ((x[i >> 6] >> (i & 63)) & 1);
        if (r[4] || sc_geq_q(r)) sc_sub_q(r); /* r < 2q < 2^256 so r[4] is always 0; kept for clarity */
    }
    sc s; memcpy(s.l, r, 32); return s;
}
static sc sc_mul(sc a, sc b) {
    u64 t[8] = {0};
    for (int i = 0; i < 4; i++) {
        u128 c = 0;
        for (int j = 0; j < 4; j++) { c += (u128)a.l[i] * b.l[j] + t[i + j]; t[i + j] = (u64)c; c >>= 64; }
        t[i + 4] = (u64)c;
    }
    return sc_reduce(t, 8);
}
static sc sc_add(sc a, sc b) {
    u64 t[5]; u128 c = 0;
    for (int i = 0; i < 4; i++) { c += (u128)a.l[i] + b.l[i]; t[i] = (u64)c; c >>= 64; }
    t[4] = (u64)c;
    return sc_reduce(t, 5);
}
static sc sc_neg(sc a) {
    sc r; u128 b = 0; int z = !(a.l[0] | a.l[1] | a.l[2] | a.l[3]);
    if (z) return a;
    for (int i = 0; i < 4; i++) { u128 d = (u128)CHEETAH_Q64[i] - a.l[i] - b; r.l[i] = (u64)d; b = (d >> 64) & 1; }
    return r;
}
static sc sc_from_bytes_reduce(const u8 *b32) { u64 t[4]; memcpy(t, b32, 32); return sc_reduce(t, 4); }

/* ------------------------------------------------------------------ Rescue-Prime 64/12/8 */
static void rescue_permutation(u64 *s) {
    u64 t[12];
    for (int r = 0; r < RESCUE_ROUNDS; r++) {
        for (int i = 0; i < 12; i++) { u64 x = s[i], x2 = fp_mul(x, x), x3 = fp_mul(x2, x), x4 = fp_mul(x2, x2); t[i] = fp_mul(x3, x4); }
        for (int i = 0; i < 12; i++) {
            u128 acc = RESCUE_ARK[(2 * r) * 12 + i];
            for (int j = 0; j < 12; j++) acc += (u128)RESCUE_MDS[i * 12 + j] * t[j];
            s[i] = fp_red(acc);
        }
        for (int i = 0; i < 12; i++) t[i] = fp_pow(s[i], RESCUE_INV_ALPHA);
        for (int i = 0; i < 12; i++) {
            u128 acc = RESCUE_ARK[(2 * r + 1) * 12 + i];
            for (int j = 0; j < 12; j++) acc += (u128)RESCUE_MDS[i * 12 + j] * t[j];
            s[i] = fp_red(acc);
        }
    }
}
/* sponge absorb helper: RescueHash::hash_field restated in streaming form */
typedef struct { u64 s[12]; int i; } sponge;
static void sponge_init(sponge *sp) { memset(sp, 0, sizeof *sp); }
static void sponge_absorb(sponge *sp, u64 e) {
    sp->s[sp->i] = fp_add(sp->s[sp->i], e);
    if (++sp->i == 8) { rescue_permutation(sp->s); sp->i = 0; }
}
static void sponge_finish(sponge *sp, u8 *out32) {
    if (sp->i > 0) { sp->s[sp->i] = fp_add(sp->s[sp->i], 1); rescue_permutation(sp->s); }
    memcpy(out32, sp->s, 32);
}
/* hash_message, src/signature.rs:274-306 */
static void hash_message(fp6 rx, aff pk, const u8 *msg, u64 len, u8 *out32) {
    sponge sp; sponge_init(&sp);
    for (int i = 0; i < 6; i++) sponge_absorb(&sp, rx.c[i]);
    for (int i = 0; i < 6; i++) sponge_absorb(&sp, pk.x.c[i]);
    sponge_absorb(&sp, pk.y.c[0]);
    u64 nb = len / 7;
    for (u64 c = 0; c < nb; c++) { u64 v = 0; memcpy(&v, msg + 7 * c, 7); sponge_absorb(&sp, v); }
    u64 rem = len - 7 * nb;
    if (rem) { u8 buf[8] = {0}; memcpy(buf, msg + 7 * nb, rem); buf[rem] = 1; u64 v; memcpy(&v, buf, 8); sponge_absorb(&sp, v); }
    sponge_finish(&sp, out32);
}

/* ------------------------------------------------------------------ protocol */
static aff load_pk(const u8 *pk96, const u8 *pk_inf, u64 i) {
    aff p; memcpy(p.x.c, pk96 + 96 * i, 48); memcpy(p.y.c, pk96 + 96 * i + 48, 48);
    p.inf = pk_inf ? pk_inf[i] != 0 : 0;
    if (p.inf) { p.x = F6_ZERO; p.y = F6_ZERO; }
    return p;
}
/* Signature::verify, src/signature.rs:181-205 -> 0 ok, 1 InvalidPublicKey, 2 InvalidSignature, 3 malformed (reference panics) */
static int pk_canonical(const u8 *pk96, u64 i) {
    for (int k = 0; k < 12; k++) { u64 v; memcpy(&v, pk96 + 96 * i + 8 * k, 8); if (v >= PP) return 0; }
    return 1;
}
static int scalar_canonical(const u8 *b32) { u64 t[4]; memcpy(t, b32, 32); return !sc_geq_q(t); }
/* States the reference's types cannot hold (Scalar >= q, Fp limb >= p inside a PublicKey) are
 * reported as 3 before anything else; a canonical-but-off-subgroup key is 1; a non-canonical sig.x
 * limb (Fp6::from_bytes(..).unwrap() panics AFTER the subgroup check, :182-186) is 3. */
static u8 verify_one(const u8 *sig81, aff pk, const u8 *msg, u64 len) {
    if (!scalar_canonical(sig81 + 49)) return 3;
    if (!is_torsion_free(pk)) return 1;
    fp6 x;
    if (!f6_from_bytes(sig81, &x)) return 3; /* flag byte sig81[48] ignored */
    u8 h[32];
    hash_message(x, pk, msg, len, h);
    sc hs = sc_from_bytes_reduce(h);
    /* sig.e arrives as a canonical Scalar in the reference; reduce defensively (no-op when canonical) */
    sc es = sc_from_bytes_reduce(sig81 + 49);
    aff r = to_aff(jac_mul2(pk, (const u8 *)hs.l, generator(), (const u8 *)es.l));
    return f6_eq(r.x, x) ? 0 : 2; /* identity has x = 0 */
}

typedef struct {
    int tid, nt; u64 n;
    const u8 *a, *b, *c, *d, *e; const u64 *off; u8 *o1, *o2;
    jac partial; sc lin; int bad;
} job;

static void run_jobs(void *(*fn)(void *), job *tmpl, int nt) {
    if (nt < 1) nt = 1;
    pthread_t *th = malloc(sizeof(pthread_t) * nt);
    job *js = malloc(sizeof(job) * nt);
    for (int t = 0; t < nt; t++) { js[t] = *tmpl; js[t].tid = t; js[t].nt = nt; pthread_create(&th[t], 0, fn, &js[t]); }
    for (int t = 0; t < nt; t++) pthread_join(th[t], 0);
    memcpy(tmpl, js, sizeof(job)); /* thread 0's result slot (callers that reduce use js directly) */
    free(th); free(js);
}

static void *verify_worker(void *arg) {
    job *j = arg;
    for (u64 i = j->tid; i < j->n; i += j->nt)
        j->o1[i] = (!(j->c && j->c[i]) && !pk_canonical(j->b, i))
                       ? 3
                       : verify_one(j->a + 81 * i, load_pk(j->b, j->c, i), j->d + j->off[i], j->off[i + 1] - j->off[i]);
    return 0;
}
API int cref_verify_many(u64 n, const u8 *sigs81, const u8 *pk96, const u8 *pk_inf, const u8 *msgs,
                         const u64 *off, u8 *verdicts, int nthreads) {
    job t = {0}; t.n = n; t.a = sigs81; t.b = pk96; t.c = pk_inf; t.d = msgs; t.off = off; t.o1 = verdicts;
    run_jobs(verify_worker, &t, nthreads);
    return 0;
}

static void *hash_worker(void *arg) {
    job *j = arg;
    for (u64 i = j->tid; i < j->n; i += j->nt) {
        fp6 x; memcpy(x.c, j->a + 48 * i, 48);
        hash_message(x, load_pk(j->b, 0, i), j->d + j->off[i], j->off[i + 1] - j->off[i], j->o1 + 32 * i);
    }
    return 0;
}
/* rx48: n x 48 bytes (canonical limbs assumed), pk96: n x (x||y) */
API int cref_hash_messages(u64 n, const u8 *rx48, const u8 *pk96, const u8 *msgs, const u64 *off, u8 *digests, int nthreads) {
    job t = {0}; t.n = n; t.a = rx48; t.b = pk96; t.d = msgs; t.off = off; t.o1 = digests;
    run_jobs(hash_worker, &t, nthreads);
    return 0;
}

static void *keygen_worker(void *arg) {
    job *j = arg;
    for (u64 i = j->tid; i < j->n; i += j->nt) {
        sc k = sc_from_bytes_reduce(j->a + 32 * i);
        aff p = to_aff(jac_mul(generator(), (const u8 *)k.l));
        memcpy(j->o1 + 96 * i, p.x.c, 48); memcpy(j->o1 + 96 * i + 48, p.y.c, 48);
        if (j->o2) j->o2[i] = (u8)p.inf;
    }
    return 0;
}
/* PublicKey::from(&PrivateKey), src/public.rs:26-32 */
API int cref_keygen(u64 n, const u8 *sk32, u8 *pk96, u8 *pk_inf, int nthreads) {
    job t = {0}; t.n = n; t.a = sk32; t.o1 = pk96; t.o2 = pk_inf;
    run_jobs(keygen_worker, &t, nthreads);
    return 0;
}

static void *sign_worker(void *arg) {
    job *j = arg;
    for (u64 i = j->tid; i < j->n; i += j->nt) {
        sc sk = sc_from_bytes_reduce(j->a + 32 * i), r = sc_from_bytes_reduce(j->e + 32 * i);
        aff pk = load_pk(j->b, j->c, i);
        aff R = to_aff(jac_mul(generator(), (const u8 *)r.l));
        u8 h[32];
        hash_message(R.x, pk, j->d + j->off[i], j->off[i + 1] - j->off[i], h);
        sc hs = sc_from_bytes_reduce(h);
        sc e = sc_add(r, sc_neg(sc_mul(sk, hs)));
        compress(R, j->o1 + 81 * i);
        memcpy(j->o1 + 81 * i + 49, e.l, 32);
    }
    return 0;
}
/* KeyPair::sign, src/signature.rs:114-129, nonce supplied by the caller */
API int cref_sign_many(u64 n, const u8 *sk32, const u8 *pk96, const u8 *pk_inf, const u8 *msgs, const u64 *off,
                       const u8 *nonce32, u8 *sigs81, int nthreads) {
    job t = {0}; t.n = n; t.a = sk32; t.b = pk96; t.c = pk_inf; t.d = msgs; t.off = off; t.e = nonce32; t.o1 = sigs81;
    run_jobs(sign_worker, &t, nthreads);
    return 0;
}

/* verify_batch, src/batch.rs:31-130, caller-supplied randomisers (32 LE bytes each, reduced mod q) */
static void *batch_worker(void *arg) {
    job *j = arg;
    jac acc = jac_inf(); sc lin = {{0, 0, 0, 0}}; j->bad = 0;
    for (u64 i = j->tid; i < j->n; i += j->nt) {
        const u8 *sig = j->a + 81 * i;
        aff pk = load_pk(j->b, j->c, i);
        fp6 x;
        if (!scalar_canonical(sig + 49) || (!(j->c && j->c[i]) && !pk_canonical(j->b, i))) { j->bad = 1; break; }
        if (!f6_from_bytes(sig, &x)) { j->bad = 1; break; }            /* unwrap panic, batch.rs:67 */
        u8 h[32];
        hash_message(x, pk, j->d + j->off[i], j->off[i + 1] - j->off[i], h);
        sc hs = sc_from_bytes_reduce(h), s = sc_from_bytes_reduce(j->e + 32 * i), e = sc_from_bytes_reduce(sig + 49);
        lin = sc_add(lin, sc_mul(s, e));                                 /* batch.rs:92-97 */
        aff R;
        if (!decompress(sig, &R)) { j->bad = 1; break; }                 /* unwrap panic, batch.rs:104 */
        acc = jac_add(acc, jac_mul(R, (const u8 *)s.l));
        aff np = pk; np.y = f6_neg(pk.y);                                /* batch.rs:106 */
        sc hsx = sc_mul(hs, s);                                          /* batch.rs:109-111 */
        acc = jac_add(acc, jac_mul(np, (const u8 *)hsx.l));
    }
    j->partial = acc; j->lin = lin;
    return 0;
}
API int cref_verify_batch(u64 n, const u8 *sigs81, const u8 *pk96, const u8 *pk_inf, const u8 *msgs, const u64 *off,
                          const u8 *rand32, int *verdict, u8 *lhs97, u8 *rhs97, int nthreads) {
    if (nthreads < 1) nthreads = 1;
    pthread_t *th = malloc(sizeof(pthread_t) * nthreads);
    job *js = calloc(nthreads, sizeof(job));
    for (int t = 0; t < nthreads; t++) {
        js[t].tid = t; js[t].nt = nthreads; js[t].n = n; js[t].a = sigs81; js[t].b = pk96; js[t].c = pk_inf;
        js[t].d = msgs; js[t].off = off; js[t].e = rand32;
        pthread_create(&th[t], 0, batch_worker, &js[t]);
    }
    jac acc = jac_inf(); sc lin = {{0, 0, 0, 0}}; int bad = 0;
    for (int t = 0; t < nthreads; t++) {
        pthread_join(th[t], 0);
        acc = jac_add(acc, js[t].partial); lin = sc_add(lin, js[t].lin); bad |= js[t].bad;
    }
    free(th); free(js);
    if (bad) { *verdict = 3; return 0; }
    aff lhs = to_aff(acc), rhs = to_aff(jac_mul(generator(), (const u8 *)lin.l));   /* batch.rs:98-100,123 */
    if (lhs97) { memcpy(lhs97, lhs.x.c, 48); memcpy(lhs97 + 48, lhs.y.c, 48); lhs97[96] = (u8)lhs.inf; }
    if (rhs97) { memcpy(rhs97, rhs.x.c, 48); memcpy(rhs97 + 48, rhs.y.c, 48); rhs97[96] = (u8)rhs.inf; }
    *verdict = f6_eq(lhs.x, rhs.x) ? 0 : 2;                                          /* batch.rs:125-129 */
    return 0;
}

/* ------------------------------------------------------------------ low-level probes for unit tests */
API void cref_fp6_mul(const u64 *a, const u64 *b, u64 *r) { fp6 x, y; memcpy(x.c, a, 48); memcpy(y.c, b, 48); x = f6_mul(x, y); memcpy(r, x.c, 48); }
API void cref_fp6_inv(const u64 *a, u64 *r) { fp6 x; memcpy(x.c, a, 48); x = f6_inv(x); memcpy(r, x.c, 48); }
API int cref_fp6_sqrt(const u64 *a, u64 *r) { fp6 x, y; memcpy(x.c, a, 48); int ok = f6_sqrt(x, &y); if (ok) memcpy(r, y.c, 48); return ok; }
API void cref_rescue_permutation(u64 *state12) { rescue_permutation(state12); }
API void cref_pt_mul(const u8 *pt96, int inf, const u8 *k32, u8 *out96, int *out_inf) {
    u8 z = (u8)inf; aff p = load_pk(pt96, &z, 0);
    aff r = to_aff(jac_mul(p, k32));
    memcpy(out96, r.x.c, 48); memcpy(out96 + 48, r.y.c, 48); *out_inf = r.inf;
}
API void cref_pt_add(const u8 *a96, int ainf, const u8 *b96, int binf, u8 *out96, int *out_inf) {
    u8 za = (u8)ainf, zb = (u8)binf;
    aff r = to_aff(jac_add(to_jac(load_pk(a96, &za, 0)), to_jac(load_pk(b96, &zb, 0))));
    memcpy(out96, r.x.c, 48); memcpy(out96 + 48, r.y.c, 48); *out_inf = r.inf;
}
API int cref_is_torsion_free(const u8 *pt96, int inf) { u8 z = (u8)inf; return is_torsion_free(load_pk(pt96, &z, 0)); }
API void cref_compress(const u8 *pt96, int inf, u8 *out49) { u8 z = (u8)inf; compress(load_pk(pt96, &z, 0), out49); }
API int cref_decompress(const u8 *in49, u8 *out96, int *out_inf) {
    aff p; int ok = decompress(in49, &p);
    memcpy(out96, p.x.c, 48); memcpy(out96 + 48, p.y.c, 48); *out_inf = p.inf; return ok;
}
API void cref_scalar_mul(const u8 *a32, const u8 *b32, u8 *r32) { sc r = sc_mul(sc_from_bytes_reduce(a32), sc_from_bytes_reduce(b32)); memcpy(r32, r.l, 32); }
API void cref_scalar_reduce(const u8 *a32, u8 *r32) { sc r = sc_from_bytes_reduce(a32); memcpy(r32, r.l, 32); }
