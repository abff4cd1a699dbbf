// Affine-coordinate fast path of the verification (Signature::verify, src/signature.rs:181-205).
//
// In Fp6 an inversion is cheap relative to a multiplication: through the tower Fp6 = Fp3[u]/(u^2 - v)
// and the norm Fp3 -> Fp it costs ~45 base-field products plus ONE inversion in the 64-bit field Fp,
// and the Fp inversions of several independent denominators are shared (Montgomery's trick on the
// norms).  An affine addition is then 2M + 1S + (share of an inversion) instead of the 11M + 5S of a
// Jacobian addition, which more than halves the cost of the 150 bucket / table additions of a
// verification; the doubling chain D_j = 2^j P costs about the same as in Jacobian form.
//
// The affine formulas have exceptional inputs (P + P, P + (-P), the identity).  They cannot occur for
// honest keys except with negligible probability, but adversarial small-order keys reach them, so the
// fast path DETECTS every such event (a zero denominator, an identity operand it cannot represent)
// and reports FAST_EXCEPTIONAL; the caller then re-runs that signature through the exact Jacobian
// routine (curve.cuh: torsion_check_and_mul).  A result that is not flagged is the exact group
// element, so verdicts stay bit-identical to the reference.
#pragma once
#include "curve.cuh"

namespace sb {

struct aff_pt {
    fp6 x, y;
};

// ---- Fp3 squaring, 6 products -------------------------------------------------------------------
SB_DEV fp3 fp3_sqr6(const fp3& a) {
    fp_t a2_7 = fp_mul7_nc(a.c[2]);
    fp3 r;
    wide_acc w;
    wide_zero(w);
    wide_mac(w, a.c[1], a2_7);
    wide_double(w);
    wide_mac_sqr(w, a.c[0]);
    r.c[0] = wide_reduce(w);
    wide_zero(w);
    wide_mac(w, a.c[0], a.c[1]);
    wide_double(w);
    wide_mac(w, a.c[2], a2_7);
    r.c[1] = wide_reduce(w);
    wide_zero(w);
    wide_mac(w, a.c[0], a.c[2]);
    wide_double(w);
    wide_mac_sqr(w, a.c[1]);
    r.c[2] = wide_reduce(w);
    return r;
}

// adjugate and norm of d in Fp3 with lazily accumulated products (12 products, 4 reductions):
//   d^-1 = (t0 + t1 v + t2 v^2) / n,   t0 = d0^2 - 7 d1 d2,  t1 = 7 d2^2 - d0 d1,  t2 = d1^2 - d0 d2,
//   n = d0 t0 + 7 (d2 t1 + d1 t2)
SB_DEV void fp3_adj_norm_lazy(const fp3& d, fp3& adj, fp_t& norm) {
    fp_t d0 = d.c[0], d1 = d.c[1], d2 = d.c[2];
    fp_t d1_7 = fp_mul7_nc(d1), d2_7 = fp_mul7_nc(d2);
    fp_t nd0 = FP_P - d0, nd1 = FP_P - d1;  // negatives as 64-bit representatives (p itself for 0)
    wide_acc w;
    wide_zero(w);
    wide_mac_sqr(w, d0);
    wide_mac(w, nd1, d2_7);
    fp_t t0 = wide_reduce(w);
    wide_zero(w);
    wide_mac(w, d2, d2_7);
    wide_mac(w, nd0, d1);
    fp_t t1 = wide_reduce(w);
    wide_zero(w);
    wide_mac_sqr(w, d1);
    wide_mac(w, nd0, d2);
    fp_t t2 = wide_reduce(w);
    wide_zero(w);
    wide_mac(w, d0, t0);
    wide_mac(w, d2_7, t1);
    wide_mac(w, d1_7, t2);
    norm = wide_reduce(w);
    adj = fp3{{t0, t1, t2}};
}

// a^-1 for any 64-bit representative a of a non-zero element: 64 squarings + 10 multiplications,
// the whole chain in non-canonical form.
//   t31 = a^(2^31 - 1),  t32 = t31^2 a = a^(2^32 - 1),  a^(p-2) = t31^(2^33) t32
SB_DEV_NOINLINE fp_t fp_inv_chain(fp_t a) {
    fp_t t2 = fp_mul_nc(fp_sqr_nc(a), a);
    fp_t t4 = fp_mul_nc(fp_sqr_n_nc(t2, 2), t2);
    fp_t t8 = fp_mul_nc(fp_sqr_n_nc(t4, 4), t4);
    fp_t t16 = fp_mul_nc(fp_sqr_n_nc(t8, 8), t8);
    fp_t t24 = fp_mul_nc(fp_sqr_n_nc(t16, 8), t8);
    fp_t t28 = fp_mul_nc(fp_sqr_n_nc(t24, 4), t4);
    fp_t t30 = fp_mul_nc(fp_sqr_n_nc(t28, 2), t2);
    fp_t t31 = fp_mul_nc(fp_sqr_nc(t30), a);
    fp_t t32 = fp_mul_nc(fp_sqr_nc(t31), a);
    return fp_mul_nc(fp_sqr_n_nc(t31, 33), t32);
}

static constexpr int AFF_MAX_BATCH = 10;

// d[i] <- d[i]^-1 for i < k (k <= AFF_MAX_BATCH) with a single Fp inversion.  Returns the bit mask of the
// elements that are zero (their slots are left unspecified; the others are still inverted correctly).
SB_DEV_NOINLINE uint32_t fp6_batch_inv(fp6* d, int k) {
    fp3 adj[AFF_MAX_BATCH];
    fp_t nrm[AFF_MAX_BATCH], pre[AFF_MAX_BATCH];
    uint32_t zero_mask = 0;
    fp_t run = 1;
#pragma unroll 1
    for (int i = 0; i < k; i++) {
        fp3 a0, a1;
        fp6_split(d[i], a0, a1);
        fp3 s0 = fp3_sqr6(a0), s1 = fp3_sqr6(a1);
        // N = a0^2 - v a1^2,  v (x0, x1, x2) = (7 x2, x0, x1)
        fp3 nn = fp3{{fp_sub(s0.c[0], fp_mul7(s1.c[2])), fp_sub(s0.c[1], s1.c[0]), fp_sub(s0.c[2], s1.c[1])}};
        fp_t n;
        fp3_adj_norm_lazy(nn, adj[i], n);
        bool z = n == 0;  // the norm of a field element vanishes only for zero
        if (z) zero_mask |= 1u << i;
        n = z ? 1 : n;
        nrm[i] = n;
        pre[i] = run;
        run = fp_mul_nc(run, n);
    }
    fp_t inv = fp_inv_chain(run);
#pragma unroll 1
    for (int i = k - 1; i >= 0; i--) {
        fp_t ni = fp_mul_nc(inv, pre[i]);  // n_i^-1
        inv = fp_mul_nc(inv, nrm[i]);
        fp3 a0, a1;
        fp6_split(d[i], a0, a1);
        fp3 s;
#pragma unroll
        for (int c = 0; c < 3; c++) s.c[c] = fp_mul_nc(adj[i].c[c], ni);
        fp3 na1 = fp3{{FP_P - a1.c[0], FP_P - a1.c[1], FP_P - a1.c[2]}};
        d[i] = fp6_join(fp3_mul(a0, s), fp3_mul(na1, s));
    }
    return zero_mask;
}

// ---- batched affine point operations ---------------------------------------------------------------
// Per-thread mode of one operation (data-dependent, evaluated with selects: control flow stays uniform)
enum aff_mode : uint8_t {
    AOP_NOP = 0,     // leave acc unchanged
    AOP_ADD = 1,     // acc <- acc + src      (or acc <- 2 acc for a doubling entry)
    AOP_SUB = 2,     // acc <- acc - src
    AOP_SET = 3,     // acc <- src            (acc was the identity)
    AOP_SETNEG = 4,  // acc <- -src
};
struct aff_op {
    aff_pt* acc;
    const aff_pt* src;  // nullptr (warp-uniform) marks a doubling of acc
    uint8_t mode;
};
SB_DEV uint8_t aff_add_mode(bool acc_empty, bool src_empty, bool neg) {
    return src_empty ? AOP_NOP : (acc_empty ? (neg ? AOP_SETNEG : AOP_SET) : (neg ? AOP_SUB : AOP_ADD));
}

// Runs k independent operations with one shared inversion.  All operands are read before anything is
// written EXCEPT that operation i may read (as src) a point that a LATER operation j > i overwrites
// (as acc) -- the main loop adds D_j into buckets and then doubles it in the same batch.
// Returns true when an active operation met a zero denominator (P + P, P - P, doubling a 2-torsion
// point): the caller must abandon the fast path.
SB_DEV_NOINLINE bool aff_batch(const aff_op* ops, int k) {
    fp6 den[AFF_MAX_BATCH];
#pragma unroll 1
    for (int i = 0; i < k; i++) {
        const aff_pt* a = ops[i].acc;
        const aff_pt* s = ops[i].src;
        fp6 dd = s ? fp6_sub(a->x, s->x) : fp6_dbl(a->y);
        bool active = ops[i].mode == AOP_ADD || ops[i].mode == AOP_SUB;
        den[i] = active ? dd : fp6_one();
    }
    bool exceptional = fp6_batch_inv(den, k) != 0;
#pragma unroll 1
    for (int i = 0; i < k; i++) {
        aff_pt* a = ops[i].acc;
        const aff_pt* s = ops[i].src;
        uint8_t mode = ops[i].mode;
        fp6 ax = a->x, ay = a->y, num, bx, sy;
        if (s) {
            bx = s->x;
            sy = s->y;
            if (mode == AOP_SUB || mode == AOP_SETNEG) sy = fp6_neg(sy);
            num = fp6_sub(ay, sy);
        } else {
            bx = ax;
            sy = ay;
            fp6 xx = fp6_sqr(ax);
            num = fp6_add(fp6_dbl(xx), xx);
            num.c[0] = fp_add(num.c[0], 1);  // 3 x^2 + a, a = 1
        }
        fp6 lam = fp6_mul(num, den[i]);
        fp6 x3 = fp6_sub(fp6_sub(fp6_sqr(lam), ax), bx);
        fp6 y3 = fp6_sub(fp6_mul(lam, fp6_sub(ax, x3)), ay);
        bool active = mode == AOP_ADD || mode == AOP_SUB;
        bool set = mode == AOP_SET || mode == AOP_SETNEG;
#pragma unroll
        for (int c = 0; c < 6; c++) {
            a->x.c[c] = active ? x3.c[c] : (set ? bx.c[c] : ax.c[c]);
            a->y.c[c] = active ? y3.c[c] : (set ? sy.c[c] : ay.c[c]);
        }
    }
    return exceptional;
}

enum fast_result : int {
    FAST_TORSION_FREE = 0,      // [q]P == O, h*P + e*G computed
    FAST_NOT_TORSION_FREE = 1,  // [q]P != O
    FAST_EXCEPTIONAL = 2,       // an exceptional case was met: use the exact routine
};

// [q]P == O ?  and  R = h*P + e*G  (affine), sharing the doubling chain D_j = 2^j P as torsion_check_and_mul
// does, but with every point affine.  `Dp` is caller-provided storage for D_j (shared memory in the kernels).
// On FAST_TORSION_FREE, *R is the result (never the identity: that case is reported as exceptional).
SB_DEV int verify_core_affine(const fp6& px, const fp6& py, const scalar& h, const scalar& e,
                              const uint64_t* __restrict__ gtab, aff_pt* R, aff_pt* Dp) {
    aff_pt W[GTAB_WINDOWS];  // buckets during the chain (Bq = W[0..8), Bh = W[8..16)), table points afterwards
    aff_pt* Bq = W;
    aff_pt* Bh = W + 8;
    aff_op ops[AFF_MAX_BATCH];
    int8_t hd[64];
    recode_signed_w4(h, hd);
    uint32_t q_seen = 0, h_seen = 0;
    bool exc = false;
#pragma unroll 1
    for (int b = 0; b < 16; b++) W[b] = aff_pt{fp6_zero(), fp6_zero()};  // dummy operations read empty buckets
    Dp->x = px;
    Dp->y = py;
#pragma unroll 1
    for (int j = 0; j < 256; j++) {
        if ((j & 3) == 0) SB_PHASE_SYNC(1);
        int k = 0;
        int dq = SB_QWNAF(j);
        if (dq != 0) {  // warp-uniform
            int idx = (dq < 0 ? -dq : dq) >> 1;
            ops[k].acc = &Bq[idx];
            ops[k].src = Dp;
            ops[k].mode = aff_add_mode(!((q_seen >> idx) & 1), false, dq < 0);
            q_seen |= 1u << idx;
            k++;
        }
        if ((j & 3) == 0) {
            int dh = hd[j >> 2];
            int mag = dh < 0 ? -dh : dh;
            int idx = mag ? mag - 1 : 0;
            ops[k].acc = &Bh[idx];
            ops[k].src = Dp;
            ops[k].mode = aff_add_mode(!((h_seen >> idx) & 1), mag == 0, dh < 0);
            if (mag) h_seen |= 1u << idx;
            k++;
        }
        if (j < 255) {
            ops[k].acc = Dp;
            ops[k].src = nullptr;
            ops[k].mode = AOP_ADD;
            k++;
        }
        if (k) exc |= aff_batch(ops, k);
    }
    // Bucket aggregation, both scalars in lockstep (R_k = sum_{m>=k} B_m, O_k = sum_{m>=k} R_m):
    //   q (odd digits 2k+1):  [q]P = 2 O_1 + R_0        h (digits m = k+1):  h*P = O_0
    aff_pt Rq = Bq[7], Oq = Bq[7], Rh = Bh[7], Oh = Bh[7];
    bool eRq = !((q_seen >> 7) & 1), eOq = eRq, eRh = !((h_seen >> 7) & 1), eOh = eRh;
#pragma unroll 1
    for (int t = 1; t <= 8; t++) {
        SB_PHASE_SYNC(1);
        int k = 0;
        if (t >= 2) {  // O += R (the value of R before this round's update)
            if (t <= 7) {
                ops[k].acc = &Oq;
                ops[k].src = &Rq;
                ops[k].mode = aff_add_mode(eOq, eRq, false);
                eOq = eOq && eRq;
            } else {  // t == 8: O_q <- 2 O_1
                ops[k].acc = &Oq;
                ops[k].src = nullptr;
                ops[k].mode = eOq ? AOP_NOP : AOP_ADD;
            }
            k++;
            ops[k].acc = &Oh;
            ops[k].src = &Rh;
            ops[k].mode = aff_add_mode(eOh, eRh, false);
            eOh = eOh && eRh;
            k++;
        }
        if (t <= 7) {  // R += B_{7-t}
            int b = 7 - t;
            bool eb = !((q_seen >> b) & 1);
            ops[k].acc = &Rq;
            ops[k].src = &Bq[b];
            ops[k].mode = aff_add_mode(eRq, eb, false);
            eRq = eRq && eb;
            k++;
            eb = !((h_seen >> b) & 1);
            ops[k].acc = &Rh;
            ops[k].src = &Bh[b];
            ops[k].mode = aff_add_mode(eRh, eb, false);
            eRh = eRh && eb;
            k++;
        }
        exc |= aff_batch(ops, k);
    }
    // [q]P = Oq + Rq is the identity  <=>  Oq == -Rq
    bool x_eq = fp6_eq(Oq.x, Rq.x);
    bool y_opp = fp6_eq(Oq.y, fp6_neg(Rq.y));
    if (eOq || eRq) exc = true;                // degenerate digit pattern: leave it to the exact routine
    if (x_eq && !y_opp) exc = true;            // 2 O_1 == R_0: a doubling
    bool torsion_free = x_eq && y_opp;

    // e*G: the 20 table points and h*P summed as a tree (5 shared inversions)
    aff_pt hP = Oh;
    bool e_hP = eOh;
    uint32_t t_empty = 0;
    {
        int carry = 0;
#pragma unroll 1
        for (int i = 0; i < GTAB_WINDOWS; i++) {
            int raw = (int)sc_bits(e, GTAB_W * i, GTAB_W) + carry;
            bool neg = raw > (1 << (GTAB_W - 1));
            carry = neg ? 1 : 0;
            int d = neg ? (1 << GTAB_W) - raw : raw;
            const uint64_t* ent = gtab + ((size_t)i * GTAB_ENTRIES + (d ? d : 1)) * GTAB_ENTRY_U64;
            fp6 qx, qy;
#if defined(__CUDA_ARCH__)
            const ulonglong2* e2 = reinterpret_cast<const ulonglong2*>(ent);
            ulonglong2 a = e2[0], b = e2[1], c = e2[2], dd = e2[3], ee = e2[4], f = e2[5];
            qx = fp6{{a.x, a.y, b.x, b.y, c.x, c.y}};
            qy = fp6{{dd.x, dd.y, ee.x, ee.y, f.x, f.y}};
#else
            for (int c = 0; c < 6; c++) {
                qx.c[c] = ent[c];
                qy.c[c] = ent[6 + c];
            }
#endif
            if (neg) qy = fp6_neg(qy);
            W[i].x = qx;
            W[i].y = qy;
            if (d == 0) t_empty |= 1u << i;
        }
    }
    // level strides 1, 2, 4, 8, 16 over W[0..20); h*P joins W[16] at the third level
#pragma unroll 1
    for (int lvl = 0; lvl < 5; lvl++) {
        SB_PHASE_SYNC(1);
        int stride = 1 << lvl, k = 0;
#pragma unroll 1
        for (int i = 0; i + stride < GTAB_WINDOWS; i += 2 * stride) {
            bool ea = (t_empty >> i) & 1, eb = (t_empty >> (i + stride)) & 1;
            ops[k].acc = &W[i];
            ops[k].src = &W[i + stride];
            ops[k].mode = aff_add_mode(ea, eb, false);
            if (!(ea && eb)) t_empty &= ~(1u << i);
            k++;
        }
        if (lvl == 2) {
            bool ea = (t_empty >> 16) & 1;
            ops[k].acc = &W[16];
            ops[k].src = &hP;
            ops[k].mode = aff_add_mode(ea, e_hP, false);
            if (!(ea && e_hP)) t_empty &= ~(1u << 16);
            k++;
        }
        exc |= aff_batch(ops, k);
    }
    if (t_empty & 1) exc = true;  // the result is the identity: exact routine
    *R = W[0];
    if (exc) return FAST_EXCEPTIONAL;
    return torsion_free ? FAST_TORSION_FREE : FAST_NOT_TORSION_FREE;
}

}  // namespace sb
