// Lowest-latency form of Signature::verify (src/signature.rs:181-205) for calls of a FEW signatures -- the reference's own
// Criterion case is one verification (benches/schnorr.rs:22-66): one thread block per signature.
//
// The six-lane kernel (dist.cuh) walks one dependency chain per signature: challenge hash, 252 doublings interleaved with
// ~100 bucket additions, 27 additions of bucket aggregation, 20 table additions -- 0.93 ms however idle the GPU is.  Most
// of that chain is not a true dependency: only the doublings D_j = 2^j P depend on each other.  Here the four warps of a
// block split the work of ONE signature in three barrier-separated phases (same arithmetic, same six-lane primitives):
//   A  warp 0: the doubling chain, every D_j stored to shared memory (253 x 104 B), with warp 3 computing the one product
//      of a doubling that does not depend on the cofactor (X^2);   warp 1: the challenge hash and its signed digits;
//      warp 2: e*G from the fixed-base table  -- side by side
//   B  the sixteen bucket accumulations (eight of the subgroup check, eight of the challenge) on sixteen six-lane groups
//      at once, each group adding the chain points of ITS bucket (a compacted list, so a warp's groups stay in step)
//   C  warp 0: aggregation of the subgroup-check buckets and the [q]P == O test;   warp 1: aggregation of the challenge
//      buckets, + e*G, x-only comparison
// Latency = the doubling chain + ~25 additions instead of chain + ~150 additions + hash.  Group elements, exceptional-case
// policy (anything the chord-and-tangent formulas cannot evaluate is handed to the exact kernel) and verdicts are those
// of dverify_core / verify_points_fast: additions are merely re-associated.
#pragma once
#include "dist.cuh"

namespace sb {

static constexpr int ONE_THREADS = 128;
static constexpr int ONE_SLOTS = 16;  // bucket accumulations: 0..7 subgroup check (odd digits of q), 8..15 challenge (|d| = 1..8)

struct one_shared {
    fp_t cx[SB_CHAIN_STEPS][6], cy[SB_CHAIN_STEPS][6], cw[SB_CHAIN_STEPS];  // the chain D_j
    fp_t bx[ONE_SLOTS][6], by[ONE_SLOTS][6], bw[ONE_SLOTS];                 // bucket sums
    fp_t ex[6], ey[6], ew;                                                  // e*G
    fp_t xx[6];                                                             // X^2 of the current chain point (helper warp)
    uint32_t mds2[24];
    int8_t hd[64];                          // signed 4-bit digits of the challenge
    uint8_t lst[ONE_SLOTS][64];             // per bucket: chain steps to add (bit 7 of neg[] separately: steps reach 252)
    uint8_t neg[ONE_SLOTS][64];
    int cnt[ONE_SLOTS];
    int e_empty, exc, torsion_free;
};

__device__ __forceinline__ dpt one_load_chain(const one_shared& S, int j, int k) { return dpt{S.cx[j][k], S.cy[j][k], S.cw[j]}; }
__device__ __forceinline__ dpt one_load_bucket(const one_shared& S, int s, int k) { return dpt{S.bx[s][k], S.by[s][k], S.bw[s]}; }

__global__ void __launch_bounds__(ONE_THREADS) k_verify_one(soa_batch in, const uint8_t* __restrict__ msgs,
                                                            const uint64_t* __restrict__ msg_off,
                                                            const uint64_t* __restrict__ gtab, uint8_t* __restrict__ verdicts,
                                                            uint32_t* __restrict__ work_list, uint32_t* __restrict__ work_count) {
    __shared__ one_shared S;
    // Every lane of a warp runs the same instruction stream (full-warp shuffles, dist.cuh: FULL): the five six-lane
    // groups of a warp either replicate one computation (phases A and C) or work on five buckets side by side (phase B);
    // lanes 30 and 31 shadow lanes 0 and 1.  Only `real` lanes of the group in charge store results.
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const bool ghost = lane >= 30;
    const int g = ghost ? 0 : lane / 6, k = ghost ? lane - 30 : lane % 6;
    const bool first = lane < 6;             // the lanes that publish a replicated result
    const int gbase = 6 * g;
    const unsigned mask = 0xffffffffu;
    constexpr bool F = true;
    const size_t i = blockIdx.x, n = in.n;
    const uint8_t fl = in.flags[i];          // block-uniform from here on
    if (fl & (FL_MALFORMED | FL_PK_INF)) {
        if (tid == 0) {
            uint8_t v = (fl & FL_MALFORMED) ? VERDICT_MALFORMED : VERDICT_NEEDS_EXACT;  // the identity key is the exact kernel's business
            verdicts[i] = v;
            if (v == VERDICT_NEEDS_EXACT) work_list[atomicAdd(work_count, 1u)] = (uint32_t)i;
        }
        return;
    }
    if (tid < 24) S.mds2[tid] = c_mds_row[tid % 12];
    if (tid < ONE_SLOTS) S.cnt[tid] = 0;
    if (tid == 0) {
        S.exc = 0;
        S.e_empty = 1;
        S.torsion_free = 0;
    }
    __syncthreads();
    const bool x_ok = !(fl & FL_X_BAD);
    const uint64_t* pl = reinterpret_cast<const uint64_t*>(in.planes);
    fp_t sx = pl[((size_t)(0 + (k >> 1)) * n + i) * 2 + (k & 1)];
    fp_t px = pl[((size_t)(5 + (k >> 1)) * n + i) * 2 + (k & 1)];
    fp_t py = pl[((size_t)(8 + (k >> 1)) * n + i) * 2 + (k & 1)];
    bool exc = false;

    // ---- phase A: doubling chain (+ helper) | challenge | e*G -------------------------------------------------------
    // The chain is the critical path of the whole kernel.  Inside one doubling only X^2 is independent of the
    // cofactor of Y: warp 3 computes it while warp 0 works on the cofactor (two named barriers per step between the 64
    // threads of the two warps; jf_dbl of affine.cuh / djf_dbl of dist.cuh otherwise).
    if (warp == 0 || warp == 3) {
        dpt D{px, py, 1};
#pragma unroll 1
        for (int j = 0; j < SB_CHAIN_STEPS - 1; j++) {
            fp_t c = 0, nrm = 0;
            if (warp == 0 && first) {
                S.cx[j][k] = D.X;
                S.cy[j][k] = D.Y;
                if (k == 0) S.cw[j] = D.w;
            }
            asm volatile("bar.sync 1, 64;" ::: "memory");
            if (warp == 3) {
                fp_t xx = dfp6_mul<F>(mask, S.cx[j][k], S.cx[j][k], k, gbase);
                if (first) S.xx[k] = xx;
            } else {
                c = dfp6_cofactor_norm<F>(mask, D.Y, k, gbase, &nrm);   // 1 / (2 Y) = c / (2 n)
            }
            asm volatile("bar.sync 2, 64;" ::: "memory");
            if (warp == 0) {
                fp_t X = D.X, Y = D.Y, w = D.w;
                fp_t m = fp_add(nrm, nrm);
                fp_t w4 = fp_sqr(fp_sqr_nc(w));
                fp_t xx = S.xx[k];
                fp_t num = fp_add(fp_dbl(xx), xx);
                num = fp_add(num, k == 0 ? w4 : 0);                     // 3 X^2 + a w^4, a = 1
                fp_t L = dfp6_mul<F>(mask, num, c, k, gbase);           // slope = L / (m w)
                fp_t m2 = fp_sqr_nc(m), m3 = fp_mul_nc(m2, m);
                fp_t A = fp_mul(X, m2);
                fp_t X3 = fp_sub(fp_sub(dfp6_mul<F>(mask, L, L, k, gbase), A), A);
                fp_t Y3 = fp_sub(dfp6_mul<F>(mask, L, fp_sub(A, X3), k, gbase), fp_mul(Y, m3));
                D.X = X3;
                D.Y = Y3;
                D.w = fp_mul(m, w);
                exc |= nrm == 0;                                        // a point of order 2
            }
        }
        if (warp == 0 && first) {
            S.cx[SB_CHAIN_STEPS - 1][k] = D.X;
            S.cy[SB_CHAIN_STEPS - 1][k] = D.Y;
            if (k == 0) S.cw[SB_CHAIN_STEPS - 1] = D.w;
        }
    }
    {
        if (warp == 1) {
            scalar h = sc_zero();
            uint64_t off = msg_off[i];
            if (x_ok) h = dchallenge_scalar<F>(mask, sx, px, py, msgs + off, msg_off[i + 1] - off, k, gbase, S.mds2);
            if (lane == 0) recode_signed_w4(h, S.hd);
        } else if (warp == 2) {
            scalar e = load_scalar_planes(in.planes, 3, n, i);
            dpt R{0, 0, 1}, T;
            T.w = 1;
            bool e_acc = true;   // R still empty
            int carry = 0;
#pragma unroll 1
            for (int w = 0; w < GTAB_WINDOWS; w++) {
                int raw = (int)sc_bits(e, GTAB_W * w, GTAB_W) + carry;
                bool neg = raw > (1 << (GTAB_W - 1));
                carry = neg ? 1 : 0;
                int dg = neg ? (1 << GTAB_W) - raw : raw;
                const uint64_t* ent = gtab + ((size_t)w * GTAB_ENTRIES + (dg ? dg : 1)) * GTAB_ENTRY_U64;
                T.X = ent[k];
                T.Y = ent[6 + k];
                exc |= djf_add<F>(mask, &R, &T, jf_add_mode(e_acc, dg == 0, neg), k, gbase);
                e_acc = e_acc && dg == 0;
            }
            if (first) {
                S.ex[k] = R.X;
                S.ey[k] = R.Y;
            }
            if (lane == 0) {
                S.ew = R.w;
                S.e_empty = e_acc;
            }
        }
    }
    __syncthreads();

    // ---- phase B: sixteen bucket accumulations side by side -------------------------------------------------------
    const int slot = warp * 5 + g;           // 20 groups, the first 16 own a bucket
    const bool owner = !ghost && slot < ONE_SLOTS;
    if (owner && k == 0) {                   // compacted list of the chain steps that belong to this bucket
        int c = 0;
        if (slot < 8) {
#pragma unroll 1
            for (int j = 0; j < CHEETAH_Q_WNAF5_LEN; j++) {
                int dq = SB_QWNAF(j);
                if (dq != 0 && ((dq < 0 ? -dq : dq) >> 1) == slot) {
                    S.lst[slot][c] = (uint8_t)j;
                    S.neg[slot][c] = dq < 0;
                    c++;
                }
            }
        } else {
#pragma unroll 1
            for (int w = 0; w < 64; w++) {
                int dh = S.hd[w];
                if (dh != 0 && (dh < 0 ? -dh : dh) == slot - 7) {
                    S.lst[slot][c] = (uint8_t)(4 * w);
                    S.neg[slot][c] = dh < 0;
                    c++;
                }
            }
        }
        S.cnt[slot] = c;
    }
    __syncwarp();
    {
        int cnt = owner ? S.cnt[slot] : 0;
        int wmax = cnt;                      // longest list among the warp's groups: everybody runs that many (masked) steps
#pragma unroll
        for (int d = 16; d > 0; d >>= 1) wmax = max(wmax, __shfl_xor_sync(0xffffffffu, wmax, d));
        dpt B{0, 0, 1};
#pragma unroll 1
        for (int t = 0; t < wmax; t++) {     // warp-uniform trip count; groups without a t-th entry run a masked step
            bool live = t < cnt;
            int j = live ? S.lst[slot][t] : 0;
            bool neg = live && S.neg[slot][t];
            dpt src = one_load_chain(S, j, k);
            exc |= djf_add<F>(mask, &B, &src, jf_add_mode(t == 0, !live, neg), k, gbase);
        }
        if (owner) {
            S.bx[slot][k] = B.X;
            S.by[slot][k] = B.Y;
            if (k == 0) S.bw[slot] = B.w;
        }
    }
    __syncthreads();

    // ---- phase C: aggregation (R_k = sum_{m>=k} B_m, O_k = sum_{m>=k} R_m) ------------------------------------------
    //   subgroup check (odd digits 2k+1):  [q]P = 2 O_1 + R_0      challenge (digits m = k+1):  h*P = O_0
    uint8_t v = VERDICT_NEEDS_EXACT;
    if (warp == 0) {
        dpt Rq = one_load_bucket(S, 7, k), Oq = Rq;
        bool eRq = S.cnt[7] == 0, eOq = eRq;
#pragma unroll 1
        for (int b = 6; b >= 0; b--) {
            dpt Bb = one_load_bucket(S, b, k);
            bool eb = S.cnt[b] == 0;
            exc |= djf_add<F>(mask, &Rq, &Bb, jf_add_mode(eRq, eb, false), k, gbase);
            eRq = eRq && eb;
            if (b >= 1) {
                exc |= djf_add<F>(mask, &Oq, &Rq, jf_add_mode(eOq, eRq, false), k, gbase);
                eOq = eOq && eRq;
            }
        }
        if (eOq || eRq) exc = true;  // degenerate digit pattern: leave it to the exact routine
        else exc |= djf_dbl<F>(mask, &Oq, k, gbase);
        // [q]P == O  <=>  2 O_1 == -R_0:  X_O w_R^2 == X_R w_O^2  and  Y_O w_R^3 == -Y_R w_O^3
        fp_t wos = fp_sqr_nc(Oq.w), wrs = fp_sqr_nc(Rq.w);
        bool x_eq = dall<F>(mask, gbase, fp_mul(Oq.X, wrs) == fp_mul(Rq.X, wos));
        bool y_opp = dall<F>(mask, gbase, fp_mul(Oq.Y, fp_mul_nc(wrs, Rq.w)) == fp_neg(fp_mul(Rq.Y, fp_mul_nc(wos, Oq.w))));
        bool torsion_free = x_eq && y_opp;
        if (x_eq && !torsion_free) exc = true;  // 2 O_1 == R_0: a doubling the fast path does not evaluate
        if (lane == 0) S.torsion_free = torsion_free;
    }
    dpt R{0, 0, 1};
    bool r_empty = true;
    if (warp == 1) {
        dpt Rh = one_load_bucket(S, 15, k), Oh = Rh;
        bool eRh = S.cnt[15] == 0, eOh = eRh;
        bool same_h = !eRh;  // O_h and R_h are the same (finite) point: O += R is then a doubling
#pragma unroll 1
        for (int b = 6; b >= 0; b--) {
            dpt Bb = one_load_bucket(S, 8 + b, k);
            bool eb = S.cnt[8 + b] == 0;
            exc |= djf_add<F>(mask, &Rh, &Bb, jf_add_mode(eRh, eb, false), k, gbase);
            if (!eb && !eRh) same_h = false;
            eRh = eRh && eb;
            if (same_h && !eOh) {  // group-uniform
                exc |= djf_dbl<F>(mask, &Oh, k, gbase);
                same_h = false;
            } else {
                exc |= djf_add<F>(mask, &Oh, &Rh, jf_add_mode(eOh, eRh, false), k, gbase);
                same_h = eOh && !eRh;
            }
            eOh = eOh && eRh;
        }
        dpt EG{S.ex[k], S.ey[k], S.ew};
        bool e_empty = S.e_empty != 0;
        exc |= djf_add<F>(mask, &Oh, &EG, jf_add_mode(eOh, e_empty, false), k, gbase);
        R = Oh;
        r_empty = eOh && e_empty;
        if (r_empty) exc = true;  // the result is the identity: exact routine
    }
    if (exc) atomicOr(&S.exc, 1);
    __syncthreads();
    if (warp == 1) {
        bool eq = dall<F>(mask, gbase, R.X == fp_mul(sx, fp_sqr_nc(R.w)));  // x(R) == sig.x  <=>  X == sig.x w^2
        v = S.exc ? VERDICT_NEEDS_EXACT
            : (!S.torsion_free ? VERDICT_INVALID_PUBLIC_KEY : (!x_ok ? VERDICT_MALFORMED : (eq ? VERDICT_OK : VERDICT_INVALID_SIGNATURE)));
        if (lane == 0) {
            verdicts[i] = v;
            if (v == VERDICT_NEEDS_EXACT) work_list[atomicAdd(work_count, 1u)] = (uint32_t)i;
        }
    }
}

}  // namespace sb
