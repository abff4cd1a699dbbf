"""Oracle #1 — pure-Python big-integer restatement of the schnorr-sig hot path.

TEST INFRASTRUCTURE ONLY. Nothing under oracle/ is imported, linked or executed by the
product path (schnorr-sig_b200/); only tests/, __graft_entry__.smoke() and bench.py's
cpu_baseline / --impl reference legs may use it, and only as the checker.

PARITY STATUS: **parity unpinned** for value-level results.  The reference crate
(/root/reference) contains no field/curve/hash arithmetic: it calls the un-vendored,
un-pinned git dependencies `cheetah` and `hash` (Cargo.toml:16,18; no Cargo.lock), whose
sources are not on this box, and its tests hold no digest/signature known answers.
What IS pinned by the reference and checked in tests/test_oracle_pins.py:
  * p, the curve equation and the Fp6 modulus            (README.md:4-8)
  * the off-subgroup KAT point                            (src/signature.rs:385-406, src/error.rs:47-64)
  * identity / scalar encodings, byte lengths             (src/public.rs:94-101, src/constants.rs:12-46)
  * every behavioural accept/reject case of the reference's own tests
Everything tagged PLACEHOLDER / SPEC-DERIVED / RECALLED in params.py is shared by the oracle
and the CUDA engine, so GPU-vs-oracle parity is like-for-like.

The protocol glue follows the reference line by line:
  hash_message          src/signature.rs:274-306
  sign                  src/signature.rs:114-129
  Signature::verify     src/signature.rs:181-205
  verify_batch          src/batch.rs:31-130
  encodings             src/signature.rs:208-227, src/public.rs:49-56
Arithmetic (cheetah/hash are absent) follows the published definitions: Goldilocks field,
Fp6 = Fp[u]/(u^6-7), short-Weierstrass group law, Rescue-Prime (ePrint 2020/1143).
"""
from __future__ import annotations

import hashlib

P = 2**64 - 2**32 + 1                       # README.md:4
Q = 0x7AF2599B3B3F22D0563FBF0F990A37B5327AA72330157722D443623EAED4ACCF   # SURVEY App. A [VERIFIED]
COFACTOR = 708537115134665106932687062569690615370                      # SURVEY App. A [VERIFIED]
NONRES = 7                                   # u^6 = 7, README.md:8

# off-subgroup known-answer point, src/signature.rs:387-404
KAT_X = (0x9BFCD3244AFCB637, 0x39005E478830B187, 0x7046F1C03B42C6CC,
         0xB5EEAC99193711E5, 0x7FD272E724307B98, 0xCC371DD6DD5D8625)
KAT_Y = (0x9D03FDC216DFAAE8, 0xBF4ADE2A7665D9B8, 0xF08B022D5B3262B7,
         0x2EAF583A3CF15C6F, 0xA92531E4B1338285, 0x5B8157814141A7A7)

# ----------------------------------------------------------------------------- Fp6
F6_ZERO = (0, 0, 0, 0, 0, 0)
F6_ONE = (1, 0, 0, 0, 0, 0)
CURVE_A = F6_ONE                              # y^2 = x^3 + x + B, README.md:4-5
CURVE_B = (395, 1, 0, 0, 0, 0)                # B = u + 395


def f6_add(a, b):
    return tuple((x + y) % P for x, y in zip(a, b))


def f6_sub(a, b):
    return tuple((x - y) % P for x, y in zip(a, b))


def f6_neg(a):
    return tuple((-x) % P for x in a)


def f6_mul(a, b):
    t = [0] * 11
    for i in range(6):
        ai = a[i]
        if ai:
            for j in range(6):
                t[i + j] += ai * b[j]
    return tuple((t[k] + NONRES * t[k + 6]) % P if k < 5 else t[k] % P for k in range(6))


def f6_sqr(a):
    return f6_mul(a, a)


def f6_scalar(a, k):
    return tuple((x * k) % P for x in a)


def f6_pow(a, e):
    r = F6_ONE
    base = a
    while e:
        if e & 1:
            r = f6_mul(r, base)
        base = f6_sqr(base)
        e >>= 1
    return r


def f6_inv_fermat(a):
    """a^(p^6-2); slow, used to cross-check f6_inv."""
    return f6_pow(a, P**6 - 2)


def f6_inv(a):
    """Tower inversion Fp6 = Fp3[u]/(u^2 - v), Fp3 = Fp[v]/(v^3 - 7); checked against Fermat in tests."""
    a0 = (a[0], a[2], a[4])
    a1 = (a[1], a[3], a[5])
    # d = a0^2 - v*a1^2  in Fp3
    s0 = _f3_sqr(a0)
    s1 = _f3_mulv(_f3_sqr(a1))
    d = tuple((x - y) % P for x, y in zip(s0, s1))
    di = _f3_inv(d)
    r0 = _f3_mul(a0, di)
    r1 = _f3_mul(tuple((-x) % P for x in a1), di)
    return (r0[0], r1[0], r0[1], r1[1], r0[2], r1[2])


def _f3_mul(a, b):
    a0, a1, a2 = a
    b0, b1, b2 = b
    return ((a0 * b0 + 7 * (a1 * b2 + a2 * b1)) % P,
            (a0 * b1 + a1 * b0 + 7 * a2 * b2) % P,
            (a0 * b2 + a1 * b1 + a2 * b0) % P)


def _f3_sqr(a):
    return _f3_mul(a, a)


def _f3_mulv(a):
    return ((7 * a[2]) % P, a[0], a[1])


def _f3_inv(d):
    d0, d1, d2 = d
    t0 = (d0 * d0 - 7 * d1 * d2) % P
    t1 = (7 * d2 * d2 - d0 * d1) % P
    t2 = (d1 * d1 - d0 * d2) % P
    n = (d0 * t0 + 7 * (d2 * t1 + d1 * t2)) % P
    ni = pow(n, P - 2, P)
    return ((t0 * ni) % P, (t1 * ni) % P, (t2 * ni) % P)


def f6_is_square(a):
    return a == F6_ZERO or f6_pow(a, (P**6 - 1) // 2) == F6_ONE


_TS = {}


def _ts_setup():
    """Tonelli-Shanks constants in Fp6: p^6 - 1 = 2^33 * t (2-adicity 33)."""
    if _TS:
        return _TS
    n = P**6 - 1
    s = 0
    while n % 2 == 0:
        n //= 2
        s += 1
    assert s == 33
    # deterministic non-residue search over small elements c + u
    c = 0
    while True:
        z = (c, 1, 0, 0, 0, 0)
        if not f6_is_square(z):
            break
        c += 1
    _TS.update(s=s, t=n, z=f6_pow(z, n))
    return _TS


def f6_sqrt(a):
    """Generic Tonelli-Shanks square root in Fp6 (returns one root or None)."""
    if a == F6_ZERO:
        return F6_ZERO
    ts = _ts_setup()
    s, t = ts["s"], ts["t"]
    x = f6_pow(a, (t + 1) // 2)
    b = f6_pow(a, t)
    g = ts["z"]
    r = s
    while b != F6_ONE:
        m = 0
        bb = b
        while bb != F6_ONE:
            bb = f6_sqr(bb)
            m += 1
            if m == r:
                return None
        gs = g
        for _ in range(r - m - 1):
            gs = f6_sqr(gs)
        g = f6_sqr(gs)
        x = f6_mul(x, gs)
        b = f6_mul(b, g)
        r = m
    return x


def fp_lex_largest(c):
    return c > (P - 1) // 2


def f6_lex_largest(a):
    """'lexicographically largest' of {a, -a}: decided by the highest non-zero coefficient
    [RECALLED convention of cheetah's compressed encoding; unpinned, see params.py]."""
    for c in reversed(a):
        if c != 0:
            return fp_lex_largest(c)
    return False


def f6_to_bytes(a):
    return b"".join(int(c).to_bytes(8, "little") for c in a)


def f6_from_bytes(b):
    """Fp6::from_bytes: None (reference: CtOption none -> unwrap panics, src/signature.rs:186)
    when a limb is not canonical."""
    assert len(b) == 48
    cs = tuple(int.from_bytes(b[8 * i:8 * i + 8], "little") for i in range(6))
    if any(c >= P for c in cs):
        return None
    return cs


# ----------------------------------------------------------------------------- curve (Jacobian)
INF = None            # affine identity


def on_curve(pt):
    if pt is INF:
        return True
    x, y = pt
    return f6_sqr(y) == f6_add(f6_add(f6_mul(f6_sqr(x), x), x), CURVE_B)


def _jac(pt):
    return (F6_ONE, F6_ONE, F6_ZERO) if pt is INF else (pt[0], pt[1], F6_ONE)


def _jac_to_affine(j):
    X, Y, Z = j
    if Z == F6_ZERO:
        return INF
    zi = f6_inv(Z)
    zi2 = f6_sqr(zi)
    return (f6_mul(X, zi2), f6_mul(Y, f6_mul(zi2, zi)))


def _jac_dbl(j):
    X, Y, Z = j
    if Z == F6_ZERO or Y == F6_ZERO:
        return (F6_ONE, F6_ONE, F6_ZERO)
    YY = f6_sqr(Y)
    S = f6_scalar(f6_mul(X, YY), 4)
    ZZ = f6_sqr(Z)
    M = f6_add(f6_scalar(f6_sqr(X), 3), f6_sqr(ZZ))        # a = 1
    X3 = f6_sub(f6_sqr(M), f6_scalar(S, 2))
    Y3 = f6_sub(f6_mul(M, f6_sub(S, X3)), f6_scalar(f6_sqr(YY), 8))
    Z3 = f6_scalar(f6_mul(Y, Z), 2)
    return (X3, Y3, Z3)


def _jac_add(j1, j2):
    X1, Y1, Z1 = j1
    X2, Y2, Z2 = j2
    if Z1 == F6_ZERO:
        return j2
    if Z2 == F6_ZERO:
        return j1
    Z1Z1 = f6_sqr(Z1)
    Z2Z2 = f6_sqr(Z2)
    U1 = f6_mul(X1, Z2Z2)
    U2 = f6_mul(X2, Z1Z1)
    S1 = f6_mul(Y1, f6_mul(Z2, Z2Z2))
    S2 = f6_mul(Y2, f6_mul(Z1, Z1Z1))
    if U1 == U2:
        if S1 == S2:
            return _jac_dbl(j1)
        return (F6_ONE, F6_ONE, F6_ZERO)
    H = f6_sub(U2, U1)
    R = f6_sub(S2, S1)
    HH = f6_sqr(H)
    HHH = f6_mul(H, HH)
    V = f6_mul(U1, HH)
    X3 = f6_sub(f6_sub(f6_sqr(R), HHH), f6_scalar(V, 2))
    Y3 = f6_sub(f6_mul(R, f6_sub(V, X3)), f6_mul(S1, HHH))
    Z3 = f6_mul(f6_mul(Z1, Z2), H)
    return (X3, Y3, Z3)


def pt_neg(pt):
    return INF if pt is INF else (pt[0], f6_neg(pt[1]))


def pt_add(p1, p2):
    return _jac_to_affine(_jac_add(_jac(p1), _jac(p2)))


def pt_mul(pt, k):
    """Plain left-to-right double-and-add, k >= 0 (any size)."""
    acc = (F6_ONE, F6_ONE, F6_ZERO)
    base = _jac(pt)
    for bit in bin(k)[2:] if k else "":
        acc = _jac_dbl(acc)
        if bit == "1":
            acc = _jac_add(acc, base)
    return _jac_to_affine(acc)


def pt_mul2(p1, k1, p2, k2):
    """k1*p1 + k2*p2 (Shamir, bitwise)."""
    acc = (F6_ONE, F6_ONE, F6_ZERO)
    j1, j2 = _jac(p1), _jac(p2)
    j12 = _jac_add(j1, j2)
    for i in reversed(range(max(k1.bit_length(), k2.bit_length()))):
        acc = _jac_dbl(acc)
        b1, b2 = (k1 >> i) & 1, (k2 >> i) & 1
        if b1 and b2:
            acc = _jac_add(acc, j12)
        elif b1:
            acc = _jac_add(acc, j1)
        elif b2:
            acc = _jac_add(acc, j2)
    return _jac_to_affine(acc)


def is_torsion_free(pt):
    """AffinePoint::is_torsion_free (call site src/signature.rs:182): [q]P == O."""
    return pt_mul(pt, Q) is INF


def compress(pt):
    """CompressedPoint: 48 bytes of x, then a flag byte: bit 7 = infinity (src/public.rs:95-101),
    bit 6 = y is the lexicographically largest root [RECALLED, unpinned]."""
    if pt is INF:
        return bytes(48) + b"\x80"
    x, y = pt
    return f6_to_bytes(x) + bytes([0x40 if f6_lex_largest(y) else 0x00])


def decompress(b):
    """AffinePoint::from_compressed; returns (ok, point).  Rejections follow the reference tests
    (src/public.rs:114-129,150-151): non-canonical limbs, x not on the curve, unknown flag bits,
    infinity flag with non-zero x / sign bit."""
    assert len(b) == 49
    flags = b[48]
    if flags & 0x3F:
        return False, INF
    inf = bool(flags & 0x80)
    sign = bool(flags & 0x40)
    x = f6_from_bytes(b[:48])
    if x is None:
        return False, INF
    if inf:
        if x != F6_ZERO or sign:
            return False, INF
        return True, INF
    rhs = f6_add(f6_add(f6_mul(f6_sqr(x), x), x), CURVE_B)
    y = f6_sqrt(rhs)
    if y is None:
        return False, INF
    if f6_lex_largest(y) != sign:
        y = f6_neg(y)
    return True, (x, y)


# ----------------------------------------------------------------------------- parameters
def upstream_dump():
    """params/upstream_dump.json = output of rust/dump_params run against the REAL cheetah / hash crates (absent on this
    box: no cargo, no network).  When it exists the generator comes from it and tests/test_oracle_pins.py checks every
    dumped known answer; until then the oracle is "parity unpinned" at value level."""
    global _DUMP
    try:
        return _DUMP
    except NameError:
        import json
        import os
        path = os.path.join(os.path.dirname(os.path.abspath(__file__)), "..", "params", "upstream_dump.json")
        _DUMP = json.load(open(path)) if os.path.exists(path) else None
        return _DUMP


def _limbs_from_hex(h):
    b = bytes.fromhex(h)
    return tuple(int.from_bytes(b[8 * i:8 * i + 8], "little") for i in range(6))


def generator():
    """The generator of the prime-order subgroup.  Upstream value when params/upstream_dump.json exists; otherwise the
    PLACEHOLDER G := [cofactor] * KAT point, reproducible from reference data alone and of order q."""
    global _G
    try:
        return _G
    except NameError:
        d = upstream_dump()
        if d is not None:
            _G = (_limbs_from_hex(d["generator"]["x"]), _limbs_from_hex(d["generator"]["y"]))
            assert on_curve(_G) and pt_mul(_G, Q) is INF
        else:
            _G = pt_mul((KAT_X, KAT_Y), COFACTOR)
        return _G


RESCUE_WIDTH = 12
RESCUE_RATE = 8
RESCUE_ROUNDS = 7            # [RECALLED] upstream `rescue_64_12_8`; unpinned
ALPHA = 7
INV_ALPHA = 10540996611094048183    # 7^-1 mod (p-1)  [VERIFIED]
MDS_FIRST_ROW = (7, 23, 8, 26, 13, 10, 9, 7, 6, 22, 21, 8)   # [RECALLED] circulant; unpinned


def rescue_mds():
    """Circulant matrix: row i is the first row rotated right by i."""
    return [[MDS_FIRST_ROW[(j - i) % 12] for j in range(12)] for i in range(12)]


def rescue_round_constants():
    """SPEC-DERIVED (ePrint 2020/1143, Rescue-Prime reference generator): SHAKE256 of
    "Rescue-XLIX(p,m,capacity,security)" cut in (ceil(bits/8)+1)-byte little-endian chunks mod p.
    Returned as 2*N rows of 12: row 2r is added after the forward S-box half-round of round r,
    row 2r+1 after the inverse S-box half-round."""
    m, cap, sec, n = RESCUE_WIDTH, RESCUE_WIDTH - RESCUE_RATE, 128, RESCUE_ROUNDS
    bytes_per_int = (P.bit_length() + 7) // 8 + 1
    seed = "Rescue-XLIX(%i,%i,%i,%i)" % (P, m, cap, sec)
    stream = hashlib.shake_256(seed.encode("ascii")).digest(bytes_per_int * 2 * m * n)
    cs = [int.from_bytes(stream[bytes_per_int * i:bytes_per_int * (i + 1)], "little") % P
          for i in range(2 * m * n)]
    return [cs[12 * r:12 * r + 12] for r in range(2 * n)]


_MDS = rescue_mds()
_ARK = rescue_round_constants()


def rescue_permutation(state):
    s = list(state)
    for r in range(RESCUE_ROUNDS):
        s = [pow(x, ALPHA, P) for x in s]
        s = [(sum(_MDS[i][j] * s[j] for j in range(12)) + _ARK[2 * r][i]) % P for i in range(12)]
        s = [pow(x, INV_ALPHA, P) for x in s]
        s = [(sum(_MDS[i][j] * s[j] for j in range(12)) + _ARK[2 * r + 1][i]) % P for i in range(12)]
    return s


def rescue_hash_field(elems):
    """RescueHash::hash_field (call site src/signature.rs:303): sponge, rate = state[0..8],
    additive absorption, '1' padding element only when the last block is partial, digest =
    state[0..4]  [RECALLED absorb/padding rule; unpinned]."""
    s = [0] * 12
    i = 0
    for e in elems:
        s[i] = (s[i] + e) % P
        i += 1
        if i == RESCUE_RATE:
            s = rescue_permutation(s)
            i = 0
    if i > 0:
        s[i] = (s[i] + 1) % P
        s = rescue_permutation(s)
    return s[:4]


def digest_to_bytes(d):
    return b"".join(int(c).to_bytes(8, "little") for c in d)


# ----------------------------------------------------------------------------- protocol
def message_to_felts(message: bytes):
    """src/signature.rs:284-301 — 7-byte little-endian chunks; a short tail gets a 0x01 marker."""
    out = []
    nb = len(message) // 7
    for i in range(0, len(message), 7):
        chunk = message[i:i + 7]
        if i // 7 < nb:
            out.append(int.from_bytes(chunk + b"\x00", "little"))
        else:
            buf = bytearray(8)
            buf[:len(chunk)] = chunk
            buf[len(chunk)] = 1
            out.append(int.from_bytes(bytes(buf), "little"))
    return out


def hash_message(rx, pk, message: bytes) -> bytes:
    """src/signature.rs:274-306.  pk is an affine point (x, y); the identity hashes as x=y=0."""
    px, py = (F6_ZERO, F6_ZERO) if pk is INF else pk
    data = list(rx) + list(px) + [py[0]] + message_to_felts(message)
    return digest_to_bytes(rescue_hash_field(data))


def scalar_from_digest(h: bytes) -> int:
    """Scalar::from_bits(_vartime) over the 256 Lsb0 bits of the digest = LE integer mod q."""
    return int.from_bytes(h, "little") % Q


def public_key(sk: int):
    return pt_mul(generator(), sk % Q)


def sign(sk: int, pk, message: bytes, r: int):
    """KeyPair::sign src/signature.rs:114-129 with the nonce r supplied by the caller.
    Returns (compressed R: 49 bytes, e)."""
    R = pt_mul(generator(), r % Q)
    rx = F6_ZERO if R is INF else R[0]
    h = scalar_from_digest(hash_message(rx, pk, message))
    e = (r - sk * h) % Q
    return compress(R), e


OK, INVALID_PUBLIC_KEY, INVALID_SIGNATURE, MALFORMED = 0, 1, 2, 3


def verify(sig_x49: bytes, e: int, message: bytes, pk) -> int:
    """Signature::verify src/signature.rs:181-205 -> verdict code (0 ok, 1 InvalidPublicKey,
    2 InvalidSignature, 3 = the reference panics: non-canonical x limb, :186)."""
    if e >= Q or (pk is not INF and any(c >= P for c in pk[0] + pk[1])):
        return MALFORMED                     # states the reference's Scalar / Fp types cannot hold
    if not is_torsion_free(pk):
        return INVALID_PUBLIC_KEY
    x = f6_from_bytes(sig_x49[:48])          # flag byte ignored
    if x is None:
        return MALFORMED
    h = scalar_from_digest(hash_message(x, pk, message))
    r = pt_mul2(pk, h, generator(), e)
    rx = F6_ZERO if r is INF else r[0]       # identity reads as x = 0
    return OK if rx == x else INVALID_SIGNATURE


def verify_batch(sigs, pks, messages, randomizers):
    """verify_batch src/batch.rs:31-130 with caller-supplied randomisers s_i.
    sigs: list of (x49, e).  Returns (verdict, lhs_affine, rhs_affine)."""
    assert len(sigs) == len(pks) == len(messages) == len(randomizers)
    hashes = []
    for (x49, _e), pk, m in zip(sigs, pks, messages):
        x = f6_from_bytes(x49[:48])
        if x is None:
            return MALFORMED, INF, INF
        hashes.append(scalar_from_digest(hash_message(x, pk, m)))
    lin = sum(s * e for (_x, e), s in zip(sigs, randomizers)) % Q
    rhs = pt_mul(generator(), lin)
    acc = (F6_ONE, F6_ONE, F6_ZERO)
    for (x49, _e), s in zip(sigs, randomizers):
        ok, R = decompress(x49)
        if not ok:
            return MALFORMED, INF, INF
        acc = _jac_add(acc, _jac(pt_mul(R, s % Q)))
    for pk, h, s in zip(pks, hashes, randomizers):
        acc = _jac_add(acc, _jac(pt_mul(pt_neg(pk), (h * s) % Q)))
    lhs = _jac_to_affine(acc)
    lx = F6_ZERO if lhs is INF else lhs[0]
    gx = F6_ZERO if rhs is INF else rhs[0]
    return (OK if lx == gx else INVALID_SIGNATURE), lhs, rhs


# ----------------------------------------------------------------------------- synthetic inputs
def prf(seed: int, domain: str, index: int, nbytes: int) -> bytes:
    """Counter-based PRF of SURVEY §8(d): SHAKE256(seed_le8 || domain || 0x00 || index_le8)."""
    return hashlib.shake_256(seed.to_bytes(8, "little") + domain.encode() + b"\x00"
                             + index.to_bytes(8, "little")).digest(nbytes)


def synth_scalar(seed, domain, index):
    k = int.from_bytes(prf(seed, domain, index, 64), "little") % Q
    return k if k else 1


# ----------------------------------------------------------------------------- HD derivation (src/derivation.rs)
def _hmac_sha512(key: bytes, data: bytes) -> bytes:
    import hashlib
    import hmac
    return hmac.new(key, data, hashlib.sha512).digest()


def hd_master_key(seed32: bytes):
    """ExtendedPrivateKey::generate_master_key (src/derivation.rs:66-84) -> (sk, chain code, is_some)."""
    I = _hmac_sha512(b"Cheetah - Master extended key seed", seed32)
    sk = int.from_bytes(I[:32], "little") % Q          # Scalar::from_bytes_non_canonical [INFERRED: value mod q]
    return sk, I[32:], sk != 0


def hd_derive_private(sk: int, chain: bytes, index: int):
    """ExtendedPrivateKey::derive_private (src/derivation.rs:90-153): hardened (index >= 2^31) children hash
    0^17 || sk, normal children hash PublicKey(sk).to_bytes(); child = I_L + sk."""
    i4 = index.to_bytes(4, "little")
    if index >> 31:
        data = bytes(17) + sk.to_bytes(32, "little") + i4
    else:
        data = bytes(compress(pt_mul(generator(), sk))) + i4
    I = _hmac_sha512(chain, data)
    child = (int.from_bytes(I[:32], "little") + sk) % Q
    return child, I[32:], child != 0


def hd_derive_public(pk, chain: bytes, index: int):
    """ExtendedPublicKey::derive_normal_public (src/derivation.rs:249-277): child = I_L G + pk; None for a hardened
    index or an identity tweak point."""
    I = _hmac_sha512(chain, bytes(compress(pk)) + index.to_bytes(4, "little"))
    t = int.from_bytes(I[:32], "little") % Q
    point = pt_mul(generator(), t)
    return pt_add(point, pk), I[32:], (point is not INF) and (index >> 31) == 0
