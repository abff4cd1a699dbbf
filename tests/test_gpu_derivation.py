"""f4: hierarchical deterministic derivation on the device (HMAC-SHA512 + fixed-base multiplication per child)
against the oracle's restatement of src/derivation.rs:66-277, plus the behavioural cases of the reference's own test
(src/derivation.rs:401-447)."""
import numpy as np
import pytest

import pyref as o

pytestmark = pytest.mark.gpu


def test_master_private_and_public_children_match_the_oracle():
    import schnorr_sig_b200 as s
    eng = s.default_engine(0)
    rng = np.random.default_rng(81)
    seeds = rng.integers(0, 256, (40, 32), dtype=np.uint8)
    xsk, ok = eng.derive_master_keys(seeds)
    for k in range(40):
        sk, chain, some = o.hd_master_key(bytes(seeds[k]))
        assert bool(ok[k]) == some and bytes(xsk[k]) == sk.to_bytes(32, "little") + chain
    sk, chain, _ = o.hd_master_key(bytes(seeds[0]))
    idx = np.concatenate([np.array([0, 1, 2**31 - 1, 2**31, 2**32 - 1], dtype=np.uint32),
                          rng.integers(0, 2**32, 120, dtype=np.uint64).astype(np.uint32)])
    kids, ok = eng.derive_private_children(bytes(xsk[0]), idx)
    for k, i in enumerate(idx):
        ck, cc, some = o.hd_derive_private(sk, chain, int(i))
        assert bool(ok[k]) == some and bytes(kids[k]) == ck.to_bytes(32, "little") + cc
    # public children of the master's public key: first 24 against the big-int oracle (a scalar multiplication each)
    pk = o.pt_mul(o.generator(), sk)
    xpk = bytes(o.compress(pk)) + chain
    pkids, pok = eng.derive_public_children(xpk, idx)
    for k, i in enumerate(idx[:24]):
        child, cc, some = o.hd_derive_public(pk, chain, int(i))
        assert bool(pok[k]) == some
        assert bytes(pkids[k]) == bytes(o.compress(child)) + cc
    assert list(pok[:5]) == [1, 1, 1, 0, 0]                      # hardened indices cannot be derived from a public key


def test_derivation_behaviour_like_the_reference_tests():
    """src/derivation.rs:401-447: private and public derivation paths of a non-hardened child agree; wrong-kind indices
    give None; encodings round-trip."""
    import schnorr_sig_b200 as s
    rng = np.random.default_rng(82)
    skey = s.ExtendedPrivateKey.generate_master_key(bytes(rng.integers(0, 256, 32, dtype=np.uint8)))
    assert skey is not None
    pkey = s.ExtendedPublicKey.from_extended_private_key(skey)
    indices = [int(v) & 0x7FFFFFFF for v in rng.integers(0, 2**32, 100, dtype=np.uint64)]
    priv = skey.derive_private_many(indices)
    pub = pkey.derive_normal_public_many(indices)
    for a, b in list(zip(priv, pub))[:100]:
        assert a is not None and b is not None and b.chaincode == a.chaincode
    # public key of each private child on the device in one call
    pk96, inf = s.default_engine().keygen(np.stack([np.frombuffer(a.key.to_bytes(), dtype=np.uint8) for a in priv]))
    for k, b in enumerate(pub):
        assert b.key.xy == bytes(pk96[k]) and not inf[k]
    i = (0xFFFFFFFF & 0x7FFFFFFF).to_bytes(4, "little")           # 2^31 - 1
    assert skey.derive_hardened_private(i) is None
    assert skey.derive_normal_private(b"\x00\x00\x00\x80") is None
    assert pkey.derive_normal_public(b"\x00\x00\x00\x80") is None
    assert skey.derive_private(b"\x00\x00\x00\x80") is not None and skey.derive_public(7) == pkey.derive_normal_public(7)
    assert s.ExtendedPrivateKey.from_bytes(skey.to_bytes()) == skey
    assert s.ExtendedPublicKey.from_bytes(pkey.to_bytes()) == pkey
    one = s.ExtendedPrivateKey(s.PrivateKey(1), s.ChainCode(bytes([1] * 32)))
    assert one.to_bytes() == bytes([1] + [0] * 31 + [1] * 32)     # src/derivation.rs:451-462
    assert s.ExtendedPrivateKey.from_bytes(bytes(64)) is None      # zero key
    # signatures made with a derived key verify under the publicly derived key
    child = priv[3]
    kp = s.KeyPair(child.key, pub[3].key)
    assert kp.sign(b"derived").verify(b"derived", pub[3].key).is_ok()
