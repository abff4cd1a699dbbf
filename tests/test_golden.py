"""Committed golden fixtures (tests/golden/vectors_v1.json, made by tools/gen_golden.py from the
big-int oracle): oracle #2 on CPU, the CUDA engine on GPU."""
import json
import os

import numpy as np
import pytest

import cref
from util import pack_msgs

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def load():
    with open(os.path.join(ROOT, "tests", "golden", "vectors_v1.json")) as f:
        g = json.load(f)
    it = g["items"]
    arr = lambda k, w: np.frombuffer(b"".join(bytes.fromhex(i[k]) for i in it), dtype=np.uint8).reshape(len(it), w).copy()
    msgs = [bytes.fromhex(i["msg"]) for i in it]
    blob, off = pack_msgs(msgs)
    return g, dict(sk=arr("sk", 32), nonce=arr("nonce", 32), pk=arr("pk", 96), sigs=arr("sig", 81), rand=arr("rand", 32),
                   digest=arr("digest", 32), comp=arr("pk_compressed", 49), msgs=msgs, blob=blob, off=off,
                   inf=np.zeros(len(it), np.uint8))


def check(impl, g, w):
    pk, inf = impl.keygen(w["sk"])
    assert np.array_equal(pk, w["pk"]) and not inf.any()
    assert np.array_equal(impl.sign_many(w["sk"], pk, inf, w["blob"], w["off"], w["nonce"]), w["sigs"])
    assert np.array_equal(impl.hash_messages(w["sigs"][:, :48].copy(), w["pk"], w["blob"], w["off"]), w["digest"])
    assert list(impl.verify_many(w["sigs"], w["pk"], w["inf"], w["blob"], w["off"])) == [i["verdict"] for i in g["items"]]
    v, lhs, rhs = impl.verify_batch(w["sigs"], w["pk"], w["inf"], w["blob"], w["off"], w["rand"])
    assert v == g["batch"]["verdict"] and bytes(lhs[:96]).hex() == g["batch"]["lhs"] and bytes(rhs[:96]).hex() == g["batch"]["rhs"]
    pk2 = w["pk"].copy(); pk2[[1, 2]] = pk2[[2, 1]]
    v, lhs, _ = impl.verify_batch(w["sigs"], pk2, w["inf"], w["blob"], w["off"], w["rand"])
    assert v == g["batch"]["swapped_1_2_verdict"] and bytes(lhs[:96]).hex() == g["batch"]["swapped_lhs"]
    neg = g["negative"]
    one = lambda sig, pkk, m: int(impl.verify_many(sig.reshape(1, 81), pkk.reshape(1, 96), np.zeros(1, np.uint8), *pack_msgs([m]))[0])
    kat = np.frombuffer(bytes.fromhex(neg["off_subgroup_key"]), dtype=np.uint8)
    assert one(w["sigs"][4], kat, w["msgs"][4]) == neg["off_subgroup_verdict"] == 1
    assert one(w["sigs"][4], w["pk"][4], b"\x2a" + w["msgs"][4][1:]) == neg["wrong_message_verdict"] == 2
    s = w["sigs"][4].copy(); s[:48] = 0; s[48] = 0x80
    assert one(s, w["pk"][4], w["msgs"][4]) == neg["identity_x_verdict"] == 2
    s = w["sigs"][4].copy(); s[49:] = 0
    assert one(s, w["pk"][4], w["msgs"][4]) == neg["zero_e_verdict"] == 2


class CrefImpl:
    keygen = staticmethod(lambda sk: cref.keygen(sk, 2))
    sign_many = staticmethod(lambda *a: cref.sign_many(*a, 2))
    hash_messages = staticmethod(lambda *a: cref.hash_messages(*a, 2))
    verify_many = staticmethod(lambda *a: cref.verify_many(*a, 2))
    verify_batch = staticmethod(lambda *a: cref.verify_batch(*a, 2))


def test_c_oracle_against_golden():
    g, w = load()
    check(CrefImpl, g, w)
    assert [hex(int(c)) for c in cref.rescue_permutation(list(range(12)))] == g["rescue_permutation_of_0_to_11"]
    for i in range(len(g["items"])):
        assert bytes(cref.compress(w["pk"][i])) == bytes(w["comp"][i])


@pytest.mark.gpu
def test_cuda_engine_against_golden():
    import schnorr_sig_b200 as s
    g, w = load()
    eng = s.default_engine(0)
    check(eng, g, w)
    assert np.array_equal(eng.compress(w["pk"], w["inf"]), w["comp"])
    q1 = np.frombuffer(((int(s.api._Q) - 1).to_bytes(32, "little")), dtype=np.uint8).reshape(1, 32)
    assert bytes(eng.keygen(q1)[0][0]).hex() == g["generator_times_q_minus_1"]
