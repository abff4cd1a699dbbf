#!/usr/bin/env python3
"""Generates tests/golden/vectors_v1.json from oracle #1 (oracle/pyref.py, pure big-int).

The reference cannot run here (Rust, un-vendored deps) and holds no known answers of its own, so
these fixtures freeze the RESTATED oracle: they catch drift between oracle #1, oracle #2 and the
CUDA engine, and are the vectors to regenerate once true upstream constants are substituted
(INTEGRATION.md §5).  Run:  python tools/gen_golden.py
"""
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "oracle"))
import pyref as o  # noqa: E402

SEED = 0x5343484E4F525231


def main():
    lens = [0, 1, 6, 7, 8, 13, 14, 24, 48, 80, 160, 163]
    items = []
    for i, L in enumerate(lens):
        sk = o.synth_scalar(SEED, "sk", i)
        r = o.synth_scalar(SEED, "nonce", i)
        msg = o.prf(SEED, "msg", i, L)
        pk = o.public_key(sk)
        x49, e = o.sign(sk, pk, msg, r)
        rx = o.f6_from_bytes(x49[:48])
        items.append({
            "sk": sk.to_bytes(32, "little").hex(), "nonce": r.to_bytes(32, "little").hex(), "msg": msg.hex(),
            "pk": (o.f6_to_bytes(pk[0]) + o.f6_to_bytes(pk[1])).hex(), "pk_compressed": o.compress(pk).hex(),
            "sig": (x49 + e.to_bytes(32, "little")).hex(),
            "digest": o.hash_message(rx, pk, msg).hex(),
            "verdict": o.verify(x49, e, msg, pk),
            "rand": o.synth_scalar(SEED, "rand", i).to_bytes(32, "little").hex(),
        })
        assert items[-1]["verdict"] == 0
    sigs = [(bytes.fromhex(it["sig"])[:49], int.from_bytes(bytes.fromhex(it["sig"])[49:], "little")) for it in items]
    pks = [o.public_key(int.from_bytes(bytes.fromhex(it["sk"]), "little")) for it in items]
    msgs = [bytes.fromhex(it["msg"]) for it in items]
    rand = [int.from_bytes(bytes.fromhex(it["rand"]), "little") for it in items]
    v, lhs, rhs = o.verify_batch(sigs, pks, msgs, rand)
    assert v == 0
    swapped = list(pks); swapped[1], swapped[2] = swapped[2], swapped[1]
    v2, lhs2, _ = o.verify_batch(sigs, swapped, msgs, rand)
    assert v2 == 2
    kat = (o.KAT_X, o.KAT_Y)
    out = {
        "_about": "frozen outputs of oracle/pyref.py (restated reference; parity unpinned, DESIGN.md §3); made by tools/gen_golden.py",
        "seed": hex(SEED), "items": items,
        "batch": {"verdict": v, "lhs": (o.f6_to_bytes(lhs[0]) + o.f6_to_bytes(lhs[1])).hex(),
                  "rhs": (o.f6_to_bytes(rhs[0]) + o.f6_to_bytes(rhs[1])).hex(),
                  "swapped_1_2_verdict": v2, "swapped_lhs": (o.f6_to_bytes(lhs2[0]) + o.f6_to_bytes(lhs2[1])).hex()},
        "negative": {
            "off_subgroup_key": (o.f6_to_bytes(kat[0]) + o.f6_to_bytes(kat[1])).hex(),   # src/signature.rs:387-404
            "off_subgroup_verdict": o.verify(sigs[4][0], sigs[4][1], msgs[4], kat),
            "wrong_message_verdict": o.verify(sigs[4][0], sigs[4][1], b"\x2a" + msgs[4][1:], pks[4]),
            "identity_x_verdict": o.verify(bytes(48) + b"\x80", sigs[4][1], msgs[4], pks[4]),
            "zero_e_verdict": o.verify(sigs[4][0], 0, msgs[4], pks[4]),
        },
        "rescue_permutation_of_0_to_11": [hex(c) for c in o.rescue_permutation(list(range(12)))],
        "generator_times_q_minus_1": (lambda p: (o.f6_to_bytes(p[0]) + o.f6_to_bytes(p[1])).hex())(o.pt_mul(o.generator(), o.Q - 1)),
    }
    os.makedirs(os.path.join(ROOT, "tests", "golden"), exist_ok=True)
    with open(os.path.join(ROOT, "tests", "golden", "vectors_v1.json"), "w") as f:
        json.dump(out, f, indent=1)
        f.write("\n")
    print("wrote tests/golden/vectors_v1.json (%d items)" % len(items))


if __name__ == "__main__":
    main()
