#!/usr/bin/env python3
"""Generate tests/golden/*.json by RUNNING THE REFERENCE's own arithmetic (BLST, oracle/_ref, built from
/root/reference/vendor/blst) on the scenarios of /root/reference/tests/t_batch_verifier.nim.  Run in the build
container (where /root/reference exists); the fixtures travel to the GPU box.  python tools/gen_golden.py"""
import hashlib
import json
import os
import random
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from oracle import blst_ref as br  # noqa: E402

GOLD = os.path.join(ROOT, "tests", "golden")
srb = hashlib.sha256(b"Mr F was here").digest()          # t_batch_verifier.nim:60


def forged_pair(seed1, m1, seed2, m2):
    a, b = br.make_set(seed1, m1), br.make_set(seed2, m2)
    p = br.make_set(seed1 * seed2 + seed1 + seed2, b"rekt")
    sp = p[128:]
    _, f1 = br.aggregate_g2(a[128:] + sp)
    _, f2 = br.aggregate_g2(b[128:] + br.g2_neg(sp))
    return a[:128] + f1 + b[:128] + f2


def scenario(name, sets, chunks_list=(0, 4)):
    out = []
    for chunks in chunks_list:
        ok, gt = br.batch_verify(sets, srb, chunks)
        out.append({"name": name, "chunks": chunks, "n": len(sets) // 320, "sets": sets.hex(), "ok": ok,
                    "gt": gt.hex(), "scalars": [str(x) for x in br.rlc_scalars(srb, len(sets) // 320, chunks)]})
    return out


def main():
    sc = []
    for n in (1, 2, 15, 16, 17):
        sc += scenario(f"valid_{n}", br.make_sets(0, n, b"msg"))
    s1, s2 = br.make_set(1, b"msg1"), br.make_set(2, b"msg2")
    sc += scenario("wrong_signature", s1 + s2[:128] + s1[128:])
    sc += scenario("forged_pair", forged_pair(1, b"msg1", 2, b"msg2"))
    many = br.make_sets(0, 16, b"msg") + forged_pair(1, b"msg100", 2, b"msg200")
    items = [many[i:i + 320] for i in range(0, len(many), 320)]
    random.Random(1234).shuffle(items)
    sc += scenario("one_forgery_among_many", b"".join(items))
    inf = bytearray(br.make_sets(0, 4, b"msg"))
    inf[320 + 128:640] = bytes(192)
    sc += scenario("one_infinite_signature", bytes(inf), (0,))
    infpk = bytearray(br.make_sets(0, 3, b"msg"))
    infpk[320:320 + 96] = bytes(96)
    sc += scenario("infinite_pubkey", bytes(infpk), (0,))
    json.dump({"srb": srb.hex(), "source": "BLST e7f90de via oracle/ref_batch.c", "scenarios": sc},
              open(os.path.join(GOLD, "batch_scenarios.json"), "w"))

    dst = b"BLS_SIG_BLS12381G2_XMD:SHA-256_SSWU_RO_POP_"
    msgs = [hashlib.sha256(b"golden%d" % i).digest() for i in range(16)]
    comp, aff = br.hash_to_g2(b"".join(msgs), 32, dst)
    json.dump({"dst": dst.decode(), "msgs": [m.hex() for m in msgs],
               "compressed": [comp[96 * i:96 * i + 96].hex() for i in range(16)],
               "affine": [aff[192 * i:192 * i + 192].hex() for i in range(16)]},
              open(os.path.join(GOLD, "hash_to_g2_eth2.json"), "w"))

    msm = []
    for n in (1, 2, 33, 200):
        pts, scal = br.msm_points(0xFACADE, n)
        msm.append({"n": n, "nbits": 255, "points": pts.hex(), "scalars": scal.hex(), "result": br.msm_g1(pts, scal, 255).hex()})
    json.dump({"cases": msm}, open(os.path.join(GOLD, "msm_g1.json"), "w"))

    sets = br.make_sets(500, 12)
    pks = b"".join(sets[i:i + 96] for i in range(0, len(sets), 320))
    sigs = b"".join(sets[i + 128:i + 320] for i in range(0, len(sets), 320))
    agg_pk, agg_sig = br.aggregate_g1(pks)[1], br.aggregate_g2(sigs)[1]
    # subtractAll (blst_min_pubkey_sig_core.nim:197-209): the aggregate minus its first five members
    json.dump({"pubkeys": pks.hex(), "agg_pubkey": agg_pk.hex(),
               "signatures": sigs.hex(), "agg_signature": agg_sig.hex(),
               "sub5_pubkey": br.subtract_all(agg_pk, pks[:5 * 96]).hex(),
               "sub5_signature": br.subtract_all(agg_sig, sigs[:5 * 192]).hex()},
              open(os.path.join(GOLD, "aggregate.json"), "w"))
    for f in sorted(os.listdir(GOLD)):
        print(f, os.path.getsize(os.path.join(GOLD, f)))


if __name__ == "__main__":
    main()
