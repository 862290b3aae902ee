"""GPU parity of the rows SURVEY.md §8(f) marks next: N2 batched fromBytes with subgroup checks, N3 aggregateVerify /
fastAggregateVerify / verify, the segmented aggregateAll, and BASELINE config 2 (Eth2 block batch) — all through the C ABI,
against BLST (oracle/_ref) and against the reference's own KATs (tests/golden/eth2_keys_and_proofs.json)."""
import hashlib
import json
import os
import random

import pytest

pytestmark = pytest.mark.gpu
GOLD = os.path.join(os.path.dirname(__file__), "golden")
P = 0x1a0111ea397fe69a4b1ba7b6434bacd764774b84f38512bf6730d2a0f6b0f6241eabfffeb153ffffb9feffffffffaaab


def g1_generator_mem():
    from oracle import pyref as pr
    return pr.g1_to_mem(pr.G1_GEN)


def test_reference_kats_keys_and_proofs(cache, br):
    """eth2_vectors.nim:33-75: pk/proof deserialise, pk == [sk]G1, proof == [sk]H_pop(pk), popVerify, wrong key fails."""
    import nim_blscurve_b200 as bg
    kat = json.load(open(os.path.join(GOLD, "eth2_keys_and_proofs.json")))
    dst_pop = kat["dst_pop"].encode()
    gen = g1_generator_mem()
    tr = kat["pop_triples"]
    pk48 = b"".join(bytes.fromhex(t["pk"]) for t in tr)
    pr96 = b"".join(bytes.fromhex(t["proof"]) for t in tr)
    pks, st = bg.publicKeysFromBytes(cache, pk48)
    assert st == [0, 0, 0]
    proofs, st = bg.signaturesFromBytes(cache, pr96)
    assert st == [0, 0, 0]
    assert bg.publicKeysToBytes(cache, pks) == pk48 and bg.signaturesToBytes(cache, proofs) == pr96
    for i, t in enumerate(tr):
        sk_le = bytes.fromhex(t["sk"])[::-1]
        pk, proof = pks[96 * i:96 * i + 96], proofs[192 * i:192 * i + 192]
        assert bg.msmG1(cache, gen, sk_le, 255) == pk                       # publicFromSecret
        _, h = bg.hashToG2(cache, pk48[48 * i:48 * i + 48], 48, dst_pop)
        assert bg.msmG2(cache, h, sk_le, 255) == proof                      # popProve
        assert bg.verify(cache, pk, pk48[48 * i:48 * i + 48], proof, dst_pop) is True          # popVerify
        wrong = pks[96 * ((i + 1) % 3):96 * ((i + 1) % 3) + 96]
        assert bg.verify(cache, wrong, pk48[48 * ((i + 1) % 3):48 * ((i + 1) % 3) + 48], proof, dst_pop) is False
    for t in kat["priv_to_pub"]:                                            # priv_to_pub.nim:32-81
        pk = bg.msmG1(cache, gen, bytes.fromhex(t["sk"])[::-1], 255)
        assert bg.publicKeysToBytes(cache, pk) == bytes.fromhex(t["pk"])


def _corruptions_g1(br, rng):
    """Encodings covering every exit of blst_p1_uncompress / _deserialize + the subgroup check."""
    sets = br.make_sets(500, 12)
    cases = []
    for i in range(12):
        aff = sets[320 * i:320 * i + 96]
        cases.append(br.g1_compress(aff))
    good = cases[0]
    cases.append(bytes([good[0] ^ 0x20]) + good[1:])                        # other square root: valid, -P
    cases.append(bytes([good[0] & 0x7f]) + good[1:])                        # compressed bit missing
    cases.append(bytes([0xc0]) + bytes(47))                                 # infinity
    cases.append(bytes([0xc0]) + bytes(46) + b"\x01")                       # infinity with garbage
    cases.append(bytes([0xe0]) + bytes(47))                                 # infinity + sign bit
    cases.append(bytes([0x80 | 0x1a]) + (P.to_bytes(48, "big"))[1:])        # x == p
    cases.append(bytes([0x80 | 0x1f]) + b"\xff" * 47)                       # x > p
    cases.append(bytes([0x80]) + bytes(47))                                 # x == 0 -> (0, +-2): not in group
    x = 1
    while len(cases) < 40:                                                  # small x: on/off curve, never in G1
        cases.append(bytes([0x80 | (0x20 if rng.random() < 0.5 else 0)]) + x.to_bytes(47, "big"))
        x += 1
    return cases


def test_pubkeys_from_bytes(cache, br):
    import nim_blscurve_b200 as bg
    rng = random.Random(5)
    cases = _corruptions_g1(br, rng)
    want = [br.pubkey_from_bytes(c) for c in cases]
    assert {w[0] for w in want} >= {0, 1, 2, 3, 6}                          # every BLST_ERROR is exercised
    pts, st = bg.publicKeysFromBytes(cache, b"".join(cases), 48)
    assert st == [w[0] for w in want]
    for i, w in enumerate(want):
        assert pts[96 * i:96 * i + 96] == w[1], i
    # fromBytesKnownOnCurve: no subgroup check
    want2 = [br.pubkey_from_bytes(c, group_check=False) for c in cases]
    pts2, st2 = bg.publicKeysFromBytes(cache, b"".join(cases), 48, group_check=False)
    assert st2 == [w[0] for w in want2] and any(a != b for a, b in zip(st, st2))
    for i, w in enumerate(want2):
        assert pts2[96 * i:96 * i + 96] == w[1], i
    # 96-byte form: serialized, compressed-in-96, infinity, y corrupted, y >= p
    sets = br.make_sets(600, 4)
    ser = [br.g1_serialize(sets[320 * i:320 * i + 96]) for i in range(4)]
    c96 = list(ser)
    c96.append(ser[0][:95] + bytes([ser[0][95] ^ 1]))                       # not on curve
    c96.append(ser[1][:48] + P.to_bytes(48, "big"))                         # y == p
    c96.append(bytes([0x40]) + bytes(95))                                   # infinity
    c96.append(bytes([0x40]) + bytes(94) + b"\x07")
    c96.append(bytes(96))                                                   # all zero, no flag
    c96.append(br.g1_compress(sets[:96]) + bytes(48))                       # compressed form in a 96-byte slot
    c96.append(bytes([0x20]) + ser[2][1:])                                  # stray sign bit on an uncompressed point
    want3 = [br.pubkey_from_bytes(c) for c in c96]
    pts3, st3 = bg.publicKeysFromBytes(cache, b"".join(c96), 96)
    assert st3 == [w[0] for w in want3]
    for i, w in enumerate(want3):
        assert pts3[96 * i:96 * i + 96] == w[1], i


def test_signatures_from_bytes(cache, br):
    import nim_blscurve_b200 as bg
    rng = random.Random(6)
    sets = br.make_sets(700, 10)
    cases = [br.g2_compress(sets[320 * i + 128:320 * i + 320]) for i in range(10)]
    good = cases[0]
    cases.append(bytes([good[0] ^ 0x20]) + good[1:])
    cases.append(bytes([good[0] & 0x7f]) + good[1:])
    cases.append(bytes([0xc0]) + bytes(95))                                 # infinity signature is allowed
    cases.append(bytes([0xc0]) + bytes(94) + b"\x01")
    cases.append(bytes([0x80 | 0x1a]) + P.to_bytes(48, "big")[1:] + bytes(48))   # x.im == p
    cases.append(good[:48] + P.to_bytes(48, "big"))                         # x.re == p
    cases.append(bytes([0x80]) + bytes(95))                                 # x == 0
    while len(cases) < 36:                                                  # random x: half on curve, none in G2
        xi, xr = rng.randrange(P), rng.randrange(P)
        b = bytearray(xi.to_bytes(48, "big") + xr.to_bytes(48, "big"))
        b[0] = (b[0] & 0x1f) | 0x80 | (0x20 if rng.random() < 0.5 else 0)
        cases.append(bytes(b))
    want = [br.signature_from_bytes(c) for c in cases]
    assert {w[0] for w in want} >= {0, 1, 2, 3}
    pts, st = bg.signaturesFromBytes(cache, b"".join(cases), 96)
    assert st == [w[0] for w in want]
    for i, w in enumerate(want):
        assert pts[192 * i:192 * i + 192] == w[1], i
    want2 = [br.signature_from_bytes(c, group_check=False) for c in cases]
    pts2, st2 = bg.signaturesFromBytes(cache, b"".join(cases), 96, group_check=False)
    assert st2 == [w[0] for w in want2] and any(a != b for a, b in zip(st, st2))
    for i, w in enumerate(want2):
        assert pts2[192 * i:192 * i + 192] == w[1], i
    ser = [br.g2_serialize(sets[320 * i + 128:320 * i + 320]) for i in range(3)]
    c192 = list(ser)
    c192.append(ser[0][:191] + bytes([ser[0][191] ^ 1]))
    c192.append(ser[1][:96] + P.to_bytes(48, "big") + ser[1][144:])
    c192.append(bytes([0x40]) + bytes(191))
    c192.append(bytes(192))
    c192.append(br.g2_compress(sets[128:320]) + bytes(96))
    want3 = [br.signature_from_bytes(c) for c in c192]
    pts3, st3 = bg.signaturesFromBytes(cache, b"".join(c192), 192)
    assert st3 == [w[0] for w in want3]
    for i, w in enumerate(want3):
        assert pts3[192 * i:192 * i + 192] == w[1], i
    # round trip through the device compressor
    assert bg.signaturesToBytes(cache, b"".join(sets[320 * i + 128:320 * i + 320] for i in range(10))) == b"".join(cases[:10])
    assert bg.publicKeysToBytes(cache, b"".join(sets[320 * i:320 * i + 96] for i in range(10))) == \
        b"".join(br.g1_compress(sets[320 * i:320 * i + 96]) for i in range(10))


@pytest.mark.parametrize("n", [1, 2, 7, 64])
def test_aggregate_verify(cache, br, n):
    """aggregateVerify over ragged message lengths (0..101 bytes), valid and failing, verdict + GT bytes vs BLST."""
    import nim_blscurve_b200 as bg
    rng = random.Random(n)
    msgs = [bytes(rng.randrange(256) for _ in range(rng.choice([0, 1, 31, 32, 33, 55, 56, 64, 101]))) for _ in range(n)]
    if n > 1:
        msgs[0], msgs[1] = b"first", b""
    signed = [br.sign(900 + i, msgs[i]) for i in range(n)]
    pks, sigs = [s[0] for s in signed], [s[1] for s in signed]
    ok, agg = br.aggregate_g2(b"".join(sigs))
    assert ok
    got = bg.aggregateVerify(cache, pks, msgs, agg, want_gt=True)
    assert got == br.aggregate_verify(b"".join(pks), msgs, agg) and got[0] is True
    if n > 1:                                                              # keys permuted against messages -> false
        got = bg.aggregateVerify(cache, pks[::-1], msgs, agg, want_gt=True)
        assert got == br.aggregate_verify(b"".join(pks[::-1]), msgs, agg) and got[0] is False
    assert bg.aggregateVerify(cache, pks, msgs[:-1], agg) is False         # length mismatch (bls_sig_min_pubkey.nim:164)
    assert bg.aggregateVerify(cache, [], [], agg) is False                 # :167
    # infinite public key -> false; infinite signature -> the plain product, false
    got = bg.aggregateVerify(cache, [bytes(96)] + pks[1:], msgs, agg, want_gt=True)
    assert got == br.aggregate_verify(bytes(96) + b"".join(pks[1:]), msgs, agg) and got[0] is False
    got = bg.aggregateVerify(cache, pks, msgs, bytes(192), want_gt=True)
    assert got == br.aggregate_verify(b"".join(pks), msgs, bytes(192)) and got[0] is False


def test_verify_other_dst_and_long_messages(cache, br):
    """verify() = the one-pair case, under a non-Eth2 DST and messages up to 200 bytes."""
    import nim_blscurve_b200 as bg
    rng = random.Random(77)
    dst = b"QUUX-V01-CS02-with-BLS12381G2_XMD:SHA-256_SSWU_RO_"
    for L in (0, 1, 55, 56, 119, 200):
        m = bytes(rng.randrange(256) for _ in range(L))
        pk, sig = br.sign(40 + L, m, dst)
        assert bg.verify(cache, pk, m, sig, dst) is True
        assert bg.verify(cache, pk, m + b"x", sig, dst) is False
        assert bg.verify(cache, pk, m, sig) is False                       # Eth2 DST: different hash
        got = bg.aggregateVerify(cache, [pk], [m + b"x"], sig, dst=dst, want_gt=True)
        assert got == br.aggregate_verify(pk, [m + b"x"], sig, dst=dst)


@pytest.mark.parametrize("nkeys", [1, 2, 31, 32, 33, 128, 512])
def test_fast_aggregate_verify(cache, br, nkeys):
    import nim_blscurve_b200 as bg
    h = hashlib.sha256(b"sync committee %d" % nkeys).digest()
    member_pks, agg_set = br.fast_aggregate_set(3000, nkeys, h, threads=8)
    pks = [member_pks[96 * i:96 * i + 96] for i in range(nkeys)]
    sig = agg_set[128:]
    got = bg.fastAggregateVerify(cache, pks, h, sig, want_gt=True)
    assert got == br.fast_aggregate_verify(member_pks, h, sig) and got[0] is True
    if nkeys > 1:                                                           # one participant missing -> false, same GT
        got = bg.fastAggregateVerify(cache, pks[:-1], h, sig, want_gt=True)
        assert got == br.fast_aggregate_verify(member_pks[:-96], h, sig) and got[0] is False
    assert bg.fastAggregateVerify(cache, [], h, sig) is False


def test_fast_aggregate_verify_cancelling_keys(cache, br):
    """pk + (-pk) aggregates to infinity: coreVerify refuses it (aggregate.c:296)."""
    import nim_blscurve_b200 as bg
    s = br.make_set(1, b"m")
    pk = s[:96]
    y = int.from_bytes(pk[48:], "little")
    # negate y in the Montgomery domain: (p - y) is the Montgomery form of -y
    npk = pk[:48] + (P - y).to_bytes(48, "little")
    got = bg.fastAggregateVerify(cache, [pk, npk], b"m", s[128:], want_gt=True)
    assert got == br.fast_aggregate_verify(pk + npk, b"m", s[128:]) and got[0] is False


def test_segmented_aggregate_all(cache, br):
    import nim_blscurve_b200 as bg
    sets = br.make_sets(5000, 300, threads=8)
    pks = [sets[320 * i:320 * i + 96] for i in range(300)]
    sizes = [1, 2, 3, 31, 32, 33, 64, 100, 0, 34]
    groups, o = [], 0
    for sz in sizes:
        groups.append(pks[o:o + sz])
        o += sz
    got = bg.aggregateAllSegments(cache, groups)
    for g, (ok, pt) in zip(groups, got):
        rok, rpt = br.aggregate_g1(b"".join(g))
        assert ok == rok
        if rok:
            assert pt == rpt
        else:
            assert pt == bytes(96)


def test_eth2_block_batch_config2(cache, br, srb):
    """BASELINE configs[1] / SURVEY §8d config 2: 128 attestation sets whose public key is the aggregate of a 128-key
    committee + one sync-committee set over 512 keys -> a 129-set batch; the key aggregation runs on the GPU (segmented
    aggregateAll) and must give the sets BLST gives; the batch verifies (serial and 4-chunk scalars), GT bit-exact."""
    import nim_blscurve_b200 as bg
    groups, ref_sets = [], []
    for i in range(128):
        h = hashlib.sha256(b"attestation %d" % i).digest()
        member_pks, agg_set = br.fast_aggregate_set(100000 + 1000 * i, 128, h, threads=8)
        groups.append([member_pks[96 * k:96 * k + 96] for k in range(128)])
        ref_sets.append(agg_set)
    h = hashlib.sha256(b"sync committee").digest()
    member_pks, agg_set = br.fast_aggregate_set(900000, 512, h, threads=8)
    groups.append([member_pks[96 * k:96 * k + 96] for k in range(512)])
    ref_sets.append(agg_set)
    agg = bg.aggregateAllSegments(cache, groups)
    sets = b""
    for (ok, pk), rs in zip(agg, ref_sets):
        assert ok and pk == rs[:96]
        sets += pk + rs[96:]
    assert sets == b"".join(ref_sets)
    for chunks in (0, 4):
        got = cache.verify_raw(sets, srb, chunks, want_gt=True)
        assert got == br.batch_verify(sets, srb, chunks) and got[0] is True
    bad = bytearray(sets)
    bad[320 * 77 + 128:320 * 78] = sets[320 * 5 + 128:320 * 6]              # attestation 77 carries another signature
    got = cache.verify_raw(bytes(bad), srb, 4, want_gt=True)
    assert got == br.batch_verify(bytes(bad), srb, 4) and got[0] is False
