"""GPU parity: G1 multi-scalar multiplication vs blst_p1s_mult_pippenger (affine result, bit-exact)."""
import ctypes as C
import random

import pytest

pytestmark = pytest.mark.gpu


@pytest.mark.parametrize("n", [1, 2, 3, 31, 32, 33, 100, 1000, 4096, 20000])
def test_msm_g1_255(cache, br, n):
    import nim_blscurve_b200 as bg
    pts, sc = br.msm_points(0xFACADE, n)
    assert bg.msmG1(cache, pts, sc, 255) == br.msm_g1(pts, sc, 255)


def test_msm_g1_edge_scalars(cache, br):
    import nim_blscurve_b200 as bg
    n = 64
    pts, sc = br.msm_points(5, n)
    sc = bytearray(sc)
    sc[0:32] = bytes(32)                                  # zero scalar
    sc[32:64] = b"\xff" * 31 + b"\x7f"                    # 2^255 - 1 (all digits carry)
    sc[64:96] = b"\x01" + bytes(31)
    sc[96:128] = bytes(31) + b"\x40"                      # single top bit
    pts = bytearray(pts)
    pts[96 * 5:96 * 6] = bytes(96)                        # point at infinity
    pts[96 * 7:96 * 8] = pts[96 * 6:96 * 7]               # repeated point (bucket doubling)
    sc[32 * 7:32 * 8] = sc[32 * 6:32 * 7]
    assert bg.msmG1(cache, bytes(pts), bytes(sc), 255) == br.msm_g1(bytes(pts), bytes(sc), 255)


@pytest.mark.parametrize("nbits", [64, 128, 200])
def test_msm_g1_short_scalars(cache, br, nbits):
    """nbits=64 is the shape MultiSignatureSet.combine uses (blst_min_pubkey_sig_core.nim:629-636)."""
    import nim_blscurve_b200 as bg
    n = 257
    pts, _ = br.msm_points(9, n)
    rng = random.Random(nbits)
    sb = (nbits + 7) // 8
    sc = b"".join(rng.getrandbits(nbits).to_bytes(sb, "little") for _ in range(n))
    assert bg.msmG1(cache, pts, sc, nbits) == br.msm_g1(pts, sc, nbits)


def test_msm_device_inputs_match_oracle(cache, br):
    """The benchmark's device-side input generator + MSM against BLST on the same bytes."""
    import torch
    import nim_blscurve_b200 as bg
    n = 3000
    dp = torch.empty(n * 96, dtype=torch.uint8, device="cuda")
    ds = torch.empty(n * 32, dtype=torch.uint8, device="cuda")
    L = bg.lib()
    assert L.blsgpu_msm_make_inputs(cache.handle, 11, n, C.c_void_p(dp.data_ptr()), C.c_void_p(ds.data_ptr())) == 0
    out = (C.c_uint8 * 96)()
    assert L.blsgpu_msm_g1_dev(cache.handle, C.c_void_p(dp.data_ptr()), C.c_void_p(ds.data_ptr()), n, 255, out) == 1
    assert bytes(out) == br.msm_g1(dp.cpu().numpy().tobytes(), ds.cpu().numpy().tobytes(), 255)


def _g2_points(br, n, seed=3):
    """n affine G2 points: the signatures of deterministic signature sets (192 B each)."""
    sets = br.make_sets(seed, n)
    return b"".join(sets[i * 320 + 128:(i + 1) * 320] for i in range(n))


@pytest.mark.parametrize("n,nbits", [(1, 64), (2, 64), (33, 64), (257, 64), (3000, 64), (500, 255), (129, 128)])
def test_msm_g2(cache, br, n, nbits):
    """G2 MSM vs blst_p2s_mult_pippenger; nbits=64 is the signature half of MultiSignatureSet.combine
    (blst_min_pubkey_sig_core.nim:637-644)."""
    import nim_blscurve_b200 as bg
    pts = _g2_points(br, n)
    rng = random.Random(n * 1000 + nbits)
    sb = (nbits + 7) // 8
    sc = b"".join(rng.getrandbits(nbits).to_bytes(sb, "little") for _ in range(n))
    assert bg.msmG2(cache, pts, sc, nbits) == br.msm_g2(pts, sc, nbits)


def test_msm_skewed_scalars(cache, br):
    """Every scalar equal: each window has ONE full bucket (the task split + warp-cooperative combine path)."""
    import nim_blscurve_b200 as bg
    n = 5000
    pts, _ = br.msm_points(21, n)
    sc = (0x1234567890abcdef1122334455667788).to_bytes(32, "little") * n
    assert bg.msmG1(cache, pts, sc, 255) == br.msm_g1(pts, sc, 255)
    g2 = _g2_points(br, 700)
    sc2 = (0xfedcba9876543210).to_bytes(8, "little") * 700
    assert bg.msmG2(cache, g2, sc2, 64) == br.msm_g2(g2, sc2, 64)


def test_msm_g1_window_horner_edges(cache, br):
    """The window Horner runs as a branch-free dataflow program over complete projective formulas (fpprog.hpp
    build_msm_horner_g1): results at infinity, empty top windows, a lone low window and windows that cancel must come
    out exactly as blst_p1s_mult_pippenger + to_affine gives them."""
    import nim_blscurve_b200 as bg
    n = 40
    pts, sc = br.msm_points(31, n)
    zero = bytes(32 * n)
    assert bg.msmG1(cache, pts, zero, 255) == br.msm_g1(pts, zero, 255) == bytes(96)          # all windows infinite
    small = b"".join((i + 1).to_bytes(32, "little") for i in range(n))                         # only window 0 is populated
    assert bg.msmG1(cache, pts, small, 255) == br.msm_g1(pts, small, 255)
    top = b"".join(((i + 1) << 240).to_bytes(32, "little") for i in range(n))                  # only the top bits
    assert bg.msmG1(cache, pts, top, 255) == br.msm_g1(pts, top, 255)
    # P_1 = -P_0 with equal scalars: every window sum cancels -> infinity; with different scalars: they do not
    p = bytearray(pts[:192])
    from oracle import pyref as pr
    neg = pr.fp_to_mont_bytes((-pr.fp_from_mont_bytes(bytes(p[48:96]))) % pr.P)
    p[96:144] = p[0:48]
    p[144:192] = neg
    s1 = sc[:32] * 2
    assert bg.msmG1(cache, bytes(p), s1, 255) == br.msm_g1(bytes(p), s1, 255) == bytes(96)
    s2 = sc[:64]
    assert bg.msmG1(cache, bytes(p), s2, 255) == br.msm_g1(bytes(p), s2, 255)
    # one point (c = 2, 128 windows): sparse and all-ones scalars; the P + P / P - P cases of the program's addition
    # are pinned on the CPU in tests/test_hostsim.py::test_msm_horner_program
    one = pts[:96]
    for k in (1 << 16 | 1, (1 << 32) | (1 << 16) | 1, (1 << 255) - 1):
        s = (k % (1 << 255)).to_bytes(32, "little")
        assert bg.msmG1(cache, one, s, 255) == br.msm_g1(one, s, 255)


def _device_msm_inputs(cache, n, seed):
    import torch
    import nim_blscurve_b200 as bg
    dp = torch.empty(n * 96, dtype=torch.uint8, device="cuda")
    ds = torch.empty(n * 32, dtype=torch.uint8, device="cuda")
    assert bg.lib().blsgpu_msm_make_inputs(cache.handle, seed, n, C.c_void_p(dp.data_ptr()), C.c_void_p(ds.data_ptr())) == 0
    return dp, ds


def _msm_dev(cache, dp, ds, n, first=0, nbits=255):
    import nim_blscurve_b200 as bg
    out = (C.c_uint8 * 96)()
    rc = bg.lib().blsgpu_msm_g1_dev(cache.handle, C.c_void_p(dp.data_ptr() + 96 * first),
                                    C.c_void_p(ds.data_ptr() + 32 * first), n, nbits, out)
    assert rc == 1, cache.last_error()
    return bytes(out)


def test_msm_g1_full_size_vs_blst(cache, br):
    """BASELINE configs[2] at its headline size: 2^20 points, 255-bit scalars, against blst_p1s_mult_pippenger
    (one host thread, ~8 s) on the same bytes."""
    n = 1 << 20
    dp, ds = _device_msm_inputs(cache, n, 0xFACADE)
    got = _msm_dev(cache, dp, ds, n)
    assert got == br.msm_g1(dp.cpu().numpy().tobytes(), ds.cpu().numpy().tobytes(), 255)


def test_msm_g1_size_independent_properties(cache, br):
    """2^21 points (beyond what the oracle finishes quickly): the sum over the whole range equals the aggregate
    (aggregateAll, a different code path) of the sums over four unequal slices — slices of 2^19, 2^20 and two odd
    sizes use different window shapes (c = 16, 16, 13...), so this also checks shape independence; and a slice that
    fits the oracle budget is pinned against BLST."""
    import nim_blscurve_b200 as bg
    n = 1 << 21
    dp, ds = _device_msm_inputs(cache, n, 77)
    whole = _msm_dev(cache, dp, ds, n)
    cuts = [0, 1 << 19, (1 << 19) + (1 << 20), (1 << 21) - 70001, n]
    parts = [_msm_dev(cache, dp, ds, cuts[i + 1] - cuts[i], cuts[i]) for i in range(4)]
    ok, total = bg.aggregateAll(cache, parts)
    assert ok and total == whole
    lo, cnt = cuts[3], cuts[4] - cuts[3]
    pts = dp[96 * lo:96 * (lo + cnt)].cpu().numpy().tobytes()
    sc = ds[32 * lo:32 * (lo + cnt)].cpu().numpy().tobytes()
    assert parts[3] == br.msm_g1(pts, sc, 255)


@pytest.mark.parametrize("world", [2, 3, 8])
def test_msm_rank_shares_on_one_device(cache, br, world):
    """SURVEY.md §8e MSM row on one device: the per-rank MSMs over shard_range slices, summed by aggregateAll as
    msm_g1_distributed does after its all-gather, equal blst_p1s_mult_pippenger over all points (n = 5 leaves ranks
    with an empty share at world 8)."""
    import nim_blscurve_b200 as bg
    be = bg.GpuBackend(cache)
    for n in (5, 1000):
        pts, sc = br.msm_points(77, n)
        parts = b""
        for r in range(world):
            f, c = bg.shard_range(n, world, r)
            parts += bg.msm_g1_distributed(be, pts[f * 96:(f + c) * 96], sc[f * 32:(f + c) * 32], 255)   # world-1 path
        assert be.aggregate_g1(parts) == br.msm_g1(pts, sc, 255)


def test_msm_g1_2pow22_slice_sum(cache, br):
    """BASELINE configs[2] at its largest size, 2^22 points (c = 19 in the reference's window rule,
    multi_scalar.c:275-289): the MSM over the whole range equals the aggregateAll of the MSMs over five unequal slices
    (2^21, 2^20, 2^19 and two odd sizes: different window shapes), and the last slice — small enough for the oracle —
    is pinned against blst_p1s_mult_pippenger."""
    import nim_blscurve_b200 as bg
    n = 1 << 22
    dp, ds = _device_msm_inputs(cache, n, 0x2222)
    whole = _msm_dev(cache, dp, ds, n)
    cuts = [0, 1 << 21, (1 << 21) + (1 << 20), (1 << 21) + (1 << 20) + (1 << 19), n - 50021, n]
    parts = [_msm_dev(cache, dp, ds, cuts[i + 1] - cuts[i], cuts[i]) for i in range(5)]
    ok, total = bg.aggregateAll(cache, parts)
    assert ok and total == whole
    lo, cnt = cuts[4], cuts[5] - cuts[4]
    pts = dp[96 * lo:96 * (lo + cnt)].cpu().numpy().tobytes()
    sc = ds[32 * lo:32 * (lo + cnt)].cpu().numpy().tobytes()
    assert parts[4] == br.msm_g1(pts, sc, 255)
