"""GPU: the multi-device context behind the C ABI (blsgpu_create_multi, include/blsgpu.h) driven by a plain C program
(tests/c_abi_multi.c, compiled by __graft_entry__.build()) — the call a Nim consumer of the drop-in makes — and by the
Python mirror.  ONE blsgpu_batch_verify call fans the batch out over the devices (replacing the Taskpools fan-out of
/root/reference/blscurve/bls_batch_verifier.nim:316-369), gathers the 576-byte partials and runs one final
exponentiation; verdict and GT must equal the one-device result and BLST's."""
import ctypes as C
import os
import subprocess

import pytest

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
EXE = os.path.join(ROOT, "tests", "c_abi_multi")


def _exe():
    if not os.path.exists(EXE):
        import __graft_entry__ as g
        g.build_c_abi_test()
    return EXE


@pytest.mark.parametrize("n,shares", [(32768, 0), (4099, 3), (257, 5)])
def test_c_program_multi_device_batch_verify(n, shares):
    """shares = 0: one share per visible GPU (two shares on a one-GPU box); otherwise `shares` shares round-robin over
    the visible devices.  The program checks verdict + GT (valid, corrupted, infinite key, tiny, empty, over capacity)
    and the sharded MSM against the one-device results; exit code 0 = all equal."""
    r = subprocess.run([_exe(), str(n), str(shares)], capture_output=True, text=True, timeout=600)
    assert r.returncode == 0, r.stdout + r.stderr
    assert "c_abi_multi OK" in r.stdout


def test_multi_context_matches_blst(br, srb):
    """The same entry point through ctypes, against the oracle: 3 shares, 300 sets, chunks 0 / 4 / 16, one corrupted."""
    import nim_blscurve_b200 as bg
    L = bg.lib()
    ndev = L.blsgpu_device_count()
    devs = (C.c_int * 3)(*[k % ndev for k in range(3)])
    h = L.blsgpu_create_multi(devs, 3, 300)
    assert h, L.blsgpu_last_error(None)
    try:
        assert L.blsgpu_device_span(h) == 3 and L.blsgpu_capacity(h) >= 300
        sets = br.make_sets(0, 300)
        bad = bytearray(sets)
        bad[299 * 320 + 97] ^= 0x40
        for s in (sets, bytes(bad)):
            for chunks in (0, 4, 16):
                gt = (C.c_uint8 * 576)()
                rc = L.blsgpu_batch_verify(h, s, 300, srb, chunks, None, gt)
                assert rc >= 0, L.blsgpu_last_error(h)
                assert (bool(rc), bytes(gt)) == br.batch_verify(s, srb, chunks)
    finally:
        L.blsgpu_destroy(h)


def test_python_mirror_multi_device_cache(br, srb):
    """BatchedBLSVerifierCache(devices=[...]): the mirror of the shim's -d:blsgpuNumDevices, growing on demand."""
    import nim_blscurve_b200 as bg
    ndev = bg.lib().blsgpu_device_count()
    cache = bg.BatchedBLSVerifierCache(max_sets=8, devices=[k % ndev for k in range(2)])
    try:
        sets = br.make_sets(0, 40)                            # past the initial capacity: the context is re-created
        tp = bg.Taskpool.new(4)
        assert bg.batchVerify(tp, cache, sets, srb) is True
        bad = bytearray(sets)
        bad[7 * 320 + 96] ^= 1
        assert cache.verify_raw(bytes(bad), srb, 4, want_gt=True) == br.batch_verify(bytes(bad), srb, 4)
        assert bg.lib().blsgpu_device_span(cache.handle) == 2
    finally:
        cache.close()
