"""CPU: the C-ABI library loads and exports every symbol include/blsgpu.h declares; host-side logic of the
reference-facing mirror (chunking, dispatch rule, packing); the product fails loudly without a GPU."""
import ctypes as C
import os
import re
import subprocess

import pytest

import nim_blscurve_b200 as bg
from nim_blscurve_b200 import _lib
from oracle import pyref as pr

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def header_symbols():
    h = open(os.path.join(ROOT, "include", "blsgpu.h")).read()
    h = re.sub(r"/\*.*?\*/", "", h, flags=re.S)
    return sorted(set(re.findall(r"\b(blsgpu_[a-z0-9_]+)\s*\(", h)))


def test_library_exports_every_declared_symbol():
    assert os.path.exists(bg.LIB_PATH), "libblsgpu.so not built (run __graft_entry__.build())"
    syms = header_symbols()
    assert len(syms) >= 20
    out = subprocess.run(["nm", "-D", "--defined-only", bg.LIB_PATH], capture_output=True, text=True, check=True).stdout
    exported = set(re.findall(r" T (blsgpu_[a-z0-9_]+)", out))
    missing = [s for s in syms if s not in exported]
    assert not missing, missing
    L = bg.lib()
    for s in syms:
        assert hasattr(L, s)
    assert sorted(_lib.SYMBOLS) == syms, "python binding list out of sync with include/blsgpu.h"


def test_library_has_sm100a_code_and_no_cpu_path():
    out = subprocess.run(["cuobjdump", "-lelf", bg.LIB_PATH], capture_output=True, text=True).stdout
    assert "sm_100a" in out


@pytest.mark.skipif(bg.lib().blsgpu_device_count() > 0, reason="a GPU is present")
def test_fails_loudly_without_gpu():
    L = bg.lib()
    assert L.blsgpu_device_count() == 0
    assert not L.blsgpu_create(0, 16)
    assert b"no CUDA device" in L.blsgpu_last_error(None)
    with pytest.raises(bg.BlsGpuError):
        bg.BatchedBLSVerifierCache()


def test_shard_range_is_parallel_chunks():
    for total in (0, 1, 2, 7, 40, 129, 32768, 1000003):
        for world in (1, 2, 3, 4, 8, 12):
            covered = 0
            for r in range(world):
                first, cnt = bg.shard_range(total, world, r)
                assert (first, cnt) == pr.parallel_chunks(world, total, r)
                assert first == covered
                covered += cnt
            assert covered == total
    # the example of blscurve/parallel_chunks.nim:27-33: 40 items on 12 threads -> 4,4,4,4,3,...
    assert [bg.shard_range(40, 12, r)[1] for r in range(12)] == [4] * 4 + [3] * 8


def test_signature_set_packing_and_dispatch_rule(monkeypatch):
    s = bg.SignatureSet(b"\x01" * 96, b"\x02" * 32, b"\x03" * 192)
    assert len(s.to_bytes()) == 320 and s.to_bytes()[96:128] == b"\x02" * 32
    calls = []

    class FakeCache:
        def verify_raw(self, sets, srb, chunks, scalars=None, want_gt=False):
            calls.append((len(sets) // 320, chunks))
            return True
    c = FakeCache()
    tp4, tp1 = bg.Taskpool.new(numThreads=4), bg.Taskpool.new(numThreads=1)
    assert bg.batchVerify(tp4, c, [s] * 3, b"\0" * 32)        # parallel: numThreads > 1 and len >= 3
    assert bg.batchVerify(tp4, c, [s] * 2, b"\0" * 32)        # serial: len < 3  (bls_batch_verifier.nim:468)
    assert bg.batchVerify(tp1, c, [s] * 5, b"\0" * 32)        # serial: one thread
    assert calls == [(3, 4), (2, 0), (5, 0)]
    assert bg.batchVerify(tp4, c, [], b"\0" * 32) is False    # empty -> false, no device call (:137, :312)
    assert bg.batchVerifySerial(c, [], b"\0" * 32) is False
    assert bg.batchVerifyParallel(tp4, c, [], b"\0" * 32) is False
    assert len(calls) == 3


def test_rlc_scalar_derivation_matches_reference_recipe():
    import hashlib
    srb = hashlib.sha256(b"Mr F was here").digest()
    # serial: seed = SHA256(srb), first scalar = LE64(SHA256(seed)[:8])
    seed = hashlib.sha256(hashlib.sha256(srb).digest()).digest()
    assert pr.rlc_scalars(srb, 1, 0) == [int.from_bytes(seed[:8], "little")]
    # chunk tag is LE64(chunk id)
    seed = hashlib.sha256(hashlib.sha256(srb + (2).to_bytes(8, "little")).digest()).digest()
    off, _ = pr.parallel_chunks(4, 10, 2)
    assert pr.rlc_scalars(srb, 10, 4)[off] == int.from_bytes(seed[:8], "little")
    assert all(x != 0 for x in pr.rlc_scalars(srb, 50, 7))
