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


# ---- the Nim shim cannot be compiled here (no Nim toolchain): check it mechanically instead -------------------------
NIM_DIR = os.path.join(ROOT, "nim_blscurve_b200", "nim", "blscurve")


def _c_prototypes():
    """{name: (return type class, [parameter type classes])} from include/blsgpu.h."""
    h = open(os.path.join(ROOT, "include", "blsgpu.h")).read()
    h = re.sub(r"/\*.*?\*/", "", h, flags=re.S)
    protos = {}
    for m in re.finditer(r"([A-Za-z_][A-Za-z0-9_ \*]*?)\b(blsgpu_[a-z0-9_]+)\s*\(([^)]*)\)\s*;", h):
        ret, name, params = m.group(1).strip(), m.group(2), m.group(3).strip()
        plist = [] if params in ("", "void") else [p.strip() for p in params.split(",")]
        protos[name] = (_c_class(ret, is_ret=True), [_c_class(p) for p in plist])
    return protos


def _c_class(decl, is_ret=False):
    d = re.sub(r"\bconst\b", "", decl).strip()
    if "[" in d:                                              # uint8_t x[576] decays to a pointer
        return "ptr"
    if "blsgpu_ctx" in d and "*" in d:
        return "ctx"
    if "char" in d and "*" in d:
        return "cstring" if is_ret else "ptr"
    if "*" in d:
        return "ptr"
    base = d if is_ret else " ".join(d.split()[:-1])          # drop the parameter name
    return {"int": "int", "size_t": "size_t", "uint32_t": "u32", "uint64_t": "u64", "double": "double",
            "float": "float", "void": "void"}[base.strip()]


def _nim_class(t):
    t = t.strip()
    if t == "BlsGpuCtx":
        return "ctx"
    if t == "cstring":
        return "cstring"
    if t == "pointer" or t.startswith("ptr "):
        return "ptr"
    return {"cint": "int", "csize_t": "size_t", "uint32": "u32", "uint64": "u64", "cdouble": "double",
            "cfloat": "float"}[t]


def _nim_prototypes():
    src = open(os.path.join(NIM_DIR, "cuda", "blsgpu_abi.nim")).read()
    src = re.sub(r"#.*", "", src)
    protos = {}
    for m in re.finditer(r"proc\s+(blsgpu_[a-z0-9_]+)\*\s*\(([^)]*)\)\s*(?::\s*([A-Za-z_0-9]+))?", src, flags=re.S):
        name, params, ret = m.group(1), " ".join(m.group(2).split()), m.group(3)
        classes = []
        if params:
            # "a, b: T" declares two parameters of type T; array[32, byte] contains a comma, so split on ':' first
            for group in re.findall(r"([A-Za-z0-9_, ]+?):\s*((?:ptr\s+)?(?:array\[[^\]]*\]|[A-Za-z0-9_]+))", params):
                names = [x for x in group[0].split(",") if x.strip()]
                classes += [_nim_class(group[1])] * len(names)
        protos[name] = (_nim_class(ret) if ret else "void", classes)
    return protos


def test_nim_ffi_declarations_match_the_c_header():
    c, nim = _c_prototypes(), _nim_prototypes()
    assert sorted(c) == header_symbols()
    assert sorted(nim) == sorted(c), (sorted(set(c) - set(nim)), sorted(set(nim) - set(c)))
    for name in sorted(c):
        assert nim[name] == c[name], (name, "nim", nim[name], "c", c[name])


def _nim_public_procs(path):
    """[(name, [parameter names])] of every exported proc/func, comments stripped."""
    src = re.sub(r"##?.*", "", open(path).read())
    out = []
    for m in re.finditer(r"\b(?:proc|func)\s+`?([A-Za-z][A-Za-z0-9_]*)`?\*\s*(?:\[[^\]]*\])?\s*\(([^)]*)\)", src, flags=re.S):
        params = " ".join(m.group(2).split())
        names = []
        for group in re.findall(r"([A-Za-z0-9_, ]+?):\s*(?:var\s+|ptr\s+|type\s+)?(?:[A-Za-z]+\[[^\]]*\]|[A-Za-z0-9_\.]+)", params):
            names += [x.strip() for x in group.split(",") if x.strip()]
        out.append((m.group(1), names))
    return out


REFERENCE_OVERLOADS = [
    # blscurve/bls_batch_verifier.nim (reference lines in the comment column), parameter names verbatim
    ("init", ["T", "pubkeys", "message", "signatures"]),                                   # :73
    ("init", ["T", "sigset"]),                                                             # :86
    ("add", ["multiSet", "sigset"]),                                                       # :93
    ("combine", ["multiSet", "secureRandomBytes"]),                                        # :100
    ("init", ["T"]),                                                                       # :108
    ("init", ["T", "tp"]),                                                                 # :115
    ("batchVerifySerial", ["cache", "input", "secureRandomBytes"]),                        # :121
    ("batchVerifySerial", ["input", "secureRandomBytes"]),                                 # :162
    ("batchVerifyParallel", ["tp", "cache", "setsPtr", "numSets", "secureRandomBytes"]),   # :296
    ("batchVerifyParallel", ["tp", "cache", "input", "secureRandomBytes"]),                # :373
    ("batchVerifyParallel", ["tp", "input", "secureRandomBytes"]),                         # :399
    ("batchVerify", ["tp", "cache", "setsPtr", "numSets", "secureRandomBytes"]),           # :420
    ("batchVerify", ["tp", "cache", "input", "secureRandomBytes"]),                        # :449
    ("batchVerify", ["tp", "input", "secureRandomBytes"]),                                 # :475
    # blscurve/blst/blst_min_pubkey_sig_core.nim
    ("aggregateAll", ["dst", "elems"]),                                                    # :179
    ("subtractAll", ["dst", "elems"]),                                                     # :197
    ("combine", ["secureRandomBytes", "publicKeys", "signatures"]),                        # :570
]


def test_nim_shim_has_every_reference_overload_with_the_reference_parameter_lists():
    have = _nim_public_procs(os.path.join(NIM_DIR, "cuda", "bls_batch_verifier_cuda.nim"))
    for want in REFERENCE_OVERLOADS:
        assert want in have, ("missing overload", want, [h for h in have if h[0] == want[0]])
    src = open(os.path.join(NIM_DIR, "cuda", "bls_batch_verifier_cuda.nim")).read()
    assert "`=copy`" in src and "{.error" in src            # a cache owns a device context: no implicit copies
    assert "ensureCapacity" in src                           # no fixed size limit (the reference has none)
    # every blsgpu_* call in the shim is a declared FFI proc
    used = set(re.findall(r"\b(blsgpu_[a-z0-9_]+)\s*\(", src))
    assert used <= set(_nim_prototypes()), used - set(_nim_prototypes())


def test_backend_switch_has_the_cuda_branch():
    src = open(os.path.join(NIM_DIR, "bls_backend_cuda.nim")).read()
    assert 'BLS_FORCE_BACKEND == "cuda"' in src and "CUDA" in src and "cuda/blsgpu_abi" in src
