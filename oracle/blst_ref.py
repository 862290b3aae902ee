"""ctypes access to the reference's own arithmetic backend (BLST) — TEST INFRASTRUCTURE ONLY.

``oracle/_ref/libblst_ref.so`` is compiled from the sources where they lie under
/root/reference/vendor/blst (see oracle/Makefile); ``oracle/_ref/libref_batch.so``
is oracle/ref_batch.c (the C restatement of the thin Nim layer) linked against it.
Nothing under ``nim_blscurve_b200/`` may import this module.
"""
import ctypes as C
import os

_HERE = os.path.dirname(os.path.abspath(__file__))
_REF = os.path.join(_HERE, "_ref")


def _load(name):
    path = os.path.join(_REF, name)
    if not os.path.exists(path):
        raise RuntimeError(f"{path} missing: run `make -C oracle` (needs /root/reference) "
                           "or __graft_entry__.build()")
    return C.CDLL(path, mode=C.RTLD_GLOBAL)


blst = _load("libblst_ref.so")
ref = _load("libref_batch.so")

u8p = C.POINTER(C.c_uint8)
u64p = C.POINTER(C.c_uint64)


def _buf(b):
    return (C.c_uint8 * len(b)).from_buffer_copy(b)


def _out(n):
    return (C.c_uint8 * n)()


ref.ref_batch_verify.restype = C.c_int
ref.ref_batch_verify_mt.restype = C.c_int
ref.ref_msm_g1.restype = C.c_int
ref.ref_aggregate_g1.restype = C.c_int
ref.ref_aggregate_g2.restype = C.c_int
ref.ref_fast_aggregate_set.restype = C.c_int
ref.ref_time_batch_verify.restype = C.c_double
ref.ref_time_msm_g1.restype = C.c_double
ref.ref_ncores.restype = C.c_int
ref.ref_partial.restype = C.c_int
ref.ref_finalize.restype = C.c_int


def make_sets(start_seed, n, msg_prefix=b"msg", threads=0):
    """n signature sets as t_batch_verifier.nim:34-47 makes them: sk=keygen(LE64(seed)‖0..),
    msg=SHA256(prefix‖dec(seed)), sig=[sk]H(msg).  Returns n*320 bytes (reference memory layout)."""
    out = _out(320 * n)
    ref.ref_make_sets(C.c_uint64(start_seed), C.c_size_t(n), _buf(msg_prefix), C.c_size_t(len(msg_prefix)),
                      out, C.c_int(threads))
    return bytes(out)


def make_sets_device_recipe(seed, first, n, threads=0):
    """The bytes blsgpu_make_sets(seed, first, n) generates on the device, computed with BLST on the host cores (the
    benchmark workload for the reference arm; equality with the device generator is a GPU test)."""
    out = _out(320 * n)
    ref.ref_make_sets_device_recipe(C.c_uint64(seed), C.c_size_t(first), C.c_size_t(n), out, C.c_int(threads))
    return bytes(out)


def make_set(seed, message):
    """One set with an explicit message string (hashed with SHA-256 first)."""
    out = _out(320)
    ref.ref_make_set(C.c_uint64(seed), _buf(message), C.c_size_t(len(message)), out)
    return bytes(out)


def sign_hashed(seed, hashed32):
    out = _out(320)
    ref.ref_make_set_hashed(C.c_uint64(seed), _buf(hashed32), out)
    return bytes(out)


def rlc_scalars(srb, n, chunks):
    out = (C.c_uint64 * n)()
    ref.ref_rlc_scalars(_buf(srb), C.c_size_t(n), C.c_uint32(chunks), out)
    return list(out)


def batch_verify(sets, srb, chunks=0, scalars=None):
    """(ok, gt576) — chunks=0 → batchVerifySerial; chunks=T → batchVerifyParallel with T threads' chunking.
    scalars: optional explicit list of 64-bit scalars (overrides the derivation)."""
    n = len(sets) // 320
    gt = _out(576)
    sc = None
    if scalars is not None:
        sc = (C.c_uint64 * n)(*scalars)
    ok = ref.ref_batch_verify(_buf(sets) if n else None, C.c_size_t(n), _buf(srb), C.c_uint32(chunks), sc, gt)
    return bool(ok), bytes(gt)


def batch_verify_mt(sets, srb, threads):
    n = len(sets) // 320
    return bool(ref.ref_batch_verify_mt(_buf(sets), C.c_size_t(n), _buf(srb), C.c_int(threads)))


def time_batch_verify(sets, srb, threads, reps=1):
    """Best-of-reps wall seconds of the pthreads replica of batchVerifyParallel (threads>=1)."""
    n = len(sets) // 320
    return float(ref.ref_time_batch_verify(_buf(sets), C.c_size_t(n), _buf(srb), C.c_int(threads), C.c_int(reps)))


def hash_to_g2(msgs, msg_len, dst):
    """Returns (compressed 96B each, affine-memory 192B each)."""
    n = len(msgs) // msg_len if msg_len else 0
    comp, aff = _out(96 * n), _out(192 * n)
    ref.ref_hash_to_g2(_buf(msgs), C.c_size_t(n), C.c_size_t(msg_len), _buf(dst), C.c_size_t(len(dst)), comp, aff)
    return bytes(comp), bytes(aff)


def msm_g1(points96, scalars32, nbits=255):
    n = len(points96) // 96
    out = _out(96)
    ref.ref_msm_g1(_buf(points96), _buf(scalars32), C.c_size_t(n), C.c_size_t(nbits), out)
    return bytes(out)


def msm_g2(points192, scalars, nbits=255):
    """blst_p2s_mult_pippenger + to_affine on n x 192-byte affine points, scalars n x ceil(nbits/8) bytes LE."""
    n = len(points192) // 192
    out = _out(192)
    ref.ref_msm_g2(_buf(points192), _buf(scalars), C.c_size_t(n), C.c_size_t(nbits), out)
    return bytes(out)


def time_msm_g1(points96, scalars32, nbits=255, reps=1):
    n = len(points96) // 96
    return float(ref.ref_time_msm_g1(_buf(points96), _buf(scalars32), C.c_size_t(n), C.c_size_t(nbits), C.c_int(reps)))


def msm_points(seed, n, threads=0):
    """(points n*96, scalars n*32) as benchmarks/bls12381_msm_g1.nim:22-44 shapes them."""
    pts, sc = _out(96 * n), _out(32 * n)
    ref.ref_msm_inputs(C.c_uint64(seed), C.c_size_t(n), pts, sc, C.c_int(threads))
    return bytes(pts), bytes(sc)


def aggregate_g1(points96):
    n = len(points96) // 96
    out = _out(96)
    ok = ref.ref_aggregate_g1(_buf(points96) if n else None, C.c_size_t(n), out)
    return bool(ok), bytes(out)


def aggregate_g2(points192):
    n = len(points192) // 192
    out = _out(192)
    ok = ref.ref_aggregate_g2(_buf(points192) if n else None, C.c_size_t(n), out)
    return bool(ok), bytes(out)


def subtract_all(dst, points):
    """subtractAll (blst_min_pubkey_sig_core.nim:197-209): dst - sum(points); 96-byte (G1) or 192-byte (G2) affine points."""
    sz = len(dst)
    assert sz in (96, 192) and len(points) % sz == 0
    n = len(points) // sz
    out = (C.c_uint8 * sz).from_buffer_copy(dst)
    (ref.ref_subtract_g1 if sz == 96 else ref.ref_subtract_g2)(out, _buf(points) if n else None, C.c_size_t(n))
    return bytes(out)


def fast_aggregate_set(start_seed, nkeys, hashed32, threads=0):
    """Committee of nkeys signers on one message: returns (member pubkeys nkeys*96, aggregate set 320 B)."""
    pks, out = _out(96 * nkeys), _out(320)
    ref.ref_fast_aggregate_set(C.c_uint64(start_seed), C.c_size_t(nkeys), _buf(hashed32), pks, out, C.c_int(threads))
    return bytes(pks), bytes(out)


def combine(srb, pubkeys96, sigs192):
    n = len(pubkeys96) // 96
    pk, sig = _out(96), _out(192)
    ref.ref_combine(_buf(srb), _buf(pubkeys96), _buf(sigs192), C.c_size_t(n), pk, sig)
    return bytes(pk), bytes(sig)


def g1_compress(aff96):
    out = _out(48)
    blst.blst_p1_affine_compress(out, _buf(aff96))
    return bytes(out)


def g2_compress(aff192):
    out = _out(96)
    blst.blst_p2_affine_compress(out, _buf(aff192))
    return bytes(out)


def g2_neg(aff192):
    out = _out(192)
    ref.ref_g2_neg(_buf(aff192), out)
    return bytes(out)


def fp_op(op, a48, b48=None):
    """Raw Montgomery-domain field ops for unit tests of fp.cuh: 'mul','sqr','add','sub','inv'."""
    out = _out(48)
    fn = {"mul": blst.blst_fp_mul, "add": blst.blst_fp_add, "sub": blst.blst_fp_sub}.get(op)
    if fn is not None:
        fn(out, _buf(a48), _buf(b48))
    elif op == "sqr":
        blst.blst_fp_sqr(out, _buf(a48))
    elif op == "inv":
        blst.blst_fp_inverse(out, _buf(a48))
    else:
        raise ValueError(op)
    return bytes(out)


def ncores():
    return int(ref.ref_ncores())


def partial(sets, first, total_n, srb, chunks):
    """One rank's 576-byte Fp12 partial (BLST in-memory layout) and its failure flag."""
    n = len(sets) // 320
    out, flag = _out(576), C.c_int(0)
    ref.ref_partial(_buf(sets) if n else None, C.c_size_t(n), C.c_size_t(first), C.c_size_t(total_n), _buf(srb),
                    C.c_uint32(chunks), out, C.byref(flag))
    return bytes(out), int(flag.value)


def finalize(partials):
    count = len(partials) // 576
    gt = _out(576)
    ok = ref.ref_finalize(_buf(partials), C.c_size_t(count), gt)
    return bool(ok), bytes(gt)


# ---- aggregateVerify / fastAggregateVerify (bls_sig_min_pubkey.nim:127-273) ----
ref.ref_aggregate_verify.restype = C.c_int
ref.ref_fast_aggregate_verify.restype = C.c_int
ref.ref_pubkey_from_bytes.restype = C.c_int
ref.ref_signature_from_bytes.restype = C.c_int
DST_ETH2 = b"BLS_SIG_BLS12381G2_XMD:SHA-256_SSWU_RO_POP_"


def _offsets(msgs):
    offs, o = [0], 0
    for m in msgs:
        o += len(m)
        offs.append(o)
    return (C.c_uint32 * len(offs))(*offs)


def aggregate_verify(pubkeys96, msgs, sig192, dst=DST_ETH2):
    """(bool, GT bytes) of aggregateVerify over n (public key, message) pairs and one signature."""
    n = len(pubkeys96) // 96
    assert n == len(msgs)
    gt = _out(576)
    blob = b"".join(msgs)
    ok = ref.ref_aggregate_verify(_buf(pubkeys96) if n else None, C.c_size_t(n), _buf(blob) if blob else None,
                                  _offsets(msgs), _buf(dst), C.c_size_t(len(dst)), _buf(sig192), gt)
    return bool(ok), bytes(gt)


def fast_aggregate_verify(pubkeys96, msg, sig192, dst=DST_ETH2):
    n = len(pubkeys96) // 96
    gt = _out(576)
    ok = ref.ref_fast_aggregate_verify(_buf(pubkeys96) if n else None, C.c_size_t(n), _buf(msg) if msg else None,
                                       C.c_size_t(len(msg)), _buf(dst), C.c_size_t(len(dst)), _buf(sig192), gt)
    return bool(ok), bytes(gt)


# ---- fromBytes with checks (bls_sig_io.nim:42-122) ----
def pubkey_from_bytes(raw, group_check=True):
    """(BLST_ERROR, 96-byte affine) of PublicKey.fromBytes on 48 (compressed) or 96 (serialized) bytes."""
    out = _out(96)
    err = ref.ref_pubkey_from_bytes(_buf(raw), C.c_size_t(len(raw)), C.c_int(1 if group_check else 0), out)
    return int(err), bytes(out)


def signature_from_bytes(raw, group_check=True):
    out = _out(192)
    err = ref.ref_signature_from_bytes(_buf(raw), C.c_size_t(len(raw)), C.c_int(1 if group_check else 0), out)
    return int(err), bytes(out)


def g1_serialize(aff96):
    c, s = _out(48), _out(96)
    ref.ref_g1_compress(_buf(aff96), c, s)
    return bytes(s)


def g2_serialize(aff192):
    c, s = _out(96), _out(192)
    ref.ref_g2_compress(_buf(aff192), c, s)
    return bytes(s)


def sign(seed, msg, dst=DST_ETH2):
    """(pk 96 B, sig 192 B) of the key derived from `seed` over an arbitrary-length message under `dst`."""
    pk, sig = _out(96), _out(192)
    ref.ref_sign(C.c_uint64(seed), _buf(msg) if msg else None, C.c_size_t(len(msg)), _buf(dst), C.c_size_t(len(dst)), pk, sig)
    return bytes(pk), bytes(sig)
