# blscurve/cuda/bls_batch_verifier_cuda.nim — drop-in bodies for blscurve/bls_batch_verifier.nim.
#
# Public names, parameter lists and results are those of the reference (file:line into nim-blscurve):
#   SignatureSet :34, BatchedBLSVerifierCache :62, init :108/:115, batchVerifySerial :121/:162,
#   batchVerifyParallel :296/:373/:399, batchVerify :420/:449/:475, aggregateAll / subtractAll (blst_min_pubkey_sig_core.nim:179, :197).
# `tp: Taskpool` stays in the signatures; it no longer fans work out — tp.numThreads only selects the reference's
# RLC-scalar chunking, so verdicts (and GT values) are identical to the BLST path for the same tp.
# NOT compiled in the build container (no Nim toolchain); see INTEGRATION.md.
{.push raises: [].}

import taskpools
import ../blst/blst_min_pubkey_sig_core   # PublicKey, Signature (in-memory blst_p1_affine / blst_p2_affine)
import ./blsgpu_abi

type
  SignatureSet* = tuple[pubkey: PublicKey, message: array[32, byte], signature: Signature]

  BatchedBLSVerifierCache* {.requiresInit.} = object
    ## device scratch instead of per-thread pairing contexts
    ctx: BlsGpuCtx
    numThreads: int

static: doAssert sizeof(SignatureSet) == 320   # pk 96 | msg 32 | sig 192, no padding

proc `=destroy`(c: var BatchedBLSVerifierCache) =
  if not pointer(c.ctx).isNil: blsgpu_destroy(c.ctx)

func init*(T: type BatchedBLSVerifierCache, maxSets = 16384, device = 0): T =
  T(ctx: blsgpu_create(device.cint, maxSets.csize_t), numThreads: 1)

func init*(T: type BatchedBLSVerifierCache, tp: Taskpool, maxSets = 16384, device = 0): T =
  T(ctx: blsgpu_create(device.cint, maxSets.csize_t), numThreads: tp.numThreads)

func verifyRaw(cache: var BatchedBLSVerifierCache, sets: ptr UncheckedArray[SignatureSet], n: int,
               srb: ptr array[32, byte], chunks: uint32): bool =
  if n == 0: return false                                    # spec precondition (:137, :312)
  doAssert not pointer(cache.ctx).isNil, "blsgpu_create failed: no CUDA device (there is no CPU fallback)"
  # 1 valid / 0 invalid / <0 CUDA failure -> false (+ blsgpu_last_error for diagnostics)
  blsgpu_batch_verify(cache.ctx, sets, n.csize_t, srb, chunks, nil, nil) == 1

func batchVerifySerial*(cache: var BatchedBLSVerifierCache, input: openArray[SignatureSet],
                        secureRandomBytes: array[32, byte]): bool =
  if input.len == 0: return false
  cache.verifyRaw(cast[ptr UncheckedArray[SignatureSet]](unsafeAddr input[0]), input.len,
                  unsafeAddr secureRandomBytes, 0'u32)

proc batchVerifyParallel*(tp: Taskpool, cache: var BatchedBLSVerifierCache, input: openArray[SignatureSet],
                          secureRandomBytes: array[32, byte]): bool =
  if input.len == 0: return false
  cache.verifyRaw(cast[ptr UncheckedArray[SignatureSet]](unsafeAddr input[0]), input.len,
                  unsafeAddr secureRandomBytes, uint32 max(1, tp.numThreads))

proc batchVerify*(tp: Taskpool, cache: var BatchedBLSVerifierCache, input: openArray[SignatureSet],
                  secureRandomBytes: array[32, byte]): bool =
  if tp.numThreads > 1 and input.len >= 3:                   # same rule as :440, :468
    tp.batchVerifyParallel(cache, input, secureRandomBytes)
  else:
    cache.batchVerifySerial(input, secureRandomBytes)

proc batchVerify*(tp: Taskpool, input: openArray[SignatureSet], secureRandomBytes: array[32, byte]): bool =
  var cache = BatchedBLSVerifierCache.init(tp, maxSets = max(1, input.len))
  tp.batchVerify(cache, input, secureRandomBytes)

type
  MultiSignatureSet* = object                                # bls_batch_verifier.nim:47-62
    pubkeys: seq[PublicKey]
    message: array[32, byte]
    signatures: seq[Signature]

func init*(T: type MultiSignatureSet, pubkeys: seq[PublicKey], message: array[32, byte],
           signatures: seq[Signature]): MultiSignatureSet =
  doAssert pubkeys.len == signatures.len
  doAssert pubkeys.len > 0
  MultiSignatureSet(pubkeys: pubkeys, message: message, signatures: signatures)

func init*(T: type MultiSignatureSet, sigset: SignatureSet): MultiSignatureSet =
  MultiSignatureSet(pubkeys: @[sigset.pubkey], message: sigset.message, signatures: @[sigset.signature])

func add*(multiSet: var MultiSignatureSet, sigset: SignatureSet) =
  doAssert multiSet.message == sigset.message
  multiSet.pubkeys.add sigset.pubkey
  multiSet.signatures.add sigset.signature

func combine*(cache: var BatchedBLSVerifierCache, multiSet: MultiSignatureSet,
              secureRandomBytes: array[32, byte]): SignatureSet =
  ## bls_batch_verifier.nim:100-106 / blst_min_pubkey_sig_core.nim:570-647 on the device
  doAssert multiSet.pubkeys.len > 0, "Must provide at least 1 signature"
  result.message = multiSet.message
  let rc = blsgpu_combine(cache.ctx, unsafeAddr secureRandomBytes, unsafeAddr multiSet.pubkeys[0],
                          unsafeAddr multiSet.signatures[0], multiSet.pubkeys.len.csize_t,
                          addr result.pubkey, addr result.signature)
  doAssert rc == 1, "blsgpu_combine failed"

func aggregateAll*(cache: var BatchedBLSVerifierCache, dst: var PublicKey, elems: openArray[PublicKey]): bool =
  if elems.len == 0: return false                            # blst_min_pubkey_sig_core.nim:183-184
  blsgpu_aggregate_g1(cache.ctx, unsafeAddr elems[0], elems.len.csize_t, addr dst) == 1

func aggregateAll*(cache: var BatchedBLSVerifierCache, dst: var Signature, elems: openArray[Signature]): bool =
  if elems.len == 0: return false
  blsgpu_aggregate_g2(cache.ctx, unsafeAddr elems[0], elems.len.csize_t, addr dst) == 1

proc subtractAll*(cache: var BatchedBLSVerifierCache, dst: var PublicKey, elems: openArray[PublicKey]) =
  ## dst <- dst - sum(elems)                                   # blst_min_pubkey_sig_core.nim:197-209
  if elems.len == 0: return
  doAssert blsgpu_subtract_g1(cache.ctx, addr dst, unsafeAddr elems[0], elems.len.csize_t) == 1, "blsgpu_subtract_g1 failed"

proc subtractAll*(cache: var BatchedBLSVerifierCache, dst: var Signature, elems: openArray[Signature]) =
  if elems.len == 0: return
  doAssert blsgpu_subtract_g2(cache.ctx, addr dst, unsafeAddr elems[0], elems.len.csize_t) == 1, "blsgpu_subtract_g2 failed"

# --- SURVEY §8f N3: bls_sig_min_pubkey.nim:108-258 on the device (proof-of-possession overloads stay as they are:
# they call popVerify per key and then these) ---
const DST = "BLS_SIG_BLS12381G2_XMD:SHA-256_SSWU_RO_POP_"        # bls_sig_min_pubkey.nim:31

func verify*[T: byte|char](cache: var BatchedBLSVerifierCache, publicKey: PublicKey, message: openArray[T],
                           signature: Signature): bool =
  ## coreVerifyNoGroupCheck (blst_min_pubkey_sig_core.nim:264-297) = aggregateVerify over one pair
  var offs = [0'u32, uint32 message.len]
  blsgpu_aggregate_verify(cache.ctx, unsafeAddr publicKey, 1, (if message.len > 0: cast[ptr byte](unsafeAddr message[0]) else: nil),
                          addr offs[0], cast[ptr byte](unsafeAddr DST[0]), DST.len.csize_t, unsafeAddr signature, nil) == 1

func fastAggregateVerify*[T: byte|char](cache: var BatchedBLSVerifierCache, publicKeys: openArray[PublicKey],
                                        message: openArray[T], signature: Signature): bool =
  if publicKeys.len == 0: return false                       # :251-253
  blsgpu_fast_aggregate_verify(cache.ctx, unsafeAddr publicKeys[0], publicKeys.len.csize_t,
                               (if message.len > 0: cast[ptr byte](unsafeAddr message[0]) else: nil), message.len.csize_t,
                               cast[ptr byte](unsafeAddr DST[0]), DST.len.csize_t, unsafeAddr signature, nil) == 1

func aggregateVerify*(cache: var BatchedBLSVerifierCache, publicKeys: openArray[PublicKey],
                      messages: openArray[seq[byte]], signature: Signature): bool =
  if publicKeys.len != messages.len or publicKeys.len < 1: return false     # :164-169
  var blob: seq[byte]
  var offs = newSeq[uint32](messages.len + 1)
  for i, m in messages:
    blob.add m
    offs[i + 1] = uint32 blob.len
  blsgpu_aggregate_verify(cache.ctx, unsafeAddr publicKeys[0], publicKeys.len.csize_t,
                          (if blob.len > 0: addr blob[0] else: nil), addr offs[0],
                          cast[ptr byte](unsafeAddr DST[0]), DST.len.csize_t, unsafeAddr signature, nil) == 1

# --- SURVEY §8f N2: bls_sig_io.nim:42-122, batched ---
func fromBytes*(cache: var BatchedBLSVerifierCache, dst: var openArray[PublicKey], raw: openArray[array[48, byte]],
                ok: var openArray[bool]): bool =
  ## every element as PublicKey.fromBytes: uncompress, reject infinity, subgroup check
  doAssert dst.len == raw.len and ok.len == raw.len
  if raw.len == 0: return true
  var status = newSeq[byte](raw.len)
  result = blsgpu_pubkeys_from_bytes(cache.ctx, cast[ptr byte](unsafeAddr raw[0]), raw.len.csize_t, 48, 1,
                                     addr dst[0], addr status[0]) == 1
  for i in 0 ..< raw.len: ok[i] = status[i] == 0

func fromBytes*(cache: var BatchedBLSVerifierCache, dst: var openArray[Signature], raw: openArray[array[96, byte]],
                ok: var openArray[bool]): bool =
  doAssert dst.len == raw.len and ok.len == raw.len
  if raw.len == 0: return true
  var status = newSeq[byte](raw.len)
  result = blsgpu_signatures_from_bytes(cache.ctx, cast[ptr byte](unsafeAddr raw[0]), raw.len.csize_t, 96, 1,
                                        addr dst[0], addr status[0]) == 1
  for i in 0 ..< raw.len: ok[i] = status[i] == 0
{.pop.}
