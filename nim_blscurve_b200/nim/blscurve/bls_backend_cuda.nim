# blscurve/bls_backend.nim with the CUDA branch — the reference's backend switch (bls_backend.nim:10-33) extended by one
# value: `-d:BLS_FORCE_BACKEND=cuda` keeps BLST for everything that is not on the batch-verification path (signing, key
# derivation, (de)serialisation, single pairing checks) and swaps the batch verifier for blscurve/cuda.
#
# Drop-in: this file REPLACES blscurve/bls_backend.nim; blscurve/bls_public_exports.nim then selects the verifier with
#
#   when BLS_BACKEND == CUDA:
#     import ./cuda/bls_batch_verifier_cuda
#     export bls_batch_verifier_cuda
#   else:
#     import ./bls_batch_verifier
#     export bls_batch_verifier
#
# NOT compiled in the build container (no Nim toolchain there).  See INTEGRATION.md section 1.

const BLS_FORCE_BACKEND*{.strdefine.} = "auto"

static: doAssert BLS_FORCE_BACKEND == "auto" or
                 BLS_FORCE_BACKEND == "blst" or
                 BLS_FORCE_BACKEND == "cuda",
                 """Only "auto", "blst" and "cuda" backends are valid."""

type BlsBackendKind* = enum
  BLST
  CUDA   ## BLST for the scalar paths + libblsgpu.so (B200, sm_100a) for batchVerify*/aggregateAll/combine

# BLST is compiled in either way: the CUDA backend replaces the batch verifier, not the library
# (PublicKey / Signature stay blst_p1_affine / blst_p2_affine, which is what crosses the C ABI byte for byte).
const UseBLST = true
const UseCUDA = BLS_FORCE_BACKEND == "cuda"

when UseBLST:
  when defined(amd64) or defined(arm64):
    # BLST has assembly routines and detects the most profitable one at runtime
    # when `__BLST_PORTABLE__` is set
    {.passc: "-D__BLST_PORTABLE__".}
  else:
    # WASM and others - no specialised assembly code available
    {.passc: "-D__BLST_NO_ASM__".}

when UseCUDA:
  const BLS_BACKEND* = CUDA
else:
  const BLS_BACKEND* = BLST

import ./blst/[blst_min_pubkey_sig_core, blst_recovery]
export blst_min_pubkey_sig_core, blst_recovery

when BLS_BACKEND == CUDA:
  # aggregateAll / subtractAll / combine of blst_min_pubkey_sig_core are shadowed by the device versions for callers that
  # import the batch verifier module; the FFI itself:
  import ./cuda/blsgpu_abi
  export blsgpu_abi
