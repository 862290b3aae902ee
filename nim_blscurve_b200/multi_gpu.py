"""One batch over several GPUs: one process per GPU, each rank verifies a contiguous share and contributes a
single 576-byte Fp12 partial; the partials are all-gathered (NCCL over NVLink on GPUs, gloo in CPU tests) and one
final exponentiation decides the batch.

This replaces the Taskpools fan-out + log-tree merge of blscurve/bls_batch_verifier.nim:296-371
(processSingleChunk / reducePartialPairings): rank k plays the role of a group of the reference's chunks, the
all-gather + product plays the role of blst_pairing_merge (vendor/blst/src/aggregate.c:410-458).  RLC scalars are
derived from the GLOBAL (total_n, chunks) derivation, so the verdict and the GT value do not depend on the number
of ranks.
"""
import ctypes as C
from typing import Callable, Optional, Tuple

from ._lib import BlsGpuError, lib


def shard_range(total: int, world: int, rank: int) -> Tuple[int, int]:
    """Balanced contiguous split (sizes differ by at most one) — the rule of blscurve/parallel_chunks.nim:42-55.
    Returns (first, count)."""
    base, rem = divmod(total, world)
    if rank < rem:
        return (base + 1) * rank, base + 1
    return base * rank + rem, base


class GpuBackend:
    """partial/finalize through the C ABI on this rank's device (host buffers)."""

    def __init__(self, cache):
        self.cache = cache

    def partial(self, sets: bytes, first: int, total_n: int, srb: bytes, chunks: int):
        n = len(sets) // 320
        out = (C.c_uint8 * 576)()
        flag = C.c_int(0)
        rc = lib().blsgpu_partial(self.cache.handle, sets if n else None, 0, n, first, total_n, srb, chunks, None, out,
                                  C.byref(flag))
        if rc < 0:
            raise BlsGpuError(f"blsgpu_partial failed ({rc}): {self.cache.last_error()}")
        return bytes(out), int(flag.value)

    def finalize(self, partials: bytes):
        gt = (C.c_uint8 * 576)()
        rc = lib().blsgpu_finalize(self.cache.handle, partials, len(partials) // 576, gt)
        if rc < 0:
            raise BlsGpuError(f"blsgpu_finalize failed ({rc}): {self.cache.last_error()}")
        return bool(rc), bytes(gt)


    def msm_g1(self, points96: bytes, scalars: bytes, nbits: int) -> bytes:
        from .batch_verifier import msmG1
        return msmG1(self.cache, points96, scalars, nbits)

    def aggregate_g1(self, points96: bytes) -> bytes:
        from .batch_verifier import aggregateAll
        ok, pt = aggregateAll(self.cache, [points96[i:i + 96] for i in range(0, len(points96), 96)])
        return pt if ok else bytes(96)


def msm_g1_distributed(backend, local_points96: bytes, local_scalars: bytes, nbits: int = 255, group=None) -> bytes:
    """Collective G1 MSM (MSM row): every rank passes its slice of the points and scalars
    (shard_range of the n inputs), runs the Pippenger MSM of blst_p1s_mult_pippenger (multi_scalar.c:415-434) on it
    and contributes ONE point; the points are all-gathered (96 bytes per rank: the affine form, so that an empty or
    cancelling share is the all-zero point) and summed on every rank.  Affine coordinates are unique, so the result
    is byte-identical to the one-context MSM whatever the number of ranks."""
    import torch
    import torch.distributed as dist

    world = dist.get_world_size(group) if dist.is_initialized() else 1
    n = len(local_points96) // 96
    assert len(local_scalars) * 96 == len(local_points96) * ((nbits + 7) // 8)
    mine = backend.msm_g1(local_points96, local_scalars, nbits) if n else bytes(96)
    if world == 1:
        return mine
    dev = "cuda" if dist.get_backend(group) == "nccl" else "cpu"
    t = torch.frombuffer(bytearray(mine), dtype=torch.uint8).to(dev)
    gathered = torch.empty(world * 96, dtype=torch.uint8, device=dev)
    dist.all_gather_into_tensor(gathered, t, group=group)
    return backend.aggregate_g1(gathered.cpu().numpy().tobytes())


def batch_verify_distributed(backend, local_sets: bytes, first: int, total_n: int, srb: bytes, chunks: int,
                             group=None, want_gt: bool = False):
    """Collective call: every rank passes its share (global indices [first, first+len)) and gets the verdict.

    Exchange = all_gather of 576 bytes + one flag per rank through torch.distributed (backend of `group`:
    nccl on GPUs, gloo on CPU).  Returns bool, or (bool, gt_bytes) with want_gt.
    """
    import torch
    import torch.distributed as dist

    if total_n == 0:                       # bls_batch_verifier.nim:312-314
        return (False, bytes(576)) if want_gt else False
    world = dist.get_world_size(group) if dist.is_initialized() else 1
    partial, flag = backend.partial(local_sets, first, total_n, srb, chunks)
    if world == 1:
        parts, flags = partial, [flag]
    else:
        dev = "cuda" if dist.get_backend(group) == "nccl" else "cpu"
        mine = torch.tensor(list(partial) + [flag, 0, 0, 0], dtype=torch.uint8, device=dev)   # 576 B + flag, 580 B
        gathered = torch.empty(world * 580, dtype=torch.uint8, device=dev)
        dist.all_gather_into_tensor(gathered, mine, group=group)
        g = bytes(gathered.cpu().numpy().tobytes())
        parts = b"".join(g[580 * r:580 * r + 576] for r in range(world))
        flags = [g[580 * r + 576] for r in range(world)]
    if any(flags):                         # a share hit an infinite public key: update() false -> batch false
        return (False, bytes(576)) if want_gt else False
    ok, gt = backend.finalize(parts)
    return (ok, gt) if want_gt else ok
