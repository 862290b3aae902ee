"""torchrun target: BASELINE configs[3] on real GPUs — one batch of N distinct-message sets (default 32 768) sharded
over WORLD_SIZE ranks, NCCL all-gather of the 576-byte partials (nim_blscurve_b200.batch_verify_distributed).

Checks, on every rank: (1) the valid batch verifies; (2) with one corrupted set the verdict is false and the GT equals
the one a single context computes for the whole batch on rank 0's GPU (independent of the number of ranks);
(3) when oracle/_ref is present, rank 0 also compares that GT with BLST's for a 2 048-set prefix batch; (4) a G1 MSM
of 16 384 points sharded over the ranks (msm_g1_distributed: 96 bytes per rank over NCCL) equals the single-context MSM on
every rank and blst_p1s_mult_pippenger on rank 0.

  python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 \
      tools/nccl_parity.py [N]
"""
import ctypes as C
import hashlib
import os
import sys

sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), ".."))
import torch
import torch.distributed as dist

import nim_blscurve_b200 as bg


def main():
    n = int(sys.argv[1]) if len(sys.argv) > 1 else 32768
    rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
    torch.cuda.set_device(local)
    dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    L = bg.lib()
    srb = hashlib.sha256(b"Mr F was here").digest()
    chunks = 64
    cache = bg.BatchedBLSVerifierCache(max_sets=n, device=local)
    be = bg.GpuBackend(cache)
    # every rank generates the whole synthetic batch (deterministic in (seed, index)) and keeps its share
    out = (C.c_uint8 * (320 * n))()
    assert L.blsgpu_make_sets(cache.handle, 4242, 0, n, out, 0) == 0
    sets = bytes(out)
    bad = bytearray(sets)
    victim = (n * 5) // 7
    bad[victim * 320 + 96] ^= 0x01            # message of one set
    bad = bytes(bad)
    first, cnt = bg.shard_range(n, world, rank)
    mine = lambda s: s[first * 320:(first + cnt) * 320]

    ok = bg.batch_verify_distributed(be, mine(sets), first, n, srb, chunks)
    assert ok is True, f"rank {rank}: valid batch rejected"
    ok_bad, gt_bad = bg.batch_verify_distributed(be, mine(bad), first, n, srb, chunks, want_gt=True)
    assert ok_bad is False, f"rank {rank}: corrupted batch accepted"
    # single-context result for the whole batch (this rank's GPU): must be identical
    one_ok, one_gt = cache.verify_raw(bad, srb, chunks, want_gt=True)
    assert (one_ok, one_gt) == (ok_bad, gt_bad), f"rank {rank}: sharded GT differs from the single-context GT"
    ref = "skipped"
    if rank == 0:
        try:
            from oracle import blst_ref as br
            m = 2048
            pre = bad[:m * 320] if victim < m else bad[(victim - 7) * 320:(victim - 7 + m) * 320]
            rok, rgt = br.batch_verify(pre, srb, 16)
            gok, ggt = cache.verify_raw(pre, srb, 16, want_gt=True)
            assert (gok, ggt) == (rok, rgt)
            ref = "BLST GT identical on a %d-set window around the corrupted set" % m
        except ImportError as ex:
            ref = f"oracle unavailable: {ex}"
    # SURVEY.md §8e MSM row: the G1 MSM sharded the same way, one 96-byte point per rank over NCCL
    msm = "skipped (oracle unavailable)"
    try:
        from oracle import blst_ref as br
        m = 16384
        pts, sc = br.msm_points(0xFACADE, m, 8)
        f, c = bg.shard_range(m, world, rank)
        got = bg.msm_g1_distributed(be, pts[f * 96:(f + c) * 96], sc[f * 32:(f + c) * 32], 255)
        assert got == bg.msmG1(cache, pts, sc, 255), f"rank {rank}: sharded MSM differs from the single-context MSM"
        if rank == 0:
            assert got == br.msm_g1(pts, sc, 255), "sharded MSM differs from blst_p1s_mult_pippenger"
        msm = f"G1 MSM of {m} points over {world} ranks == single context == BLST ({got[:8].hex()}...)"
    except ImportError:
        pass
    dist.barrier()
    if rank == 0:
        print(f"nccl_parity MSM: {msm}", flush=True)
        print(f"nccl_parity OK: {n} sets over {world} ranks, verdicts (True, False), sharded GT == single-context GT "
              f"({gt_bad[:8].hex()}...); {ref}", flush=True)
    dist.destroy_process_group()


if __name__ == "__main__":
    main()
