"""CPU: the multi-rank host logic (nim_blscurve_b200/multi_gpu.py) with world_size 2 and 3 over gloo.  The device
call is replaced by the oracle (BLST) computing each rank's 576-byte partial, so what is tested is the sharding
rule, the exchange and the single final exponentiation — the result must equal the one-context BLST verdict + GT."""
import hashlib
import os
import socket

import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


class OracleBackend:
    def partial(self, sets, first, total_n, srb, chunks):
        from oracle import blst_ref as br
        return br.partial(sets, first, total_n, srb, chunks)

    def finalize(self, partials):
        from oracle import blst_ref as br
        return br.finalize(partials)

    def msm_g1(self, points96, scalars, nbits):
        from oracle import blst_ref as br
        return br.msm_g1(points96, scalars, nbits)

    def aggregate_g1(self, points96):
        from oracle import blst_ref as br
        ok, pt = br.aggregate_g1(points96)
        return pt if ok else bytes(96)


def _worker(rank, world, port, sets, srb, chunks, q):
    import sys
    sys.path.insert(0, ROOT)
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    import nim_blscurve_b200 as bg
    total = len(sets) // 320
    first, cnt = bg.shard_range(total, world, rank)
    res = bg.batch_verify_distributed(OracleBackend(), sets[first * 320:(first + cnt) * 320], first, total, srb, chunks,
                                      want_gt=True)
    q.put((rank, res[0], res[1]))
    dist.barrier()
    dist.destroy_process_group()


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _run(world, sets, srb, chunks):
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    ps = [ctx.Process(target=_worker, args=(r, world, port, sets, srb, chunks, q)) for r in range(world)]
    for p in ps:
        p.start()
    out = [q.get(timeout=120) for _ in range(world)]
    for p in ps:
        p.join(timeout=60)
        assert p.exitcode == 0
    return sorted(out)


@pytest.mark.parametrize("world", [2, 3])
def test_distributed_matches_single_context(world):
    try:
        from oracle import blst_ref as br
    except Exception as e:
        pytest.skip(f"oracle/_ref not built: {e}")
    srb = hashlib.sha256(b"Mr F was here").digest()
    sets = br.make_sets(0, 11)
    bad = bytearray(sets)
    bad[9 * 320 + 128:10 * 320] = sets[128:320]
    infpk = bytearray(sets)
    infpk[10 * 320:10 * 320 + 96] = bytes(96)
    for s, chunks in ((sets, 4), (bytes(bad), 4), (bytes(bad), 0), (bytes(infpk), 4)):
        ok, gt = br.batch_verify(s, srb, chunks)
        for rank, rok, rgt in _run(world, s, srb, chunks):
            assert rok == ok, (world, rank)
            assert rgt == gt, (world, rank)


def _msm_worker(rank, world, port, pts, sc, q):
    import sys
    sys.path.insert(0, ROOT)
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    import nim_blscurve_b200 as bg
    first, cnt = bg.shard_range(len(pts) // 96, world, rank)
    res = bg.msm_g1_distributed(OracleBackend(), pts[first * 96:(first + cnt) * 96], sc[first * 32:(first + cnt) * 32], 255)
    q.put((rank, res))
    dist.barrier()
    dist.destroy_process_group()


@pytest.mark.parametrize("world", [2, 3])
def test_distributed_msm_matches_single_context(world):
    """SURVEY.md §8e MSM row: per-rank Pippenger over a slice, 96-byte all-gather, sum — equal to the MSM over all
    points, including a batch smaller than the world (empty shares) and a sum that cancels to infinity."""
    try:
        from oracle import blst_ref as br
    except Exception as e:
        pytest.skip(f"oracle/_ref not built: {e}")
    pts, sc = br.msm_points(0xFACADE, 37)
    one = (1).to_bytes(32, "little")
    r_minus_1 = (0x73eda753299d7d483339d80809a1d80553bda402fffe5bfeffffffff00000001 - 1).to_bytes(32, "little")
    cases = [(pts, sc), (pts[:96], sc[:32]), (pts[:96] * 2, one + r_minus_1)]        # P + (r-1)P = infinity
    for p, k in cases:
        want = br.msm_g1(p, k, 255)
        ctx = mp.get_context("spawn")
        q = ctx.Queue()
        port = _free_port()
        ps = [ctx.Process(target=_msm_worker, args=(r, world, port, p, k, q)) for r in range(world)]
        for x in ps:
            x.start()
        out = [q.get(timeout=120) for _ in range(world)]
        for x in ps:
            x.join(timeout=60)
            assert x.exitcode == 0
        assert all(res == want for _, res in out), (world, len(p) // 96)
    assert br.msm_g1(*cases[2], 255) == bytes(96)
