"""Latency of the single-call entry points on the GPU next to BLST on one host thread (the reference's
verify / fastAggregateVerify / aggregateVerify are single-threaded calls): wall-clock ms per call, best of 5."""
import hashlib, sys, time
sys.path.insert(0, '.')
import nim_blscurve_b200 as bg
from oracle import blst_ref as br

c = bg.BatchedBLSVerifierCache(max_sets=4096)
msg = hashlib.sha256(b"sync committee").digest()


def best(f, reps=5):
    f()
    b = 1e9
    for _ in range(reps):
        t = time.perf_counter(); r = f(); b = min(b, time.perf_counter() - t)
    return b * 1e3, r


for nkeys in (1, 128, 512):
    pks, aset = br.fast_aggregate_set(1000, nkeys, msg)
    sig = aset[128:320]
    keys = [pks[96 * i:96 * i + 96] for i in range(nkeys)]
    g, ok = best(lambda: bg.fastAggregateVerify(c, keys, msg, sig))
    h, rok = best(lambda: br.fast_aggregate_verify(pks, msg, sig)[0])
    print(f"fastAggregateVerify {nkeys:4d} keys: GPU {g:6.2f} ms ({ok})   BLST 1 thread {h:6.2f} ms ({rok})", flush=True)
for n in (2, 16, 128):
    sets = br.make_sets(77, n)
    # aggregateVerify: one signature = sum of the n signatures over n distinct messages
    pks = [sets[i * 320:i * 320 + 96] for i in range(n)]
    msgs = [sets[i * 320 + 96:i * 320 + 128] for i in range(n)]
    ok_agg, sig = bg.aggregateAll(c, [sets[i * 320 + 128:i * 320 + 320] for i in range(n)])
    g, ok = best(lambda: bg.aggregateVerify(c, pks, msgs, sig))
    h, rok = best(lambda: br.aggregate_verify(b"".join(pks), msgs, sig)[0])
    print(f"aggregateVerify     {n:4d} pairs: GPU {g:6.2f} ms ({ok})   BLST 1 thread {h:6.2f} ms ({rok})", flush=True)
