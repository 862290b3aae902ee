"""One pass over every route of the hot path, sized for compute-sanitizer (tools/gpu_sanitize.sh).
usage: python tools/sanitize_run.py [--host] [--chunks C] n1 n2 ...   (expects every batch to verify)"""
import ctypes as C, hashlib, sys
sys.path.insert(0, '.')
import torch
import nim_blscurve_b200 as bg
L = bg.lib()
srb = hashlib.sha256(b"Mr F was here").digest()
argv = sys.argv[1:]
host = False
chunks = 4
while argv and argv[0].startswith("--"):
    if argv[0] == "--host": host = True; argv = argv[1:]
    elif argv[0] == "--chunks": chunks = int(argv[1]); argv = argv[2:]
    else: raise SystemExit("unknown flag " + argv[0])
sizes = [int(x) for x in argv] or [129]
cap = max(sizes)
c = bg.BatchedBLSVerifierCache(max_sets=cap)
d = torch.empty(cap * 320, dtype=torch.uint8, device='cuda')
assert L.blsgpu_make_sets(c.handle, 7, 0, cap, C.c_void_p(d.data_ptr()), 1) == 0
torch.cuda.synchronize()
h = bytes(d.cpu().numpy().tobytes()) if host else None
ok = True
for n in sizes:
    gt = (C.c_uint8 * 576)()
    if host:
        rc = L.blsgpu_batch_verify(c.handle, h[:320 * n], n, srb, chunks, None, gt)
    else:
        rc = L.blsgpu_batch_verify_dev(c.handle, C.c_void_p(d.data_ptr()), n, srb, chunks, None, gt)
    print(f"n={n} host={host} rc={rc} launches={L.blsgpu_last_launches(c.handle)}", flush=True)
    ok &= rc == 1
if not host:
    # companion entry points on a small input: G1 MSM, one pairing check, aggregate
    n = 600
    sets = d[:320 * n].cpu().numpy().tobytes()
    pts = b"".join(sets[320 * i:320 * i + 96] for i in range(n))
    sc = bytes((i * 131 + b * 17 + 3) & (0x7f if b == 31 else 0xff) for i in range(n) for b in range(32))
    r = bg.msmG1(c, pts, sc)
    print("msm_g1", len(r), flush=True)
    a = bg.aggregateAll(c, [pts[96 * i:96 * i + 96] for i in range(64)])
    print("aggregate", a[0], len(a[1]), flush=True)
sys.exit(0 if ok else 3)
