"""Quick device-side stage timing probe (not the benchmark).
usage: python tools/probe.py [--chunks C] n1 n2 ...     (sets resident in HBM, blsgpu_batch_verify_dev, wall + stage times)"""
import ctypes as C, hashlib, sys, time
sys.path.insert(0, '.')
import torch
import nim_blscurve_b200 as bg
L = bg.lib()
srb = hashlib.sha256(b"Mr F was here").digest()
argv = sys.argv[1:]
chunks = 16
if argv and argv[0] == "--chunks":
    chunks = int(argv[1]); argv = argv[2:]
sizes = [int(x) for x in argv] or [64, 129, 4096, 32768, 131072]
cap = max(sizes)
c = bg.BatchedBLSVerifierCache(max_sets=cap)
print("imad peak wide: %.3e /s   lo: %.3e /s   chunks=%d" % (L.blsgpu_imad_peak(c.handle, 1), L.blsgpu_imad_peak(c.handle, 0), chunks))
d = torch.empty(cap * 320, dtype=torch.uint8, device='cuda')
t = time.time(); rc = L.blsgpu_make_sets(c.handle, 7, 0, cap, C.c_void_p(d.data_ptr()), 1); torch.cuda.synchronize()
print("make_sets", cap, rc, "%.3fs" % (time.time() - t))
names = [L.blsgpu_stage_name(i).decode() for i in range(16)]
names = [x for x in names if x]
for n in sizes:
    best = 1e9
    for rep in range(4):
        gt = (C.c_uint8 * 576)()
        t = time.time()
        rc = L.blsgpu_batch_verify_dev(c.handle, C.c_void_p(d.data_ptr()), n, srb, chunks, None, gt)
        dt = time.time() - t
        if rep: best = min(best, dt)
    ms = (C.c_float * 16)(); L.blsgpu_last_stage_ms(c.handle, ms, 16)
    print(f"n={n} rc={rc} wall={best*1e3:.2f}ms sets/s={n/best:.0f} launches={L.blsgpu_last_launches(c.handle)}")
    print("   " + " ".join(f"{nm}={v:.3f}" for nm, v in zip(names, ms)))
