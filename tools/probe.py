"""Quick device-side stage timing probe (not the benchmark)."""
import ctypes as C, hashlib, sys, time
sys.path.insert(0, '.')
import torch
import nim_blscurve_b200 as bg
L = bg.lib()
srb = hashlib.sha256(b"Mr F was here").digest()
sizes = [int(x) for x in sys.argv[1:]] or [64, 129, 4096, 32768, 131072]
cap = max(sizes)
c = bg.BatchedBLSVerifierCache(max_sets=cap)
print("imad peak wide: %.3e /s   lo: %.3e /s" % (L.blsgpu_imad_peak(c.handle, 1), L.blsgpu_imad_peak(c.handle, 0)))
d = torch.empty(cap * 320, dtype=torch.uint8, device='cuda')
t = time.time(); rc = L.blsgpu_make_sets(c.handle, 7, 0, cap, C.c_void_p(d.data_ptr()), 1); torch.cuda.synchronize()
print("make_sets", cap, rc, "%.3fs" % (time.time() - t))
names = [L.blsgpu_stage_name(i).decode() for i in range(16)]
names = [x for x in names if x]
for n in sizes:
    for rep in range(2):
        gt = (C.c_uint8 * 576)()
        t = time.time()
        rc = L.blsgpu_batch_verify_dev(c.handle, C.c_void_p(d.data_ptr()), n, srb, 1024, None, gt)
        dt = time.time() - t
    ms = (C.c_float * 16)(); L.blsgpu_last_stage_ms(c.handle, ms, 16)
    print(f"n={n} rc={rc} wall={dt*1e3:.2f}ms sets/s={n/dt:.0f} launches={L.blsgpu_last_launches(c.handle)}")
    print("   " + " ".join(f"{nm}={v:.3f}" for nm, v in zip(names, ms)))
