"""G1 MSM timing probe (device-resident inputs): ms per call for n = 2^k."""
import ctypes as C, sys, time
sys.path.insert(0, '.')
import torch
import nim_blscurve_b200 as bg
L = bg.lib()
ks = [int(x) for x in sys.argv[1:]] or [16, 18, 20]
c = bg.BatchedBLSVerifierCache(max_sets=16)
for k in ks:
    n = 1 << k
    dp = torch.empty(n * 96, dtype=torch.uint8, device='cuda')
    ds = torch.empty(n * 32, dtype=torch.uint8, device='cuda')
    assert L.blsgpu_msm_make_inputs(c.handle, 0xFACADE, n, C.c_void_p(dp.data_ptr()), C.c_void_p(ds.data_ptr())) == 0
    out = (C.c_uint8 * 96)()
    best = 1e9
    for rep in range(4):
        torch.cuda.synchronize(); t = time.perf_counter()
        rc = L.blsgpu_msm_g1_dev(c.handle, C.c_void_p(dp.data_ptr()), C.c_void_p(ds.data_ptr()), n, 255, out)
        dt = time.perf_counter() - t
        if rep: best = min(best, dt)
    print(f"msm n=2^{k} rc={rc} best={best*1e3:.2f} ms  out[:8]={bytes(out)[:8].hex()}", flush=True)
