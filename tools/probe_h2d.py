"""Where does the host-buffer call spend its extra time?  H2D of 131072 x 320 B from pinned and pageable memory,
alone (torch copy, CUDA events) and inside blsgpu_batch_verify (wall clock vs the resident call)."""
import ctypes as C, hashlib, sys, time
sys.path.insert(0, '.')
import torch
import nim_blscurve_b200 as bg
L = bg.lib()
srb = hashlib.sha256(b"Mr F was here").digest()
n = int(sys.argv[1]) if len(sys.argv) > 1 else 131072
c = bg.BatchedBLSVerifierCache(max_sets=n)
d = torch.empty(n * 320, dtype=torch.uint8, device='cuda')
assert L.blsgpu_make_sets(c.handle, 7, 0, n, C.c_void_p(d.data_ptr()), 1) == 0
hp = d.cpu().pin_memory()
hg = bytearray(hp.numpy().tobytes())
cg = (C.c_uint8 * len(hg)).from_buffer(hg)
d2 = torch.empty_like(d)
for name, src in (("pinned", hp), ("pageable", torch.frombuffer(hg, dtype=torch.uint8))):
    for _ in range(2):
        d2.copy_(src, non_blocking=True); torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    t0 = time.perf_counter(); e0.record()
    for _ in range(5):
        d2.copy_(src, non_blocking=True)
    e1.record(); torch.cuda.synchronize()
    print(f"torch H2D {name}: {e0.elapsed_time(e1)/5:.3f} ms/copy (events), {(time.perf_counter()-t0)/5*1e3:.3f} ms wall  -> {n*320/ (e0.elapsed_time(e1)/5*1e-3)/1e9:.1f} GB/s")
gt = (C.c_uint8 * 576)()
def wall(fn, reps=5):
    fn(); fn()
    t0 = time.perf_counter()
    for _ in range(reps): assert fn() == 1
    return (time.perf_counter() - t0) / reps * 1e3
names = [L.blsgpu_stage_name(i).decode() for i in range(16)]
names = [x for x in names if x]
def stages():
    ms = (C.c_float * 16)(); L.blsgpu_last_stage_ms(c.handle, ms, 16)
    return " ".join(f"{nm}={v:.2f}" for nm, v in zip(names, ms))
ds = torch.empty_like(d)
def torch_then_dev():
    ds.copy_(hp, non_blocking=True)
    return L.blsgpu_batch_verify_dev(c.handle, C.c_void_p(ds.data_ptr()), n, srb, 16, None, gt)
st = torch.cuda.Stream()
torch.cuda.set_stream(st)
L.blsgpu_set_stream(c.handle, C.c_void_p(st.cuda_stream))
print("torch copy + dev (one stream): %.2f ms" % wall(torch_then_dev)); print("   ", stages())
print("resident  : %.2f ms" % wall(lambda: L.blsgpu_batch_verify_dev(c.handle, C.c_void_p(d.data_ptr()), n, srb, 16, None, gt))); print("   ", stages())
print("pinned    : %.2f ms" % wall(lambda: L.blsgpu_batch_verify(c.handle, C.c_void_p(hp.data_ptr()), n, srb, 16, None, gt))); print("   ", stages())
print("pageable  : %.2f ms" % wall(lambda: L.blsgpu_batch_verify(c.handle, cg, n, srb, 16, None, gt)))
print("pinned    : %.2f ms" % wall(lambda: L.blsgpu_batch_verify(c.handle, C.c_void_p(hp.data_ptr()), n, srb, 16, None, gt)))
print("resident  : %.2f ms" % wall(lambda: L.blsgpu_batch_verify_dev(c.handle, C.c_void_p(d.data_ptr()), n, srb, 16, None, gt)))
