import ctypes as C, hashlib, sys
sys.path.insert(0, '.')
import torch
import nim_blscurve_b200 as bg
L = bg.lib()
srb = hashlib.sha256(b"Mr F was here").digest()
S = int(sys.argv[1]) if len(sys.argv) > 1 else 4096
seed = int(sys.argv[2]) if len(sys.argv) > 2 else 2026
c = bg.BatchedBLSVerifierCache(max_sets=S)
h = c.handle
d = torch.empty(S * 320, dtype=torch.uint8, device='cuda')
assert L.blsgpu_make_sets(h, seed, 0, S, C.c_void_p(d.data_ptr()), 1) == 0
gt = (C.c_uint8 * 576)()
print("batch_verify_dev:", L.blsgpu_batch_verify_dev(h, C.c_void_p(d.data_ptr()), S, srb, 1024, None, gt), bytes(gt)[:8].hex())
dp = torch.zeros(576, dtype=torch.uint8, device='cuda'); df = torch.zeros(1, dtype=torch.int32, device='cuda')
torch.cuda.synchronize()
rc = L.blsgpu_partial_dev(h, C.c_void_p(d.data_ptr()), S, 0, S, srb, 1024, C.c_void_p(dp.data_ptr()), C.c_void_p(df.data_ptr()))
print("partial_dev rc", rc, c.last_error())
rc = L.blsgpu_finalize_dev(h, C.c_void_p(dp.data_ptr()), 1, C.c_void_p(df.data_ptr()), gt)
print("finalize_dev rc", rc, c.last_error(), bytes(gt)[:8].hex(), "flag", df.cpu().tolist())
