"""GPU debug: small-route hash (map kernel + cofactor program) vs BLST hash_to_g2, and vs the same program run on the CPU."""
import ctypes as C, sys
sys.path.insert(0, '.')
import nim_blscurve_b200 as bg
from oracle import blst_ref as br, pyref as pr
L = bg.lib()
n = 3
c = bg.BatchedBLSVerifierCache(max_sets=64)
sets = br.make_sets(0, n)
hin = (C.c_uint8 * (n * 288))(); hout = (C.c_uint8 * (n * 288))()
print("rc", L.blsgpu_test_small_hash(c.handle, sets, n, hin, hout))
hs = C.CDLL('tests/hostsim/libhostsim.so')
P = pr.P
def hom(raw):
    v = [pr.fp_from_mont_bytes(raw[48*i:48*i+48]) for i in range(6)]
    z = (v[4], v[5])
    if z == (0, 0): return None
    zi = pr.f2_inv(z)
    return pr.f2_mul((v[0], v[1]), zi), pr.f2_mul((v[2], v[3]), zi)
for i in range(n):
    msg = sets[i*320+96:i*320+128]
    ref = pr.g2_from_mem(br.hash_to_g2(msg, 32, br.DST_ETH2)[1][:192]) if hasattr(br, 'DST_ETH2') else None
    a_in = hom(bytes(hin)[i*288:(i+1)*288]); a_out = hom(bytes(hout)[i*288:(i+1)*288])
    exp = pr.g2_clear_cofactor(a_in)
    st = (C.c_int*4)(); o = (C.c_uint8*288)()
    hs.hs_prog_g2_clear_cofactor((C.c_uint8*288).from_buffer_copy(bytes(hin)[i*288:(i+1)*288]), o, st)
    print(i, "in on curve:", pr.g2_on_curve(a_in), "| device out == pyref(clear(in)):", a_out == exp, "| hostsim prog == pyref:", hom(bytes(o)) == exp,
          "| device raw == hostsim raw:", bytes(o) == bytes(hout)[i*288:(i+1)*288], "| pyref hash:", pr.hash_to_g2(msg) == exp)
