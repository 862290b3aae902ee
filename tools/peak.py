import sys
sys.path.insert(0, '.')
import nim_blscurve_b200 as bg
L = bg.lib(); c = bg.BatchedBLSVerifierCache(max_sets=16)
print("imad wide %.3e lo %.3e" % (L.blsgpu_imad_peak(c.handle, 1), L.blsgpu_imad_peak(c.handle, 0)))
for tpb, bps in [(32, 4), (64, 4), (128, 4), (128, 8), (256, 8), (128, 2), (128, 1)]:
    r = L.blsgpu_fpmul_peak(c.handle, tpb, bps)
    print(f"fpmul peak tpb={tpb} blocks/SM={bps} warps/SMSP={tpb*bps/128:.1f}: {r:.3e} Fp-mul/s = {r*300:.3e} imad/s")
