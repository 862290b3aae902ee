#!/usr/bin/env python3
"""bench.py — batch-verified signature sets/s on N B200 (one process per GPU).

Workload (config.workload): the per-GPU share of BASELINE.json configs[4] — "1M distinct-message signature
sets on 8xB200" = 131,072 sets per GPU (weak scaling: N GPUs verify ONE batch of N*131,072 sets).  A step is
one batch verification: every rank runs the per-set pipeline on its share (RLC scalars, hash_to_G2, [r]pk,
[r]sig, Miller loops, GT product), emits one 576-byte Fp12 partial, NCCL all-gathers the partials and ONE
final exponentiation decides the batch.  configs[1] (Eth2 block: 128 aggregate-key sets + one 512-key set)
is latency-bound at 129 sets; it is timed as an extra (`block_batch`) and covered by the parity tests.

  value  : sets/s with the sets resident in HBM (device-generated synthetic valid sets, distinct messages)
  e2e    : same through the host-buffer C-ABI call: pinned host sets -> H2D -> verify -> bool back
  --impl reference : the reference's own CPU path (BLST from oracle/_ref driven by oracle/ref_batch.c, the
                     pthreads replica of batchVerifyParallel) on all host cores, bounded sample per step.
"""
import argparse
import ctypes as C
import hashlib
import json
import os
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

SETS_PER_GPU = 131072
CHUNKS_PER_GPU = 1024          # reference chunk count (tp.numThreads) used for the RLC scalar derivation
W_SET = 12725                  # Fp-mul per set of the reference algorithm
W_BATCH = 14673                # per batch finalisation
IMAD_PER_FPMUL = 300
# reference-algorithm Fp-mul per set of each stage: A3, A5 (G1), A4, and A7 split into its line
# evaluations (63 x (25+4) + 5 x (35+4)) and its accumulation (68 x 39 + shared squarings)
STAGE_FPMUL = {"hash_to_g2": 4919, "g1_mul64": 800, "pairs_affine": 23, "miller_lines": 2022, "miller_acc": 2933}
# algorithmic bytes per set moved by each stage's kernel (inputs read + outputs written)
STAGE_IO_BYTES = {"hash_to_g2": 32 + 288, "g1_mul64": 96 + 8 + 144, "miller_lines": 192 + 96 + 68 * 288,
                  "miller_acc": 68 * 288 + 576 // 8}


def clocks_sampler(stop, out, gpu_index):
    q = "clocks.sm,clocks.max.sm,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown," \
        "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap"
    while not stop.is_set():
        try:
            r = subprocess.run(["nvidia-smi", "-i", str(gpu_index), f"--query-gpu={q}", "--format=csv,noheader,nounits"],
                               capture_output=True, text=True, timeout=5)
            f = [x.strip() for x in r.stdout.strip().split(",")]
            if len(f) >= 6:
                out.append(f)
        except Exception:
            pass
        stop.wait(0.2)


def summarize_clocks(samples):
    if not samples:
        return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["unavailable"]}
    sm = sorted(int(s[0]) for s in samples if s[0].isdigit())
    names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
    reasons = [n for i, n in enumerate(names) if any(s[2 + i].lower().startswith("active") for s in samples)]
    return {"sm_mhz": sm[len(sm) // 2] if sm else None,
            "sm_max_mhz": int(samples[0][1]) if samples[0][1].isdigit() else None, "reasons": reasons}


def run_reference(args):
    """The reference's CPU implementation of the path on this host's cores (rank 0 only)."""
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return 0
    from oracle import blst_ref as br
    cores = br.ncores()
    srb = hashlib.sha256(b"Mr F was here").digest()
    n = max(64, min(SETS_PER_GPU, 1200 * cores))           # ~1-2 s of CPU work per step
    sets = br.make_sets(0, min(n, 4096))                   # distinct valid sets; tiled to n (work is identical)
    sets = (sets * (n // (len(sets) // 320) + 1))[:n * 320]
    for _ in range(args.warmup):
        assert br.batch_verify_mt(sets, srb, cores)
    t0 = time.perf_counter()
    for _ in range(args.steps):
        ok = br.batch_verify_mt(sets, srb, cores)
        assert ok
    dt = time.perf_counter() - t0
    v = n * args.steps / dt
    line = {
        "impl": "reference", "metric": "batch-verified signature sets/sec", "value": v, "unit": "sets/s",
        "n_gpus": args.gpus, "steps": args.steps, "warmup": args.warmup, "ms_per_step": dt / args.steps * 1e3,
        "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "u32x12 (Fp 381-bit Montgomery)",
        "data": "synthetic", "gpu_launches": 0,
        "config": {"workload": f"{SETS_PER_GPU} distinct-message signature sets per GPU (BASELINE configs[4] share)",
                   "sets_per_step": n, "threads": cores},
        "cpu_baseline": {"value": v, "unit": "sets/s", "cores": cores, "kind": "reference",
                         "sample": f"{n} sets per step (<=4096 distinct sets tiled), BLST batchVerifyParallel replica "
                                   f"(pthreads, {cores} threads), {args.steps} steps"},
        "e2e": {"value": v, "unit": "sets/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
    }
    print(json.dumps(line), flush=True)
    return 0


def bench_msm(L, h, cache, stream, flush, peak_wide, hbm_peak, args):
    """BASELINE metric, second half: G1 MSM over 2^20 points, ms (benchmarks/bls12381_msm_g1.nim shape: 96-bit-multiple
    points, 255-bit scalars).  Inputs generated on the device; timed with CUDA events on the launching stream, L2 flushed
    between repetitions.  Roofline: reference-algorithm work W_pt = 176 Fp-mul per point at 2^20 x 300
    IMAD against the measured IMAD.WIDE peak; HBM bytes (points + scalars read once) as the secondary figure.
    CPU beside it: BLST blst_p1s_mult_pippenger, ONE thread (the reference benchmark is single-threaded), same inputs —
    which also makes this a full-size parity check of the affine result."""
    import torch
    k = args.msm_log2
    n = 1 << k
    dev = flush.device
    dp = torch.empty(n * 96, dtype=torch.uint8, device=dev)
    ds = torch.empty(n * 32, dtype=torch.uint8, device=dev)
    assert L.blsgpu_msm_make_inputs(h, 0xFACADE, n, C.c_void_p(dp.data_ptr()), C.c_void_p(ds.data_ptr())) == 0
    out = (C.c_uint8 * 96)()
    for _ in range(3):
        assert L.blsgpu_msm_g1_dev(h, C.c_void_p(dp.data_ptr()), C.c_void_p(ds.data_ptr()), n, 255, out) == 1
    reps, tot = 5, 0.0
    for _ in range(reps):
        flush.zero_()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(stream)
        L.blsgpu_msm_g1_dev(h, C.c_void_p(dp.data_ptr()), C.c_void_p(ds.data_ptr()), n, 255, out)
        e1.record(stream)
        torch.cuda.synchronize()
        tot += e0.elapsed_time(e1)
    ms = tot / reps
    launches = L.blsgpu_last_launches(h)
    # e2e: host points + scalars -> H2D -> MSM -> affine result back
    hp, hs = dp.cpu().pin_memory(), ds.cpu().pin_memory()
    t0 = time.perf_counter()
    assert L.blsgpu_msm_g1(h, C.c_void_p(hp.data_ptr()), C.c_void_p(hs.data_ptr()), n, 255, out) == 1
    t0 = time.perf_counter()
    for _ in range(3):
        L.blsgpu_msm_g1(h, C.c_void_p(hp.data_ptr()), C.c_void_p(hs.data_ptr()), n, 255, out)
    ms_e2e = (time.perf_counter() - t0) / 3 * 1e3
    w_pt = {16: 235, 17: 220, 18: 205, 19: 190, 20: 176, 21: 170, 22: 165}.get(k, 176)
    res = {"metric": "G1 MSM 2^%d points" % k, "value": ms, "unit": "ms", "higher_is_better": False, "n": n, "nbits": 255,
           "points_per_s": n / (ms * 1e-3), "gpu_launches": launches,
           "e2e": {"value": ms_e2e, "unit": "ms", "h2d_bytes": n * 128, "d2h_bytes": 96},
           "roofline": {"bound": "int_mul_pipe", "achieved": n * w_pt * IMAD_PER_FPMUL / (ms * 1e-3) / 1e9,
                        "peak": peak_wide / 1e9, "unit": "G IMAD.WIDE/s",
                        "frac": n * w_pt * IMAD_PER_FPMUL / (ms * 1e-3) / peak_wide,
                        "algorithmic_fpmul_per_point": w_pt,
                        "hbm": {"achieved_gbs": n * 128 / (ms * 1e-3) / 1e9, "peak_gbs": hbm_peak,
                                "note": "points + scalars read once = 128 B/point: not the bound"}}}
    if not args.no_cpu_baseline:
        try:
            from oracle import blst_ref as br
            ref_out = (C.c_uint8 * 96)()
            t0 = time.perf_counter()
            br.ref.ref_msm_g1(C.c_void_p(hp.data_ptr()), C.c_void_p(hs.data_ptr()), C.c_size_t(n), C.c_size_t(255), ref_out)
            tc = time.perf_counter() - t0
            res["cpu_baseline"] = {"value": tc * 1e3, "unit": "ms", "cores": 1, "kind": "reference",
                                   "sample": "the same 2^%d inputs, blst_p1s_mult_pippenger + to_affine, one run" % k}
            res["matches_reference"] = bytes(ref_out) == bytes(out)
        except Exception as ex:
            res["cpu_baseline"] = {"value": None, "unit": "ms", "cores": 0, "kind": "reference", "sample": f"unavailable: {ex}"}
    return res


STAGE_KERNEL = {"hash_to_g2": "k_hash_sets", "miller_lines": "k_miller_lines", "miller_acc": "k_miller_acc_team",
                "g1_mul64": "k_g1_mul"}


def ncu_traffic(kernel):
    """DRAM bytes (read + write) of one launch of `kernel` from the committed `ncu --set full` capture of the same
    workload (profiles/r1/r1_raw_pick_<kernel>.txt, written by tools/gpu_profile.sh + tools/ncu_raw_pick.py)."""
    path = os.path.join(ROOT, "profiles", "r1", "r1_raw_pick_%s.txt" % kernel)
    mult = {"byte": 1.0, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9, "Tbyte": 1e12}
    tot, seen = 0.0, 0
    try:
        for ln in open(path):
            f = ln.split()
            if len(f) >= 3 and f[0] in ("dram__bytes_read.sum", "dram__bytes_write.sum"):
                tot += float(f[1].replace(",", "")) * mult.get(f[2], 1.0)
                seen += 1
    except OSError:
        return None, None
    return (tot, os.path.relpath(path, ROOT)) if seen == 2 else (None, None)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=10)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--sets-per-gpu", type=int, default=SETS_PER_GPU)
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-msm", action="store_true", help="skip the G1 MSM 2^20 extra")
    ap.add_argument("--msm-log2", type=int, default=20)
    args = ap.parse_args()
    if args.impl == "reference":
        return run_reference(args)

    import torch
    import torch.distributed as dist
    import nim_blscurve_b200 as bg

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if world > 1:
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    L = bg.lib()
    S = args.sets_per_gpu
    total = S * world
    chunks = CHUNKS_PER_GPU * world
    first = rank * S
    srb = hashlib.sha256(b"Mr F was here").digest()

    cache = bg.BatchedBLSVerifierCache(max_sets=S, device=local)
    h = cache.handle
    # ONE stream for everything in the timed region: torch's H2D copies, the library's kernels and the NCCL
    # all-gather are issued on it, so they are ordered without host synchronisation.  (torch's default stream has
    # handle 0, which blsgpu_set_stream reads as "use the context's own stream" - hence a dedicated stream.)
    stream = torch.cuda.Stream(device=dev)
    torch.cuda.set_stream(stream)
    L.blsgpu_set_stream(h, C.c_void_p(stream.cuda_stream))

    # synthetic workload, generated on the device: rank r owns global sets [r*S, (r+1)*S)
    d_sets = torch.empty(S * 320, dtype=torch.uint8, device=dev)
    rc = L.blsgpu_make_sets(h, 2026, first, S, C.c_void_p(d_sets.data_ptr()), 1)
    assert rc == 0, cache.last_error()
    h_sets = torch.empty(S * 320, dtype=torch.uint8).pin_memory()
    h_sets.copy_(d_sets)
    d_stage = torch.empty(S * 320, dtype=torch.uint8, device=dev)      # e2e staging target
    d_partial = torch.zeros(576, dtype=torch.uint8, device=dev)
    d_flag = torch.zeros(1, dtype=torch.int32, device=dev)
    d_all = torch.zeros(world * 576, dtype=torch.uint8, device=dev)
    d_flags = torch.zeros(world, dtype=torch.int32, device=dev)
    flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)      # > 126 MB L2
    gt = (C.c_uint8 * 576)()
    launches = [0]
    stage_acc = {}

    def step(src_ptr):
        rc = L.blsgpu_partial_dev(h, C.c_void_p(src_ptr), S, first, total, srb, chunks,
                                  C.c_void_p(d_partial.data_ptr()), C.c_void_p(d_flag.data_ptr()))
        assert rc == 0, cache.last_error()
        launches[0] += L.blsgpu_last_launches(h)
        if world > 1:
            dist.all_gather_into_tensor(d_all, d_partial)
            dist.all_gather_into_tensor(d_flags, d_flag)
            rc = L.blsgpu_finalize_dev(h, C.c_void_p(d_all.data_ptr()), world, C.c_void_p(d_flags.data_ptr()), gt)
        else:
            rc = L.blsgpu_finalize_dev(h, C.c_void_p(d_partial.data_ptr()), 1, C.c_void_p(d_flag.data_ptr()), gt)
        launches[0] += 1
        assert rc == 1, f"synthetic batch must verify (rc={rc}) {cache.last_error()}"
        ms = (C.c_float * 16)()
        k = L.blsgpu_last_stage_ms(h, ms, 16)
        for i in range(k):
            nm = L.blsgpu_stage_name(i).decode()
            stage_acc[nm] = stage_acc.get(nm, 0.0) + ms[i]

    def timed(fn, steps):
        """K steps bracketed by barrier + synchronize, device time by CUDA events, max over ranks."""
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(stream)
        for _ in range(steps):
            flush.zero_()
            fn()
        e1.record(stream)
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()
        ms = torch.tensor([e0.elapsed_time(e1)], device=dev)
        if world > 1:
            dist.all_reduce(ms, op=dist.ReduceOp.MAX)
        return float(ms.item())

    def resident():
        step(d_sets.data_ptr())

    def e2e():
        d_stage.copy_(h_sets, non_blocking=True)          # H2D of this step's inputs from pinned memory
        step(d_stage.data_ptr())                          # finalize_dev reads the verdict + GT back (D2H)

    for _ in range(max(args.warmup, 3)):
        resident()
    stage_acc.clear()
    launches[0] = 0
    stop, samples = threading.Event(), []
    th = threading.Thread(target=clocks_sampler, args=(stop, samples, local), daemon=True)
    if rank == 0:
        th.start()
    ms_total = timed(resident, args.steps)
    stop.set()
    n_launch = launches[0]
    stages = {k: v / args.steps for k, v in stage_acc.items()}
    for _ in range(2):
        e2e()
    ms_e2e = timed(e2e, args.steps)
    value = total * args.steps / (ms_total * 1e-3)
    e2e_value = total * args.steps / (ms_e2e * 1e-3)

    extra = {}
    if rank == 0:
        # roofline: integer-multiply pipe.  Peak = IMAD.WIDE multiply-accumulates/s measured by the microbenchmark
        # kernel in this same run; achieved = algorithmic Fp-mul of the dominant kernel x 300 / its event time.
        peak_wide = L.blsgpu_imad_peak(h, 1)
        peak_lo = L.blsgpu_imad_peak(h, 0)
        dom = max(STAGE_FPMUL, key=lambda k: stages.get(k, 0.0))
        dom_ms = stages[dom]
        achieved = S * STAGE_FPMUL[dom] * IMAD_PER_FPMUL / (dom_ms * 1e-3)
        whole = value / world * (W_SET * IMAD_PER_FPMUL)
        peaks = {}
        try:
            peaks = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))
        except Exception:
            pass
        hbm_peak = peaks.get("hbm_gbs", 6650.0)
        traffic, traffic_src = ncu_traffic(STAGE_KERNEL.get(dom, dom))
        extra["roofline"] = {
            "bound": "int_mul_pipe", "kernel": dom, "achieved": achieved / 1e9, "peak": peak_wide / 1e9,
            "unit": "G IMAD.WIDE/s", "frac": achieved / peak_wide, "traffic": traffic,
            "traffic_note": "DRAM bytes read+written per launch of %s, ncu --set full on the same 131072-set workload (%s); "
                            "algorithmic bytes per launch are %d (inputs + outputs) - the rest is per-thread stack "
                            "(local memory) traffic that overflows L2" % (STAGE_KERNEL.get(dom, dom), traffic_src,
                                                                          S * STAGE_IO_BYTES.get(dom, 320)),
            "kernel_ms": dom_ms, "algorithmic_fpmul_per_set": STAGE_FPMUL[dom], "imad_per_fpmul": IMAD_PER_FPMUL,
            "peak_source": "k_imad_peak microbenchmark in this run (mad.wide.u32); mad.lo.u32 peak %.1f G/s" % (peak_lo / 1e9),
            "whole_step_frac": whole / peak_wide,
            "hbm": {"achieved_gbs": S * 320 / (ms_total / args.steps * 1e-3) / 1e9, "peak_gbs": hbm_peak,
                    "note": "320 B of input per set: HBM is not the bound (%s)" %
                            ("of measured" if "hbm_gbs" in peaks else "of fallback")},
        }
        extra["stages_ms"] = stages
        # Eth2 block batch (configs[1]): 129 sets, latency-bound
        blk = 129
        for _ in range(3):
            L.blsgpu_batch_verify_dev(h, C.c_void_p(d_sets.data_ptr()), blk, srb, 4, None, gt)
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        reps = 5
        for _ in range(reps):
            rcb = L.blsgpu_batch_verify_dev(h, C.c_void_p(d_sets.data_ptr()), blk, srb, 4, None, gt)
        tb = (time.perf_counter() - t0) / reps
        extra["block_batch"] = {"sets": blk, "ms": tb * 1e3, "sets_per_s": blk / tb, "verified": rcb == 1,
                                "what": "BASELINE configs[1]: 129-set batch (128 attestations + 1 sync committee), host call "
                                        "blsgpu_batch_verify_dev, wall clock; latency-bound"}
        # the same block with the public-key aggregation done on the GPU first: 128 committees x 128 keys + one of 512
        # (member keys = public keys of the synthetic sets; one segmented aggregateAll launch, host buffers)
        nkeys = min(128 * 128 + 512, (S // 129) * 129)
        member = bytes(h_sets[:nkeys * 320].numpy().reshape(nkeys, 320)[:, :96].tobytes())
        offs = [min(128 * i, nkeys) for i in range(129)] + [nkeys]
        c_offs = (C.c_uint32 * len(offs))(*offs)
        agg_out = (C.c_uint8 * (96 * 129))()
        for _ in range(2):
            L.blsgpu_aggregate_g1_segments(h, member, c_offs, 129, agg_out)
        t0 = time.perf_counter()
        for _ in range(reps):
            rca = L.blsgpu_aggregate_g1_segments(h, member, c_offs, 129, agg_out)
        ta = (time.perf_counter() - t0) / reps
        extra["block_batch"].update({"key_aggregation_ms": ta * 1e3, "keys_aggregated": nkeys, "aggregation_ok": rca == 1,
                                     "ms_with_key_aggregation": (tb + ta) * 1e3})
        if not args.no_msm:
            extra["msm_g1"] = bench_msm(L, h, cache, stream, flush, peak_wide, hbm_peak, args)
        if world == 1 and not args.no_cpu_baseline:
            try:
                from oracle import blst_ref as br
                cores = br.ncores()
                n = max(64, min(S, 2500 * cores))
                sample = bytes(h_sets[:n * 320].numpy().tobytes())
                t = br.time_batch_verify(sample, srb, cores, 1)
                t1n = min(n, 4096)
                t1 = br.time_batch_verify(sample[:t1n * 320], srb, 1, 1)
                # the 129-set block batch on the host cores as well (best of 5: it is a latency comparison)
                tblk = br.time_batch_verify(sample[:blk * 320], srb, cores, 5)
                extra["block_batch"]["cpu_baseline"] = {
                    "value": tblk * 1e3, "unit": "ms", "cores": cores, "kind": "reference",
                    "sample": f"the same {blk} sets, BLST batchVerifyParallel replica, {cores} threads, best of 5"}
                extra["cpu_baseline"] = {
                    "value": n / t, "unit": "sets/s", "cores": cores, "kind": "reference",
                    "sample": f"first {n} sets of the same workload, BLST (oracle/_ref) batchVerifyParallel replica, "
                              f"{cores} threads, one run; single-thread: {t1n / t1:.0f} sets/s on {t1n} sets"}
            except Exception as ex:     # the oracle is optional at bench time
                extra["cpu_baseline"] = {"value": None, "unit": "sets/s", "cores": 0, "kind": "reference",
                                         "sample": f"unavailable: {ex}"}
        line = {
            "metric": "batch-verified signature sets/sec", "value": value, "unit": "sets/s", "n_gpus": world,
            "steps": args.steps, "warmup": max(args.warmup, 3), "ms_per_step": ms_total / args.steps,
            "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
            "dtype": "u32x12 (Fp 381-bit Montgomery)", "data": "synthetic",
            "config": {"workload": f"{S} distinct-message signature sets per GPU = per-GPU share of BASELINE "
                                   f"configs[4] (1M sets on 8 GPUs); one batch of {total} sets per step",
                       "sets_per_step": total, "rlc_chunks": chunks, "l2": "256 MiB flush write between steps",
                       "partial_exchange": "NCCL all_gather of one 576-byte Fp12 per rank" if world > 1 else "none"},
            "clocks": summarize_clocks(samples),
            "e2e": {"value": e2e_value, "unit": "sets/s", "h2d_bytes_per_step": S * 320 * world,
                    "d2h_bytes_per_step": (576 + 16) * world, "ms_per_step": ms_e2e / args.steps},
            "gpu_launches": n_launch,
        }
        line.update(extra)
        print(json.dumps(line), flush=True)
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()
    return 0


if __name__ == "__main__":
    sys.exit(main())
