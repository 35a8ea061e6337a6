#!/usr/bin/env python
"""bench.py — the driver's measurement contract for troy_b200.

Workload (BASELINE.json configs[2], the configuration its metric is quoted on): CKKS, N = 16384, L = 8 data primes +
1 special prime {60, 40x7, 60}, scale 2^40; synthetic random ciphertexts and key-switching keys (uniform residues).
One "step" = Evaluator::multiply followed by Evaluator::relinearize_inplace on a batch of independent ciphertext pairs.

    python bench.py --gpus N --steps K --warmup W            # this repo's CUDA path (one process per GPU under torchrun)
    python bench.py --impl reference --gpus N --steps K ...  # the reference's own CPU path on the host cores

One JSON line on stdout (rank 0). Extra keys `rotate`, `rescale`, `ntt` report the other quantities BASELINE.json's
metric names (rotate ops/s, limb-NTT/s and fraction of the HBM roofline).
"""
import argparse
import ctypes as C
import json
import os
import statistics
import subprocess
import sys
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

N, L, BITS = 16384, 8, [60] + [40] * 7 + [60]
SCALE = 2.0 ** 40
METRIC, UNIT = "ckks_n16384_mult_relin_ops_per_s", "ops/s"
WORKLOAD = "CKKS N=16384 L=8 {60,40x7,60} scale 2^40: multiply+relinearize on a batch of synthetic ciphertext pairs"


def bench_config(world, batch=256):
    """The `config` object of BOTH arms (the reference arm times a bounded sample of this workload; cpu_baseline.sample says which)."""
    return {"workload": WORKLOAD, "batch_per_gpu": batch, "parallelism": f"{world} independent ciphertext shards, no collective",
            "l2": f"inputs {batch * 5 * L * N * 8 >> 20} MiB per step > 126 MiB L2 (no explicit flush)"}


def algo_bytes():
    """SURVEY.md §8(d) per-unit algorithmic HBM bytes."""
    return dict(limb_ntt=16 * N, multiply=56 * L * N, keyswitch=8 * N * (2 * L * L + 7 * L),
                rotate=8 * N * (2 * L * L + 11 * L), rescale=8 * N * (4 * L - 2))


def measured_peaks():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        with open(p) as f:
            return json.load(f).get("hbm_gbs", 6650.0), "measured"
    return 6650.0, "fallback"


def bind_to_gpu_numa_node(local, world):
    """Pin this rank's host threads to the CPUs of the NUMA node its GPU hangs off (and, on a single-node host, to this
    rank's share of them) so that the pinned staging buffers are allocated and driven next to the GPU. Returns what was done."""
    try:
        import torch
        bus = torch.cuda.get_device_properties(local)
        bdf = f"{getattr(bus, 'pci_domain_id', 0):04x}:{bus.pci_bus_id:02x}:{bus.pci_device_id:02x}.0"
        node = -1
        path = f"/sys/bus/pci/devices/{bdf}/numa_node"
        if os.path.exists(path):
            node = int(open(path).read().strip())
        cpus = sorted(os.sched_getaffinity(0))
        if node >= 0 and os.path.exists(f"/sys/devices/system/node/node{node}/cpulist"):
            want = set()
            for part in open(f"/sys/devices/system/node/node{node}/cpulist").read().strip().split(","):
                lo, _, hi = part.partition("-")
                want.update(range(int(lo), int(hi or lo) + 1))
            cpus = [c for c in cpus if c in want] or cpus
        if world > 1 and len(cpus) >= world:  # ranks of one node: disjoint shares of its CPUs
            share = len(cpus) // world
            cpus = cpus[local * share:(local + 1) * share]
        os.sched_setaffinity(0, cpus)
        return {"pci": bdf, "numa_node": node, "cpus": f"{cpus[0]}-{cpus[-1]}" if cpus else ""}
    except Exception as e:  # affinity is an optimisation; never fail the bench for it
        return {"error": str(e)[:120]}


class ClockSampler:
    Q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, device):
        self.device, self.proc = device, None

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits", "-lms", "20",
                                          "-i", str(self.device)], stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
        except OSError:
            self.proc = None

    def stop(self):
        if not self.proc:
            return None
        self.proc.terminate()
        try:
            out, _ = self.proc.communicate(timeout=5)
        except Exception:
            self.proc.kill()
            out = ""
        sm, mx, reasons = [], [], set()
        for line in out.strip().splitlines():
            f = [x.strip() for x in line.split(",")]
            if len(f) < 9:
                continue
            try:
                sm.append(float(f[1]))
                mx.append(float(f[2]))
            except ValueError:
                continue
            for name, v in zip(("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"), f[5:9]):
                if v.lower().startswith("active"):
                    reasons.add(name)
        if not sm:
            return None
        load = [x for x in sm if x >= 0.5 * max(sm)]  # drop idle samples taken before / after the timed region
        return {"sm_mhz": statistics.median(load), "sm_max_mhz": max(mx), "reasons": sorted(reasons), "samples": len(load)}


# ---------------------------------------------------------------------------------------------------------------------
def run_reference(args):
    """The reference arm: the reference's own troy:: CPU implementation (oracle/_ref, compiled from /root/reference) on
    every host core, same config, metric and unit; each step is a bounded sample of the workload."""
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    import numpy as np
    from oracle import ref
    if not ref.available():
        print(json.dumps({"impl": "reference", "unavailable": "oracle/_ref/libtroyref.so was not built"}))
        return
    cores = os.cpu_count() or 1
    q = ref.coeff_modulus_create(N, BITS)
    rc = ref.RefContext(ref.CKKS, N, q)
    rng = np.random.default_rng(0)
    kq = rc.level_moduli(-1)
    key = np.stack([np.stack([np.stack([rng.integers(0, m, N, dtype=np.uint64) for m in kq]) for _ in range(2)]) for _ in range(L)])
    rlk = rc.synthetic_keys(False, 1)
    rlk.set_words(0, key)
    w = np.stack([np.stack([rng.integers(0, m, N, dtype=np.uint64) for m in q[:L]]) for _ in range(2)])
    a, b = rc.ciphertext(w, 0, True, SCALE), rc.ciphertext(w, 0, True, SCALE)
    reps = args.ref_reps
    for _ in range(args.warmup):
        rc.time_ops(0, a, b, rlk, None, 0, 1, cores)
    t = 0.0
    for _ in range(args.steps):
        t += rc.time_ops(0, a, b, rlk, None, 0, reps, cores)
    ops = args.steps * reps * cores
    value = ops / t
    sample = f"{reps} multiply+relinearize per thread x {cores} threads per step (one independent Evaluator per thread)"
    print(json.dumps({
        "impl": "reference", "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": args.gpus, "steps": args.steps,
        "warmup": args.warmup, "ms_per_step": t / args.steps * 1e3, "higher_is_better": True, "scaling": "weak",
        "vs_baseline": None, "dtype": "u64", "data": "synthetic",
        "config": bench_config(args.gpus),
        "cpu_baseline": {"value": value, "unit": UNIT, "cores": cores, "kind": "reference", "sample": sample},
        "e2e": {"value": value, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }))


# ---------------------------------------------------------------------------------------------------------------------
def run_ours(args):
    import numpy as np
    import torch
    import troy_b200 as tb
    from troy_b200.lib import library

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local)
    dist = None
    out_fd = None
    if world > 1:
        # NCCL writes its version banner to stdout when the communicator is created: keep stdout to the ONE JSON line by
        # pointing fd 1 at stderr for the run and printing the line on the saved descriptor.
        sys.stdout.flush()
        out_fd = os.dup(1)
        os.dup2(2, 1)
        import torch.distributed as dist
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    dev = torch.device("cuda", local)
    tb.KernelProvider.initialize(local)
    lib = library()
    B = args.batch
    bytes_ = algo_bytes()
    hbm_peak, peak_kind = measured_peaks()

    q = tb.CoeffModulus.Create(N, BITS)
    parms = tb.EncryptionParameters(tb.SchemeType.ckks)
    parms.set_poly_modulus_degree(N)
    parms.set_coeff_modulus(q)
    ctx = tb.SEALContextCuda(parms)
    stream = torch.cuda.current_stream().cuda_stream
    ev = tb.EvaluatorCuda(ctx, stream=stream)

    # synthetic keys and ciphertexts: uniform residues (values do not change the work)
    g = torch.Generator(device=dev)
    g.manual_seed(0x5EED0000 + rank)

    def uniform(shape_prefix, moduli):
        out = torch.empty(*shape_prefix, len(moduli), N, dtype=torch.int64, device=dev)
        for i, m in enumerate(moduli):
            out[..., i, :] = torch.randint(0, m, (*shape_prefix, N), generator=g, device=dev, dtype=torch.int64)
        return out

    kq = ctx.level_moduli(-1)
    key_t = uniform((L, 2), kq)
    gkey_t = uniform((L, 2), kq)
    rlk = tb.RelinKeysCuda(ctx)
    rlk.keys = [tb.host.ExternalArray(key_t.data_ptr(), key_t.numel())]
    elt = ctx.galois_elt_from_step(1)
    gk = tb.GaloisKeysCuda(ctx)
    gk.keys = [None] * (((elt - 1) >> 1) + 1)
    gk.keys[(elt - 1) >> 1] = tb.host.ExternalArray(gkey_t.data_ptr(), gkey_t.numel())

    a_in = uniform((B, 2), q[:L])           # pristine inputs
    b_in = uniform((B, 2), q[:L])
    a_work = torch.empty(B, 3, L, N, dtype=torch.int64, device=dev)  # capacity 3: multiply writes the third polynomial
    a_ct = tb.CiphertextCuda.from_device(ctx, a_work.data_ptr(), B, 2, 0, True, SCALE, capacity=3)
    b_ct = tb.CiphertextCuda.from_device(ctx, b_in.data_ptr(), B, 2, 0, True, SCALE)

    a_src = tb.CiphertextCuda.from_device(ctx, a_in.data_ptr(), B, 2, 0, True, SCALE)

    def reset_a():
        a_work[:, :2].copy_(a_in)
        a_ct.size, a_ct.scale = 2, SCALE

    def step():
        # multiply(a, b, destination) + relinearizeInplace(destination): the inputs stay pristine, nothing is copied
        ev.multiply(a_src, b_ct, a_ct)
        ev.relinearize_inplace(a_ct, rlk)

    def barrier():
        torch.cuda.synchronize()
        if dist is not None:
            dist.barrier()
        torch.cuda.synchronize()

    def timed(fn, steps, fork=None, join=None):
        """K steps between barriers, CUDA events on the launching stream, max over ranks (seconds). fork / join make
        side streams start after the first event and finish before the second."""
        barrier()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        if fork is not None:
            fork()
        for _ in range(steps):
            fn()
        if join is not None:
            join()
        e1.record()
        barrier()
        t = torch.tensor([e0.elapsed_time(e1) * 1e-3], dtype=torch.float64, device=dev)
        if dist is not None:
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return float(t.item())

    sampler = ClockSampler(local)
    if rank == 0:
        sampler.start()
        time.sleep(0.3)  # let nvidia-smi come up so that it samples during the timed region
    for _ in range(max(args.warmup, 3)):
        step()
    launches0 = lib.cdll.troyb_launch_count()
    t_main = timed(step, args.steps)
    launches = lib.cdll.troyb_launch_count() - launches0
    clocks = sampler.stop() if rank == 0 else None
    value = world * B * args.steps / t_main

    # ---- end to end through the public API with HOST buffers (pinned), copies inside the timed region.
    # The batch is cut into chunks that go round-robin over a few streams (one EvaluatorCuda per stream): chunk c's
    # host->device copy, its multiply+relinearize and its device->host copy overlap with those of its neighbours, so the
    # PCIe link (768 MiB per step) and the SMs work at the same time. Every word of every step crosses the link.
    affinity = bind_to_gpu_numa_node(local, world)  # before the pinned allocations, so that they land on the GPU's node
    a_host = torch.empty(B, 2, L, N, dtype=torch.int64).pin_memory()
    b_host = torch.empty(B, 2, L, N, dtype=torch.int64).pin_memory()
    out_host = torch.empty(B, 2, L, N, dtype=torch.int64).pin_memory()
    a_host.copy_(a_in)
    b_host.copy_(b_in)
    CH = max(1, min(args.e2e_chunk, B))
    n_streams = max(1, args.e2e_streams)
    cur = torch.cuda.current_stream()
    lanes = []
    for _ in range(n_streams):
        st = torch.cuda.Stream(device=dev)
        aw = torch.empty(CH, 3, L, N, dtype=torch.int64, device=dev)
        bw = torch.empty(CH, 2, L, N, dtype=torch.int64, device=dev)
        lanes.append(dict(stream=st, ev=tb.EvaluatorCuda(ctx, stream=st.cuda_stream), aw=aw, bw=bw))

    def step_e2e():
        for ci, c0 in enumerate(range(0, B, CH)):
            nb = min(CH, B - c0)
            ln = lanes[ci % n_streams]
            with torch.cuda.stream(ln["stream"]):
                ln["aw"][:nb, :2].copy_(a_host[c0:c0 + nb], non_blocking=True)
                ln["bw"][:nb].copy_(b_host[c0:c0 + nb], non_blocking=True)
                act = tb.CiphertextCuda.from_device(ctx, ln["aw"].data_ptr(), nb, 2, 0, True, SCALE, capacity=3)
                bct = tb.CiphertextCuda.from_device(ctx, ln["bw"].data_ptr(), nb, 2, 0, True, SCALE)
                ln["ev"].multiply_inplace(act, bct)
                ln["ev"].relinearize_inplace(act, rlk)
                out_host[c0:c0 + nb].copy_(ln["aw"][:nb, :2], non_blocking=True)

    def fork():
        for ln in lanes:
            ln["stream"].wait_stream(cur)

    def join():
        for ln in lanes:
            cur.wait_stream(ln["stream"])

    def step_copy_only():
        # the same chunks on the same streams with the kernels left out: the ceiling the host link sets for this arm
        for ci, c0 in enumerate(range(0, B, CH)):
            nb = min(CH, B - c0)
            ln = lanes[ci % n_streams]
            with torch.cuda.stream(ln["stream"]):
                ln["aw"][:nb, :2].copy_(a_host[c0:c0 + nb], non_blocking=True)
                ln["bw"][:nb].copy_(b_host[c0:c0 + nb], non_blocking=True)
                out_host[c0:c0 + nb].copy_(ln["aw"][:nb, :2], non_blocking=True)

    for _ in range(2):
        fork()
        step_e2e()
        join()
    e2e_steps = max(2, min(args.steps, 20))  # ~0.5 s: the PCIe arm varies by +-10 % over 5 steps on a shared host
    t_e2e = timed(step_e2e, e2e_steps, fork, join)
    if rank == 0 and world == 1:
        # the results that came back over the link are the device path's results
        step()
        torch.cuda.synchronize()
        if not torch.equal(out_host[:4], a_work[:4, :2].cpu()):
            raise SystemExit("bench: end-to-end results differ from the device-resident path")
    fork()
    step_copy_only()
    join()
    t_copy = timed(step_copy_only, e2e_steps, fork, join)
    h2d_b, d2h_b = 2 * a_host.numel() * 8, out_host.numel() * 8
    e2e = {"value": world * B * e2e_steps / t_e2e, "unit": UNIT, "h2d_bytes_per_step": h2d_b,
           "d2h_bytes_per_step": d2h_b, "chunk": CH, "streams": n_streams,
           "copy_only_ops_per_s": world * B * e2e_steps / t_copy,
           "frac_of_copy_only": t_copy / t_e2e,
           "h2d_gbs_per_rank": h2d_b * e2e_steps / t_e2e / 1e9, "d2h_gbs_per_rank": d2h_b * e2e_steps / t_e2e / 1e9,
           "copy_only_h2d_gbs_per_rank": h2d_b * e2e_steps / t_copy / 1e9,
           "limiter": "pinned host<->device copy bandwidth per rank (6 MiB cross the link per operation; no collective)",
           "cpu_affinity": affinity}

    # ---- the other quantities the metric names: rotate, rescale, standalone NTT
    def rot():
        ev.rotate_vector(a_src, 1, gk, a_ct)   # rotateVector(encrypted, steps, keys, destination)

    def resc():
        ev.rescale_to_next(a_src)              # rescaleToNext(encrypted, destination): out of place already

    extra_steps = max(2, min(args.steps, 5))
    for fn in (rot, resc):  # warm-up: first calls grow the stream-ordered scratch pool
        for _ in range(3):
            fn()
    t_rot = timed(rot, extra_steps)
    t_resc = timed(resc, extra_steps)
    ntt_polys = 512  # 512 polys x 8 limbs x 128 KiB = 512 MiB > L2
    ntt_buf = uniform((ntt_polys,), q[:L])

    def ntt_f():
        lib.call("troyb_ntt_forward", C.c_void_p(ctx.h), 0, C.c_void_p(ntt_buf.data_ptr()), ntt_polys, C.c_void_p(stream))

    def ntt_i():
        lib.call("troyb_ntt_inverse", C.c_void_p(ctx.h), 0, C.c_void_p(ntt_buf.data_ptr()), ntt_polys, C.c_void_p(stream))

    for _ in range(3):
        ntt_f()
        ntt_i()
    t_nf = timed(ntt_f, extra_steps)
    t_ni = timed(ntt_i, extra_steps)
    limbs = ntt_polys * L
    # the same transform on this configuration's 40-bit limbs alone (7 of its 8 data limbs; FP64 butterflies) - the mixed figure
    # above is the sum of the two kinds' times (60-bit limbs run integer butterflies: see ntt_sweep)
    q40 = tb.CoeffModulus.Create(N, [40] * L + [60])
    p40 = tb.EncryptionParameters(tb.SchemeType.ckks)
    p40.set_poly_modulus_degree(N)
    p40.set_coeff_modulus(q40)
    c40 = tb.SEALContextCuda(p40, expand_mod_chain=False)

    def ntt_f40():
        lib.call("troyb_ntt_forward", C.c_void_p(c40.h), 0, C.c_void_p(ntt_buf.data_ptr()), ntt_polys, C.c_void_p(stream))

    def ntt_i40():
        lib.call("troyb_ntt_inverse", C.c_void_p(c40.h), 0, C.c_void_p(ntt_buf.data_ptr()), ntt_polys, C.c_void_p(stream))

    ntt_buf.bitwise_and_((1 << 39) - 1)  # valid residues of every 40-bit prime
    for _ in range(3):
        ntt_f40()
        ntt_i40()
    t_nf40 = timed(ntt_f40, extra_steps)
    t_ni40 = timed(ntt_i40, extra_steps)

    # ---- BASELINE config 2: the standalone batched NTT / INTT sweep, N = 4096 .. 65536 x {1, 8, 32} limbs of 60-bit primes,
    # in place on at least 512 MiB of polynomials (> L2) per launch
    ntt_sweep = None
    if world == 1 and not args.no_sweep:
        ntt_sweep = []
        for n_ in (4096, 8192, 16384, 32768, 65536):
            for l_ in (1, 8, 32):
                qs = tb.CoeffModulus.Create(n_, [60] * (l_ + 1))
                ps = tb.EncryptionParameters(tb.SchemeType.ckks)
                ps.set_poly_modulus_degree(n_)
                ps.set_coeff_modulus(qs)
                cs = tb.SEALContextCuda(ps, expand_mod_chain=False)
                polys = max(1, -(-(512 << 20) // (l_ * n_ * 8)))
                sbuf = torch.randint(0, 1 << 59, (polys, l_, n_), generator=g, device=dev, dtype=torch.int64)

                def sw_f():
                    lib.call("troyb_ntt_forward", C.c_void_p(cs.h), 0, C.c_void_p(sbuf.data_ptr()), polys, C.c_void_p(stream))

                def sw_i():
                    lib.call("troyb_ntt_inverse", C.c_void_p(cs.h), 0, C.c_void_p(sbuf.data_ptr()), polys, C.c_void_p(stream))

                for _ in range(2):
                    sw_f()
                    sw_i()
                tf_, ti_ = timed(sw_f, 5), timed(sw_i, 5)
                per = polys * l_ * 5
                ntt_sweep.append({"n": n_, "limbs": l_, "bits": 60, "polys": polys,
                                  "fwd_Mps": round(per / tf_ / 1e6, 3), "inv_Mps": round(per / ti_ / 1e6, 3),
                                  "fwd_frac": round(per / tf_ * 16 * n_ / 1e9 / hbm_peak, 4),
                                  "inv_frac": round(per / ti_ * 16 * n_ / 1e9 / hbm_peak, 4)})
                del sbuf, cs

    # ---- latency of ONE multiply + relinearize (batch 1) through the same API: call to completed result, default-free stream
    one_a = tb.CiphertextCuda.from_device(ctx, a_work.data_ptr(), 1, 2, 0, True, SCALE, capacity=3)
    one_src = tb.CiphertextCuda.from_device(ctx, a_in.data_ptr(), 1, 2, 0, True, SCALE)
    one_b = tb.CiphertextCuda.from_device(ctx, b_in.data_ptr(), 1, 2, 0, True, SCALE)

    def one_op():
        ev.multiply(one_src, one_b, one_a)
        ev.relinearize_inplace(one_a, rlk)

    for _ in range(10):
        one_op()
    torch.cuda.synchronize()
    lat = []
    for _ in range(50):
        t0 = time.perf_counter()
        one_op()
        torch.cuda.synchronize()
        lat.append((time.perf_counter() - t0) * 1e6)
    # the device side alone: 50 operations back to back between two events
    t_dev1 = timed(one_op, 50) / 50 * 1e6
    latency = {"latency_us_batch1": statistics.median(lat), "device_us_batch1": t_dev1,
               "what": "multiply + relinearize of one ciphertext pair, N=16384 L=8: host call to synchronised result (median of 50) "
                       "and device time per operation when issued back to back"}

    # ---- the HE matmul of BASELINE config 5 (MatmulHelper::matmul, app/LinearHelperCKKS.cuh:225-249) at this config's
    # shape: y[b][j] = sum_i x[b][i] * W[i][j], 16 x 8 ciphertexts times 8 x 8 plaintexts, as ONE launch, and the same
    # product as the reference issues it (one multiplyPlain per term + one addInplace per term, fused here into one
    # multiply-accumulate launch per (i, j) over the batch).
    MB_, MI_, MJ_ = 16, 8, 8
    mm_x_buf = uniform((MB_ * MI_, 2), q[:L])
    mm_x = tb.Cipher2d(MB_, MI_, tb.CiphertextCuda.from_device(ctx, mm_x_buf.data_ptr(), MB_ * MI_, 2, 0, True, 2.0 ** 20))
    mm_w_buf = uniform((MI_ * MJ_,), q[:L])
    mm_w = tb.Plain2d(MI_, MJ_, tb.PlaintextCuda.from_device(ctx, mm_w_buf.data_ptr(), L * N, 0, 2.0 ** 20, batch=MI_ * MJ_))
    mm = tb.MatmulHelper(MB_, 1, 1, N // 2)
    mm.grid = (MI_, MJ_)
    mm_keep = []

    def matmul_fused():
        mm_keep[:] = [mm.matmul(ev, mm_x, mm_w)]

    for _ in range(3):
        matmul_fused()
    t_mm = timed(matmul_fused, extra_steps)
    mm_terms = MB_ * MI_ * MJ_
    mm_bytes = (MB_ * MI_ * 2 + MI_ * MJ_ + MB_ * MJ_ * 2) * L * N * 8  # every operand word once

    # ---- encryption on the device: encrypt-zero of B ciphertexts at the first level, public-key and
    # secret-key, one Blake2xb generator per ciphertext (synthetic key words and seeds; parity is the tests' business)
    enc_pk = uniform((2,), kq)
    enc_sk = uniform((), kq)
    enc_seeds = torch.randint(0, 2 ** 62, (B, 8), device=dev, dtype=torch.int64)
    enc_out = torch.empty(B, 2, L, N, dtype=torch.int64, device=dev)

    def enc_asym():
        lib.call("troyb_encrypt_zero_asymmetric", C.c_void_p(ctx.h), 0, C.c_void_p(enc_pk.data_ptr()), C.c_void_p(enc_seeds.data_ptr()),
                 C.c_void_p(enc_out.data_ptr()), 2 * L * N, B, C.c_void_p(stream))

    def enc_sym():
        lib.call("troyb_encrypt_zero_symmetric", C.c_void_p(ctx.h), 0, C.c_void_p(enc_sk.data_ptr()), C.c_void_p(enc_seeds.data_ptr()),
                 C.c_void_p(enc_out.data_ptr()), 2 * L * N, B, C.c_void_p(stream))

    for _ in range(3):
        enc_asym()
        enc_sym()
    t_enc_a = timed(enc_asym, extra_steps)
    t_enc_s = timed(enc_sym, extra_steps)
    encrypt = {"encrypt_zero_per_s": world * B * extra_steps / t_enc_a, "encrypt_zero_symmetric_per_s": world * B * extra_steps / t_enc_s,
               "what": f"Encryptor::encryptZero / encryptZeroSymmetric at the first level on the device, batch {B}, words identical to the "
                       "reference CPU encryptor's for the same seeds (reference CPU, one thread: 69 / 129 per s, profiles/r2_encrypt_bench.jsonl)"}

    # ---- per-kernel device time of one step (CUDA events around every launch) for the roofline of the dominant kernel
    lib.call("troyb_profile_enable", 1)
    step()
    buf = C.create_string_buffer(1 << 16)
    lib.call("troyb_profile_report", buf, len(buf))
    lib.call("troyb_profile_enable", 0)
    kernels = [json.loads(x) for x in buf.value.decode().strip().splitlines() if x]
    total_ms = sum(k["ms"] for k in kernels) or 1.0
    for k in kernels:
        k["share"] = k["ms"] / total_ms
    dom = max(kernels, key=lambda k: k["ms"])
    # algorithmic bytes of one launch of the dominant kernel: every NTT-family CTA (unit) reads and writes one limb
    if dom["kernel"].startswith("ntt_"):
        tiles_per_limb = max(1, N // 8192)  # a limb is a cluster of 8192-coefficient tiles, one CTA each
        units = dom["grid"] / tiles_per_limb / dom["launches"]
        if dom["kernel"] == "ntt_fwd_modup":
            units *= L / (L + 1)  # CKKS: digit i under modulus i is read in place by the inner product, its CTA exits at once
        dom_bytes = units * bytes_["limb_ntt"]
    elif dom["kernel"] == "ks_inner_product":
        dom_bytes = B * (L + 1) * N * 8 * (L + 2 * L + 2) / dom["launches"]
    else:
        dom_bytes = B * bytes_["multiply"] / dom["launches"]
    dom_s = dom["ms"] * 1e-3 / dom["launches"]
    traffic = None
    tp = os.path.join(ROOT, "profiles", "traffic.json")
    if os.path.exists(tp):
        with open(tp) as f:
            traffic = json.load(f).get(dom["kernel"])
    roofline = {"bound": "hbm", "kernel": dom["kernel"], "achieved": dom_bytes / dom_s / 1e9, "peak": hbm_peak, "unit": "GB/s",
                "frac": dom_bytes / dom_s / 1e9 / hbm_peak, "traffic": traffic, "peak_source": peak_kind,
                "share_of_step": dom["share"]}

    # ---- N > 1: ONE large key switch sharded by digit over the GPUs (BASELINE config 5: CKKS N=65536, 40 digits, a 1.7 GB key):
    # reduce-scatter of the row-major key products over NVLink, per-rank mod-down of the owned rows, all-gather. Checked
    # against, and timed next to, the same key switch on ONE GPU (rank 0, whole key).
    digit_sharded = None
    if world > 1 and not args.no_sharded:
        from troy_b200.distributed import DigitShardedKeySwitch, shard_range
        n5, bits5, b5 = 65536, [60] + [50] * 39 + [60], 4
        p5 = tb.EncryptionParameters(tb.SchemeType.ckks)
        p5.set_poly_modulus_degree(n5)
        p5.set_coeff_modulus(tb.CoeffModulus.Create(n5, bits5))
        c5 = tb.SEALContextCuda(p5, expand_mod_chain=False)
        l5 = c5.level_limbs(0)
        kq5 = c5.level_moduli(-1)
        g5 = torch.Generator(device=dev)
        g5.manual_seed(0x5EED5)                     # the same key and ciphertexts on every rank
        key5 = torch.empty(l5, 2, l5 + 1, n5, dtype=torch.int64, device=dev)
        for i_, m_ in enumerate(kq5):
            key5[:, :, i_, :] = torch.randint(0, m_, (l5, 2, n5), generator=g5, device=dev, dtype=torch.int64)
        ct5_in = torch.empty(b5, 3, l5, n5, dtype=torch.int64, device=dev)
        for i_, m_ in enumerate(kq5[:l5]):
            ct5_in[:, :, i_, :] = torch.randint(0, m_, (b5, 3, n5), generator=g5, device=dev, dtype=torch.int64)
        mine5 = shard_range(l5, world, rank)
        ks5 = DigitShardedKeySwitch(c5, None, stream=stream, key_device_ptr=key5[mine5.start:mine5.stop].data_ptr() if len(mine5) else None)
        ct5 = ct5_in.clone()
        ct5_obj = tb.CiphertextCuda.from_device(c5, ct5.data_ptr(), b5, 2, 0, True, 2.0 ** 50, capacity=3)
        tgt5 = ct5_in[:, 2].contiguous()            # the polynomial being switched, [batch][L][N]

        def sharded():
            ks5.switch_key_inplace(ct5_obj, tgt5.data_ptr(), l5 * n5)

        sharded()
        torch.cuda.synchronize()
        first = ct5[:, :2].clone()
        for _ in range(2):
            sharded()
        t_sh = timed(sharded, 5)
        # one GPU, whole key (rank 0 measures; everybody runs it so that the GPUs stay in step)
        ct5s = ct5_in.clone()

        def single():
            lib.call("troyb_switch_key", C.c_void_p(c5.h), 0, C.c_void_p(ct5s.data_ptr()), 3 * l5 * n5, C.c_void_p(tgt5.data_ptr()),
                     l5 * n5, C.c_void_p(key5.data_ptr()), b5, C.c_void_p(stream))

        single()
        torch.cuda.synchronize()
        same = bool(torch.equal(first, ct5s[:, :2]))
        for _ in range(2):
            single()
        t_si = timed(single, 5)
        ok_t = torch.tensor([1 if same else 0], device=dev)
        dist.all_reduce(ok_t, op=dist.ReduceOp.MIN)
        digit_sharded = {"config": f"CKKS N={n5} L={l5} (+1 special prime): one key switch of {l5} digits, digits sharded over {world} GPUs, batch {b5}",
                         "mode": ks5.mode, "ops_per_s": b5 * 5 / t_sh, "single_gpu_ops_per_s": b5 * 5 / t_si, "vs_single_gpu": t_si / t_sh,
                         "parity": bool(ok_t.item() == 1), "nvlink_bytes_sent_per_rank_per_call": ks5.bytes_sent,
                         "collectives": "reduce_scatter (rows) + broadcast (special row) + all_gather (finished limbs), NCCL"}
        del key5, ct5_in, ct5, ct5s, first

    if rank != 0:
        if dist is not None:
            dist.destroy_process_group()
        return
    # ---- CPU baseline: the reference's own CPU path on this box's host cores, bounded sample
    cpu = None
    if world == 1 and not args.no_cpu_baseline:
        try:
            from oracle import ref
            if ref.available():
                cores = os.cpu_count() or 1
                rc = ref.RefContext(ref.CKKS, N, q)
                key_np = key_t.cpu().numpy().astype(np.uint64)
                r_rlk = rc.synthetic_keys(False, 1)
                r_rlk.set_words(0, key_np)
                w = a_in[0].cpu().numpy().astype(np.uint64)
                ra, rb = rc.ciphertext(w, 0, True, SCALE), rc.ciphertext(w, 0, True, SCALE)
                rc.time_ops(0, ra, rb, r_rlk, None, 0, 1, cores)
                reps = 32  # about 1 s of wall time on 16 threads: a bounded sample of ~15 CPU-seconds
                t0 = time.time()
                tt = rc.time_ops(0, ra, rb, r_rlk, None, 0, reps, cores)
                cpu = {"value": reps * cores / tt, "unit": UNIT, "cores": cores, "kind": "reference",
                       "sample": f"{reps} multiply+relinearize per thread x {cores} threads, troy:: CPU Evaluator per thread "
                                 f"({time.time() - t0:.1f} s wall)"}
        except Exception as e:  # the baseline is reporting only; never fail the bench for it
            cpu = {"value": None, "unit": UNIT, "cores": 0, "kind": "reference", "sample": f"failed: {e}"}

    # ---- the reference's OWN CUDA path (troyn::, unmodified sources compiled for sm_100a: oracle/_ref/troyn_ref_bench) on this
    # GPU, same configuration, one reference call per ciphertext (it has no batch API). A reported baseline like the CPU one.
    ref_cuda = None
    tool = os.path.join(ROOT, "oracle", "_ref", "troyn_ref_bench")
    if world == 1 and not args.no_cpu_baseline and os.path.exists(tool):
        try:
            import subprocess
            torch.cuda.synchronize()
            r = subprocess.run([tool, "bench", "64", "3"], capture_output=True, text=True, timeout=600)
            for ln in r.stdout.splitlines():
                if ln.startswith("{"):
                    ref_cuda = json.loads(ln)
        except Exception as e:  # the baseline is optional; the bench line must still print
            ref_cuda = {"unavailable": str(e)[:200]}

    # ---- SEAL 4.0 (the copy the reference vendors, oracle/_ref/seal_bench: plain g++ build, no HEXL) on the host cores, same
    # configuration on real encryptions, one Evaluator loop per thread. Reported next to the troy:: CPU baseline; optional.
    seal = None
    tool = os.path.join(ROOT, "oracle", "_ref", "seal_bench")
    if world == 1 and not args.no_cpu_baseline and os.path.exists(tool):
        try:
            import subprocess
            r = subprocess.run([tool, str(os.cpu_count() or 1), "16"], capture_output=True, text=True, timeout=300)
            for ln in r.stdout.splitlines():
                if ln.startswith("{"):
                    seal = json.loads(ln)
        except Exception as e:
            seal = {"unavailable": str(e)[:200]}

    line = {
        "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps, "warmup": max(args.warmup, 3),
        "ms_per_step": t_main / args.steps * 1e3, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
        "dtype": "u64", "data": "synthetic",
        "config": bench_config(world, B),
        "e2e": e2e, "gpu_launches": int(launches), "clocks": clocks, "roofline": roofline, "cpu_baseline": cpu,
        "mult_relin_frac_of_hbm_roofline": value / world * (bytes_["multiply"] + bytes_["keyswitch"]) / 1e9 / hbm_peak,
        "rotate": {"value": world * B * extra_steps / t_rot, "unit": "ops/s",
                   "frac_of_hbm_roofline": B * extra_steps / t_rot * bytes_["rotate"] / 1e9 / hbm_peak},
        "rescale": {"value": world * B * extra_steps / t_resc, "unit": "ops/s",
                    "frac_of_hbm_roofline": B * extra_steps / t_resc * bytes_["rescale"] / 1e9 / hbm_peak},
        "ntt": {"forward_limb_ntt_per_s": world * limbs * extra_steps / t_nf, "inverse_limb_ntt_per_s": world * limbs * extra_steps / t_ni,
                "forward_gbs": limbs * extra_steps / t_nf * bytes_["limb_ntt"] / 1e9,
                "forward_frac_of_hbm_roofline": limbs * extra_steps / t_nf * bytes_["limb_ntt"] / 1e9 / hbm_peak,
                "inverse_frac_of_hbm_roofline": limbs * extra_steps / t_ni * bytes_["limb_ntt"] / 1e9 / hbm_peak,
                "shape": f"{ntt_polys} polys x {L} limbs, N={N}",
                "limbs_40bit_only": {"forward_limb_ntt_per_s": world * limbs * extra_steps / t_nf40,
                                     "inverse_limb_ntt_per_s": world * limbs * extra_steps / t_ni40,
                                     "forward_frac_of_hbm_roofline": limbs * extra_steps / t_nf40 * bytes_["limb_ntt"] / 1e9 / hbm_peak,
                                     "inverse_frac_of_hbm_roofline": limbs * extra_steps / t_ni40 * bytes_["limb_ntt"] / 1e9 / hbm_peak}},
        "matmul": {"multiply_plain_accumulate_per_s": world * mm_terms * extra_steps / t_mm, "ms": t_mm / extra_steps * 1e3,
                   "shape": f"{MB_}x{MI_} ciphertexts times {MI_}x{MJ_} plaintexts, one launch",
                   "gbs": mm_bytes * extra_steps / t_mm / 1e9, "frac_of_hbm_roofline": mm_bytes * extra_steps / t_mm / 1e9 / hbm_peak},
        "ntt_sweep": ntt_sweep,
        "encrypt": encrypt,
        "digit_sharded": digit_sharded,
        "latency": latency,
        "reference_cuda_path": ref_cuda,
        "seal_cpu_baseline": seal,
        "kernels": kernels,
    }
    if out_fd is None:
        print(json.dumps(line))
    else:
        sys.stdout.flush()
        os.write(out_fd, (json.dumps(line) + "\n").encode())
    if dist is not None:
        dist.destroy_process_group()


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=50)
    ap.add_argument("--warmup", type=int, default=5)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--batch", type=int, default=256, help="ciphertext pairs per GPU per step")
    ap.add_argument("--ref-reps", type=int, default=4, help="reference arm: ops per thread per step")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-sweep", action="store_true", help="skip the config-2 NTT sweep (development runs)")
    ap.add_argument("--no-sharded", action="store_true", help="N > 1: skip the digit-sharded large-N key switch")
    ap.add_argument("--e2e-chunk", type=int, default=16, help="end-to-end arm: ciphertext pairs per chunk")
    ap.add_argument("--e2e-streams", type=int, default=4, help="end-to-end arm: streams the chunks rotate over")
    args = ap.parse_args()
    if args.impl == "reference":
        run_reference(args)
    else:
        run_ours(args)


if __name__ == "__main__":
    main()
