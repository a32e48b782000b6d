#!/usr/bin/env python
"""bench.py -- APR training throughput (triples/s) on BASELINE.json config 4 (synthetic 10M users x 2M items, d=128),
plus full-rank evaluation users/s, with roofline / cpu_baseline / e2e objects (contract in the task brief).

  python bench.py [--gpus N] [--steps K] [--warmup W] [--batch B] [--mode 0|1] [--impl reference]

A "step" is one optimizer step = one batch of B triples through training_batch (utils.py:113-119): Delta update
+ Adagrad on L + reg_adv L_adv.  Tables (12.3 GB with Adagrad slots) are far larger than the 126 MB L2, so successive
steps touch cold rows ("inputs larger than L2").
"""
import argparse
import json
import os
import subprocess
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

METRIC = "apr_train_triples_per_s"
UNIT = "triples/s"
CFG = {"users": 10_000_000, "items": 2_000_000, "d": 128, "eps": 0.5, "reg_adv": 1.0, "lr": 0.05, "reg": 0.0}


def parse():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=2048)
    ap.add_argument("--warmup", type=int, default=128)
    ap.add_argument("--batch", type=int, default=int(os.environ.get("APR_BENCH_BATCH", "65536")))
    ap.add_argument("--mode", type=int, default=int(os.environ.get("APR_BENCH_MODE", "0")))
    ap.add_argument("--impl", default="b200")
    ap.add_argument("--users", type=int, default=CFG["users"])
    ap.add_argument("--items", type=int, default=CFG["items"])
    ap.add_argument("--dim", type=int, default=CFG["d"])
    ap.add_argument("--no-eval", action="store_true")
    ap.add_argument("--no-cpu", action="store_true")
    ap.add_argument("--no-variants", action="store_true")
    ap.add_argument("--eval-users", type=int, default=1 << 20,
                    help="users of the tiled configs[4] evaluation run (multiple of 16384; 0 = skip)")
    ap.add_argument("--cpu-seconds", type=float, default=12.0)
    return ap.parse_args()


def peaks():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        j = json.load(open(p))
        return float(j["hbm_gbs"]), float(j.get("bf16_tflops", 1590.0)), "measured"
    return 6650.0, 1590.0, "fallback"


def sustained_tc_peak(burst):
    """The sustained bf16 figure of MEASURED_PEAKS.json (cuBLAS back to back for seconds): the denominator for a kernel
    timed inside a long run; the burst figure is for a kernel timed alone."""
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        return float(json.load(open(p)).get("bf16_tflops_sustained", burst))
    return burst


class ClockSampler:
    """nvidia-smi clocks + throttle reasons sampled every 100 ms while the timed region runs."""
    Q = ("clocks.sm,clocks.max.sm,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
         "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, index=0):
        self.rows, self.proc, self.index = [], None, index

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", "-i", str(self.index), "--query-gpu=" + self.Q,
                                          "--format=csv,noheader,nounits", "-lms", "100"], stdout=subprocess.PIPE,
                                         stderr=subprocess.DEVNULL, text=True)
            self.thread = threading.Thread(target=self._read, daemon=True)
            self.thread.start()
        except Exception:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.rows.append(line.strip())

    def stop(self):
        if self.proc is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        self.proc.terminate()
        try:
            self.proc.wait(timeout=5)
        except Exception:
            self.proc.kill()
        sm, mx, reasons = [], [], set()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for r in self.rows:
            f = [x.strip() for x in r.split(",")]
            if len(f) < 6:
                continue
            try:
                sm.append(float(f[0]))
                mx.append(float(f[1]))
            except ValueError:
                continue
            for n, v in zip(names, f[2:6]):
                if v.lower().startswith("active"):
                    reasons.add(n)
        return {"sm_mhz": float(np.median(sm)) if sm else None, "sm_max_mhz": max(mx) if mx else None,
                "reasons": sorted(reasons), "samples": len(sm)}


def quiesce_gc():
    """Called BEFORE the warm-up steps.  A cyclic-GC pass of the interpreter costs tens to hundreds of ms with torch and
    numpy loaded; one that lands between the warm-up and a 2-6 ms timed region leaves the GPU (clocks, NVLink) idle in
    front of it, and one inside the region on ONE rank is seen by every rank as a stall at the next cross-rank barrier
    (both measured).  Collect now, freeze the survivors, keep the collector off while the warm-up
    and the timed launches are issued; reference counting still frees everything that is not a cycle."""
    import gc
    if os.environ.get("APR_BENCH_GC", "off") != "off":
        return False
    gc.collect()
    gc.freeze()
    gc.disable()
    return True


def resume_gc(was_off):
    if was_off:
        import gc
        gc.enable()


def synth_triples(rng, n_steps, batch, users, items):
    u = rng.integers(0, users, size=(n_steps, batch), dtype=np.int32)
    i = rng.integers(0, items, size=(n_steps, batch), dtype=np.int32)
    j = rng.integers(0, items, size=(n_steps, batch), dtype=np.int32)
    return u, i, j


# ---------------------------------------------------------------------------------------------------------
# CPU baseline: the oracle's step on touched rows only (same arithmetic as oracle.apr_step)
# ---------------------------------------------------------------------------------------------------------
def cpu_step_sparse(O, P, Q, aP, aQ, u, i, j, lr, reg, reg_adv, eps, adver):
    uu, inv_u = np.unique(u, return_inverse=True)
    ii, inv = np.unique(np.concatenate([i, j]), return_inverse=True)
    Pc, Qc, aPc, aQc = P[uu], Q[ii], aP[uu], aQ[ii]
    O.apr_step(Pc, Qc, aPc, aQc, inv_u, inv[:i.size], inv[i.size:], lr, reg, reg_adv, eps, adver)
    P[uu], Q[ii], aP[uu], aQ[ii] = Pc, Qc, aPc, aQc


def cpu_port():
    """(step function, threads, label): the C/OpenMP restatement when it builds, else the NumPy oracle (1 thread)."""
    try:
        from oracle import c_oracle as C
        # explicit thread count: torchrun exports OMP_NUM_THREADS=1 to every rank, which made the round-1 reference arm
        # single-threaded at N >= 2
        n = C.set_threads(int(os.environ.get("APR_BENCH_CPU_THREADS", os.cpu_count() or 1)))
        return (lambda O, P, Q, aP, aQ, u, i, j, lr, reg, ra, eps, adv: C.step(P, Q, aP, aQ, u, i, j, lr, reg, ra, eps, adv),
                n, "C/OpenMP oracle port (oracle/apr_oracle_c.c), %d threads" % n)
    except Exception as e:  # no compiler / no OpenMP: the NumPy oracle
        return cpu_step_sparse, 1, "NumPy oracle on touched rows (C port unavailable: %r)" % (e,)


def cpu_baseline(args, seconds, n_steps_cap=None):
    """Times the CPU restatement (kind 'port': TensorFlow cannot run here) on the host cores, same shapes."""
    from oracle import apr_oracle as O
    step_fn, _, _ = cpu_port()
    rng = np.random.default_rng(2019)
    U, I, d, B = args.users, args.items, args.dim, args.batch
    # lazily-committed tables: only touched rows ever become resident
    P, Q = np.zeros((U, d), np.float32), np.zeros((I, d), np.float32)
    aP, aQ = np.zeros((U, d), np.float32), np.zeros((I, d), np.float32)
    done, t_used, steps = 0, 0.0, 0
    while t_used < seconds and (n_steps_cap is None or steps < n_steps_cap):
        u, i, j = [x[0] for x in synth_triples(rng, 1, B, U, I)]
        for T, A, idx in ((P, aP, np.unique(u)), (Q, aQ, np.unique(np.concatenate([i, j])))):
            T[idx] = (rng.standard_normal((idx.size, d)) * 0.01).astype(np.float32)
            A[idx] = 0.1
        t0 = time.perf_counter()
        step_fn(O, P, Q, aP, aQ, u, i, j, CFG["lr"], CFG["reg"], CFG["reg_adv"], CFG["eps"], 1)
        t_used += time.perf_counter() - t0
        done += B
        steps += 1
    return done / t_used, steps, t_used


def run_reference(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    K, W = max(1, args.steps), max(0, args.warmup)
    # each step = one batch through the oracle; bounded so that the run ends within minutes
    K_run = min(K, 256)
    if W:
        cpu_baseline(args, 1e9, n_steps_cap=min(W, 2))
    tps, steps, t = cpu_baseline(args, 1e9, n_steps_cap=K_run)
    _, cores, label = cpu_port()
    line = {"impl": "reference", "metric": METRIC, "value": tps, "unit": UNIT, "n_gpus": args.gpus, "steps": K, "warmup": W,
            "ms_per_step": 1e3 * t / steps, "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f32",
            "data": "synthetic", "config": workload_config(args),
            "cpu_baseline": {"value": tps, "unit": UNIT, "cores": cores, "kind": "port",
                             "sample": "%d batches of %d triples, %s (the TensorFlow reference is not installable: "
                                       "no tensorflow/keras wheels in this image)" % (steps, args.batch, label)},
            "e2e": {"value": tps, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
            "host_cpus": os.cpu_count(), "timed_steps": steps}
    print(json.dumps(line))


def workload_config(args):
    return {"workload": "BASELINE.json configs[3]: synthetic %dM users x %dM items, d=%d, APR step (eps 0.5, reg_adv 1, "
                        "lr 0.05, Adagrad), uniform triples" % (args.users // 10 ** 6, args.items // 10 ** 6, args.dim),
            "users": args.users, "items": args.items, "d": args.dim, "batch_per_step": args.batch,
            "step_mode": {1: "persistent-cooperative", 2: "persistent, one thread-block cluster"}.get(args.mode, "fast kernel || pair kernel || general stages (3 streams)"),
            "cache": "inputs larger than L2 (tables %.1f GB vs 126 MB L2)" %
                     (2 * 4 * args.dim * (args.users + args.items) / 1e9)}


def main():
    args = parse()
    if args.impl == "reference":
        return run_reference(args)
    world = int(os.environ.get("WORLD_SIZE", "1"))
    if world > 1:
        return main_sharded(args)
    return main_single(args)


def main_sharded(args):
    """N > 1: the same tables row-sharded over the N GPUs (BASELINE.json configs[3] verbatim), weak scaling: every rank
    contributes args.batch triples to each step, so the global batch is N * batch."""
    import torch
    import torch.distributed as dist

    from apr_b200 import engine
    from apr_b200.distributed import ShardedTables, ShardedTrainer

    world = int(os.environ["WORLD_SIZE"])
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    dist.init_process_group("nccl", device_id=dev)
    U, I, d, Bl, K, W = args.users, args.items, args.dim, args.batch, max(1, args.steps), max(3, args.warmup)
    Bg = Bl * world
    hbm_peak, _, peak_kind = peaks()
    t = ShardedTables(U, I, d, Bg, dev, world=world, rank=rank, symmetric=True)
    engine.init_truncated_normal(t.local("P"), 0.01, 2019, 2 * rank)
    engine.init_truncated_normal(t.local("Q"), 0.01, 2019, 2 * rank + 1)
    engine.fill(t.local("accP"), 0.1)
    engine.fill(t.local("accQ"), 0.1)
    torch.cuda.synchronize()
    dist.barrier()
    rng = np.random.default_rng(2019 + rank)
    CH = max(2, min(32, (1 << 23) // Bg))      # steps per call of the pipelined driver (two workspaces of CH steps each)
    trainer = ShardedTrainer(t, CH, Bl)
    hp = (CFG["lr"], CFG["reg"], CFG["reg_adv"], CFG["eps"], 1)

    def local_chunk(n):
        """this rank's triples of n steps, resident in HBM (the trainer all_gathers them into the global batches)"""
        return [torch.from_numpy(x).to(dev) for x in synth_triples(rng, n, Bl, U, I)]

    def run_chunk(u, i, j, stats=None):
        trainer.train_steps(u, i, j, *hp, stats=stats)

    chunks, left = [], K
    while left > 0:
        n = min(CH, left)
        chunks.append(local_chunk(n))
        left -= n
    warm, done = [], 0
    while done < W:
        n = min(CH, W - done)
        warm.append(local_chunk(n))
        done += n
    for n in sorted({c[0].shape[0] for c in chunks}):      # every call shape of the timed region has run once
        warm.append(local_chunk(n))
        done += n
    warmup_run = done
    sampler = ClockSampler(local)
    if rank == 0 and os.environ.get("APR_BENCH_NO_SAMPLER") != "1":
        sampler.start()
        time.sleep(0.3)
    gc_off = quiesce_gc()
    dist.barrier()
    for c in warm:                                          # the timed region follows the warm-up steps directly
        run_chunk(*c)
    trainer.synchronize()
    trainer.check()
    ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    call_ev = []
    dist.barrier()
    ev0.record()
    for c in chunks:
        run_chunk(*c)
        if os.environ.get("APR_BENCH_CALL_TIMES") == "1":   # diagnostic: where inside the timed region the time goes
            call_ev.append(torch.cuda.Event(enable_timing=True))
            call_ev[-1].record()
    ev1.record()
    resume_gc(gc_off)
    trainer.synchronize()
    dist.barrier()
    trainer.check()
    ms_t = torch.tensor([ev0.elapsed_time(ev1)], device=dev)
    dist.all_reduce(ms_t, op=dist.ReduceOp.MAX)
    ms = float(ms_t.item())
    clocks = sampler.stop() if rank == 0 else None
    value = K * Bg / (ms * 1e-3)
    # algorithmic bytes of the global steps (same index arrays on every rank)
    last_ws = trainer.ws[(trainer.calls - 1) & 1]
    cnt = last_ws.unique_counts(chunks[-1][0].shape[0]).astype(np.int64)
    bytes_step = float(16 * d * cnt.sum() + 12 * Bg * cnt.shape[0]) / cnt.shape[0]
    achieved = bytes_step * K / (ms * 1e-3) / 1e9
    # NVLink bytes per step and GPU of this exchange pattern: every item row a rank touches whose owner is another rank
    # crosses the link twice (w + acc in, w + acc out), 4 d bytes per row and direction pair
    remote = (world - 1) / world
    nvl_dir = float(cnt[:, 1].mean()) / world * remote * 2 * 4 * d      # bytes per direction per GPU per step (lower bound)
    roofline = {"bound": "hbm", "achieved": achieved, "peak": hbm_peak * world, "unit": "GB/s",
                "frac": achieved / (hbm_peak * world), "traffic": None, "peak_kind": peak_kind,
                "kernel": "fast_kernel + pair_kernel + general_stage_kernel over NVLink-peer-mapped row shards",
                "nvlink_gb_s_per_direction_per_gpu": nvl_dir / (ms / K * 1e-3) / 1e9,
                "note": "aggregate over %d GPUs, whole step (index preparation, exchange and barriers included); (G-1)/G of "
                        "the item-row traffic crosses NVLink: nvlink_* = unique item rows / G x (G-1)/G x (w + acc) per "
                        "direction, against 900 GB/s nominal" % world}
    # e2e: pinned host batches on every rank -> H2D + all_gather on the trainer's side stream -> sharded steps; D2H of the
    # per-step loss
    Ke = min(max(K, 8 * CH), 16 * CH)
    host = [torch.from_numpy(x).pin_memory() for x in synth_triples(rng, Ke, Bl, U, I)]
    stats = torch.zeros((Ke, 2), dtype=torch.float32, device=dev)

    def e2e_pass():
        stats.zero_()
        for s0 in range(0, Ke, CH):
            n = min(CH, Ke - s0)
            run_chunk(*[x[s0:s0 + n] for x in host], stats=stats[s0:s0 + n])
        dist.all_reduce(stats)
        return stats.cpu()

    e2e_pass()
    trainer.synchronize()
    dist.barrier()
    t0 = time.perf_counter()
    e2e_pass()
    trainer.synchronize()
    te = torch.tensor([time.perf_counter() - t0], device=dev, dtype=torch.float64)
    dist.all_reduce(te, op=dist.ReduceOp.MAX)
    trainer.check()
    e2e = {"value": Ke * Bg / float(te.item()), "unit": UNIT, "h2d_bytes_per_step": 12 * Bl, "d2h_bytes_per_step": 8,
           "steps": Ke, "api": "distributed.ShardedTrainer.train_steps(pinned host batches per rank)"}
    cfgd = workload_config(args)
    cfgd.update({"batch_per_step": Bg, "batch_per_gpu": Bl, "parallelism": "row-sharded tables over %d GPUs (NVLink peer "
                 "loads/stores/REDs inside the kernels), data-parallel over triples" % world,
                 "step_mode": "fast kernel || pair kernel || general stages, 3 cross-rank barriers per step; index preparation "
                              "+ one packed block per sub-chunk pushed into the peers' symmetric memory, pipelined on a side "
                              "stream (ShardedTrainer)",
                 "steps_per_call": CH})
    line = {"metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": K, "warmup": W, "ms_per_step": ms / K,
            "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
            "config": cfgd, "roofline": roofline, "e2e": e2e, "gpu_launches": K * 8 + 5 * len(chunks), "clocks": clocks,   # per step: fast, pair, 3 stages, 3 barriers
            "warmup_steps_run": warmup_run}
    if call_ev:
        line["call_ms"] = [round(a.elapsed_time(b), 3) for a, b in zip([ev0] + call_ev[:-1], call_ev)]
    # ---- second headline metric at N GPUs: item-sharded full-rank evaluation (BASELINE.json configs[4]) ----------
    if not args.no_eval and args.eval_users > 0:
        del t, trainer, last_ws
        torch.cuda.empty_cache()
        _, tc_peak, _ = peaks()
        try:
            ev = {"config5_%d_users_x_10M_items_d256" % args.eval_users:
                  bench_eval_config5(torch, engine, dev, tc_peak, args.eval_users, world=world, rank=rank)}
            ku = min(args.eval_users, 1 << 18)
            ev["config5_top100_%d_users_x_10M_items_d256" % ku] = bench_eval_config5(torch, engine, dev, tc_peak, ku, k_top=100,
                                                                                     world=world, rank=rank)
            line["eval"] = ev
        except Exception as e:  # never lose the training line
            line["eval"] = {"error": repr(e)}
    if rank == 0:
        print(json.dumps(line))
    dist.destroy_process_group()


def main_single(args):
    import torch
    import torch.distributed as dist

    from apr_b200 import engine
    from apr_b200.APR import MF, Session
    from apr_b200.utils import training_batch

    world = 1
    rank = 0
    local = int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local)
    dev = engine.require_cuda()

    U, I, d, B, K, W = args.users, args.items, args.dim, args.batch, max(1, args.steps), max(3, args.warmup)
    hbm_peak, tc_peak, peak_kind = peaks()

    # ---- model through the reference-facing surface -----------------------------------------------------
    import types
    margs = types.SimpleNamespace(embed_size=d, lr=CFG["lr"], reg=CFG["reg"], dns=1, adv="grad", eps=CFG["eps"], adver=1,
                                  reg_adv=CFG["reg_adv"], epochs=0, seed=2019 + rank)
    model = MF(U, I, margs)
    model.extra_row = 0
    model.build_graph()
    sess = Session(mode=args.mode)
    rng = np.random.default_rng(2019 + rank)

    CH = max(1, min(256, (1 << 22) // B))  # steps per chunk (one prepare + one run each)

    def device_chunk(n):
        u, i, j = synth_triples(rng, n, B, U, I)
        return [torch.from_numpy(x).to(dev) for x in (u, i, j)]

    ws = sess.workspace(CH, B, d)
    P, Q, aP, aQ = model.embedding_P, model.embedding_Q, model.acc_P, model.acc_Q
    hp = (CFG["lr"], CFG["reg"], CFG["reg_adv"], CFG["eps"], 1)

    def run_chunk(u, i, j):
        engine.train_steps(P, Q, aP, aQ, u, i, j, *hp, ws, mode=args.mode)

    # ---- timed-region inputs: K steps, ids resident in HBM -----------------------------------------------
    chunks = []
    left = K
    while left > 0:
        n = min(CH, left)
        chunks.append(device_chunk(n))
        left -= n

    # ---- warm-up: W steps, then one call of every chunk SHAPE the timed region uses (so that whatever a call
    # shape builds lazily -- executable CUDA graphs, stream/event pools, kernel attributes -- exists before the clock
    # starts).  The warm-up inputs are made first and the clock sampler is started before the warm-up, so that the timed
    # region follows the warm-up steps directly (no idle GPU in between: a 20-step region is 2 ms long).
    warm, done = [], 0
    while done < W:
        n = min(CH, W - done)
        warm.append(device_chunk(n))
        done += n
    for n in sorted({c[0].shape[0] for c in chunks}):
        warm.append(device_chunk(n))
        done += n
    warmup_run = done
    sampler = ClockSampler(local)
    if rank == 0:
        sampler.start()
        time.sleep(0.3)
    gc_off = quiesce_gc()
    for c in warm:
        run_chunk(*c)
    torch.cuda.synchronize()
    ctx0 = engine.context_stats()
    ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    ev0.record()
    for c in chunks:
        run_chunk(*c)
    ev1.record()
    resume_gc(gc_off)
    torch.cuda.synchronize()
    ms = ev0.elapsed_time(ev1)
    clocks = sampler.stop() if rank == 0 else None
    ctx1 = engine.context_stats()
    value = world * K * B / (ms * 1e-3)
    pairs_on = d <= 128 and os.environ.get("APR_PAIRS", "1") != "0"   # pair work units (csrc/train.cu: pairs_enabled)

    def n_launches(S):  # index-preparation kernels per L2-sized sub-chunk + the step kernels (csrc/train.cu)
        p2 = 1 << max(0, (B - 1).bit_length())
        sc = max(1, min(S, (48 << 20) // ((max(32, 2 * p2) + max(32, 4 * p2)) * 8)))
        prep = 5 if pairs_on else 4      # insert, compact, scatter, [pair detection], pack
        step = 5 if pairs_on else 4      # fast kernel, [pair kernel], three general stages
        nsub, nadv, left, n = 0, 0, S, min(2, sc)    # sub-chunks of 2, 4, 8, ... sc steps (apr_train_steps)
        while left > 0:
            ns = min(left, n)
            nsub, left, n = nsub + 1, left - ns, min(sc, 2 * n)
            nadv += ns // 32 + bin(ns % 32).count("1")     # graph launches of 32/16/8/4/2/1 steps: one cursor-advance node each
        graphs = B >= int(os.environ.get("APR_GRAPH_MIN_BATCH", "1024")) and os.environ.get("APR_GRAPH", "1") != "0"
        return prep * nsub + (nsub if args.mode in (1, 2) else step * S + ((nadv + 1) if graphs else 0))
    launches = sum(n_launches(c[0].shape[0]) for c in chunks)

    # ---- roofline of the dominant kernel(s): the embedding step kernels, index preparation excluded -------
    # algorithmic bytes per step = 16 d (U_uniq + I_uniq) + 12 B   
    n_roof = min(len(chunks), 4)
    bytes_total, ms_run = 0.0, 0.0
    for c in chunks[:n_roof]:
        engine.train_prepare(P, Q, *c, ws)
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        # the prepared chunk runs twice back to back and the SECOND run is timed: its launches are enqueued while the
        # first run executes, so the events bracket device time only (not the host's enqueue / graph-update time)
        engine.train_run(P, Q, aP, aQ, *c, *hp, ws, mode=args.mode)
        e0.record()
        engine.train_run(P, Q, aP, aQ, *c, *hp, ws, mode=args.mode)
        e1.record()
        torch.cuda.synchronize()
        ms_run += e0.elapsed_time(e1)
        cnt = ws.unique_counts(c[0].shape[0]).astype(np.int64)
        bytes_total += float(16 * d * cnt.sum() + 12 * B * c[0].shape[0])
    achieved = bytes_total / (ms_run * 1e-3) / 1e9
    steps_roof = sum(c[0].shape[0] for c in chunks[:n_roof])
    bytes_step = bytes_total / steps_roof
    whole = bytes_step / (ms / K * 1e-3) / 1e9        # the same bytes over the TIMED region's ms_per_step (index preparation included)
    roofline = {"bound": "hbm", "achieved": achieved, "peak": hbm_peak, "unit": "GB/s", "frac": achieved / hbm_peak,
                "frac_kind": "kernel-only: the step kernels of an already prepared chunk (apr_train_run), CUDA events",
                "whole_step_achieved": whole, "whole_step_frac": whole / hbm_peak,
                "whole_step_kind": "bytes_per_step_model / ms_per_step of the timed region (index preparation + step kernels)",
                "traffic": measured_traffic(B, d, args.mode), "peak_kind": peak_kind,
                "kernel": "step_persistent_kernel" if args.mode in (1, 2) else
                ("fast_kernel || pair_kernel || 3 x general_stage_kernel per step" if pairs_on else
                 "fast_kernel || 3 x general_stage_kernel per step"),
                "bytes_per_step_model": bytes_step, "ms_per_step_kernel": ms_run / steps_roof}

    # ---- e2e: public API with HOST batches; H2D of ids and D2H of the per-step loss inside the region ------
    # (at least 256 steps: a host-timed region of 20 steps = 2 ms measures the host's launch jitter, not the path)
    Ke = min(max(K, 256), 4 * CH)
    # the steps' inputs wait in pinned host memory (bench contract); Session.train_steps streams them chunk by chunk
    hu, hi, hj = [torch.from_numpy(x).pin_memory() for x in synth_triples(rng, Ke, B, U, I)]
    stats = torch.zeros((Ke, 2), dtype=torch.float32, device=dev)

    def e2e_pass():
        sess.train_steps(model, hu, hi, hj, adver=True, stats=stats)
        return stats.cpu()

    e2e_pass()
    torch.cuda.synchronize()
    if world > 1:
        dist.barrier()
    t0 = time.perf_counter()
    e2e_pass()
    torch.cuda.synchronize()
    t_e2e = time.perf_counter() - t0
    if world > 1:
        t = torch.tensor([t_e2e], device=dev, dtype=torch.float64)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        t_e2e = float(t.item())
    e2e = {"value": world * Ke * B / t_e2e, "unit": UNIT, "h2d_bytes_per_step": 12 * B, "d2h_bytes_per_step": 8,
           "steps": Ke, "api": "Session.train_steps(model, pinned host batches): chunked H2D on a copy stream overlapped with the steps; == utils.training_batch per step"}

    line = {"metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": K, "warmup": W,
            "ms_per_step": ms / K, "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f32",
            "data": "synthetic", "config": workload_config(args), "roofline": roofline, "e2e": e2e,
            "gpu_launches": launches, "clocks": clocks, "warmup_steps_run": warmup_run,
            "graph_cache": {"instantiations_in_timed_region": ctx1["graph_instantiations"] - ctx0["graph_instantiations"],
                            "updates_in_timed_region": ctx1["graph_updates"] - ctx0["graph_updates"],
                            "launches_in_timed_region": ctx1["graph_launches"] - ctx0["graph_launches"],
                            "cached": ctx1["graphs_cached"]}}

    # ---- SURVEY 8(d) variants on the same tables: batch sweep 2^9 .. 2^20 and the Zipf(1.05)-items contention case ---
    if not args.no_variants and rank == 0:
        try:
            line["variants"] = bench_variants(torch, engine, sess, model, args, hp, rng, dev)
        except Exception as e:
            line["variants"] = {"error": repr(e)}

    # ---- second headline metric: full-rank evaluation users/s (BASELINE.json configs[2] shape) ------------
    if not args.no_eval and rank == 0:
        try:
            line["eval"] = bench_eval(torch, engine, dev, tc_peak, args.eval_users)
        except Exception as e:  # never lose the training line
            line["eval"] = {"error": repr(e)}

    if rank == 0 and not args.no_cpu:
        tps, steps, t = cpu_baseline(args, args.cpu_seconds)
        _, cores, label = cpu_port()
        line["cpu_baseline"] = {"value": tps, "unit": UNIT, "cores": cores, "kind": "port",
                                "sample": "%d batches of %d triples in %.1f s, %s (TensorFlow reference not installable "
                                          "here)" % (steps, B, t, label),
                                "host_cpus": os.cpu_count(),
                                "reference_logs": "APR phase 15-87 k triples/s on unknown CPU (BASELINE.md 1.2)"}
    if rank == 0:
        print(json.dumps(line))
    if world > 1:
        dist.destroy_process_group()


def measured_traffic(B, d, mode):
    """DRAM bytes per step (dram__bytes_read.sum + dram__bytes_write.sum over the step's kernels) from the committed ncu
    capture of this workload, or None: profiles/traffic.json = [{"batch", "d", "mode", "bytes_per_step", "source"}, ...]
    written from the `ncu --set full` CSV named in "source" (never a number typed in by hand)."""
    p = os.path.join(ROOT, "profiles", "traffic.json")
    if not os.path.exists(p):
        return None
    try:
        for r in json.load(open(p)):
            if r.get("batch") == B and r.get("d") == d and r.get("mode") == mode:
                return float(r["bytes_per_step"])
    except Exception:
        return None
    return None


def zipf_items(rng, shape, items, s=1.05):
    """Bounded Zipf(s) over [0, items): p(k) ~ (k+1)^-s, by inverse CDF."""
    cdf = np.cumsum(np.arange(1, items + 1, dtype=np.float64) ** (-s))
    cdf /= cdf[-1]
    return np.minimum(np.searchsorted(cdf, rng.random(shape)), items - 1).astype(np.int32)


def bench_variants(torch, engine, sess, model, args, hp, rng, dev):
    """Short device-resident runs of the same step on the same tables (ids in HBM, CUDA events, >= 3 warm-up steps).
    Not bench lines: they document the batch sweet spot and the duplicate-heavy regime named in SURVEY 8(d)."""
    U, I, d = args.users, args.items, args.dim
    P, Q, aP, aQ = model.embedding_P, model.embedding_Q, model.acc_P, model.acc_Q

    def run(B, ids_fn, target_triples):
        nonlocal hp
        CH = max(1, min(256, (1 << 22) // B))
        n_chunks = max(1, int(round(target_triples / (CH * B))))
        ws = sess.workspace(CH, B, d)
        chunks = [[torch.from_numpy(x).to(dev) for x in ids_fn(CH, B)] for _ in range(n_chunks + 1)]
        engine.train_steps(P, Q, aP, aQ, *chunks[0], *hp, ws, mode=args.mode)   # warm-up: CH >= 4 steps
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for c in chunks[1:]:
            engine.train_steps(P, Q, aP, aQ, *c, *hp, ws, mode=args.mode)
        e1.record()
        torch.cuda.synchronize()
        ms = e0.elapsed_time(e1)
        steps = n_chunks * CH
        cnt = ws.unique_counts(CH).astype(np.int64)          # unique rows of the last chunk's steps
        out = {"batch": B, "steps": steps, "ms_per_step": ms / steps, "triples_per_s": steps * B / (ms * 1e-3),
               "unique_rows_per_triple": float(cnt.sum()) / (CH * B)}
        del ws, chunks
        torch.cuda.empty_cache()
        return out

    uniform = lambda n, B: synth_triples(rng, n, B, U, I)
    sweep = []
    for lb in (9, 12, 14, 15, 16, 17, 18, 20):
        B = 1 << lb
        sweep.append(run(B, uniform, 32e6 if lb >= 14 else 4096 * B))
    best = max(sweep, key=lambda r: r["triples_per_s"])

    def zipf(n, B):
        u = rng.integers(0, U, size=(n, B), dtype=np.int32)
        i = zipf_items(rng, (n, B), I)
        j = rng.integers(0, I, size=(n, B), dtype=np.int32)
        return u, i, j

    z = run(args.batch, zipf, 16e6)
    z["items"] = "positives ~ Zipf(1.05) over %d items, negatives uniform" % I
    # the reference's first phase (MF-BPR pretraining, APR.py:232-236): same step without the adversarial pass
    hp_apr = hp
    hp = hp_apr[:4] + (0,)
    bpr = run(args.batch, uniform, 32e6)
    bpr["note"] = "adver = 0: one forward/backward, Adagrad; same algorithmic bytes per triple as the APR step"
    hp = hp_apr
    out = {"batch_sweep_uniform": sweep, "best_batch": best["batch"], "zipf_1.05_items": z, "bpr_step": bpr}
    out["sampler_epoch"] = bench_sampler(torch, engine, U, I, args.batch, dev)
    # the other embedding sizes of BASELINE.json's configs (d = 64: configs[0-1]; d = 256: configs[4]) on the same
    # 10M x 2M tables, each with its own roofline object (kernel-only and whole-step, like the headline line)
    for dv in (64, 256):
        if dv != d:
            try:
                out["dim_%d" % dv] = bench_dim_variant(torch, engine, args, hp_apr, rng, dev, dv)
            except Exception as e:
                out["dim_%d" % dv] = {"error": repr(e)}
    return out


def bench_dim_variant(torch, engine, args, hp, rng, dev, d, steps=512):
    """The headline measurement at another embedding size: device-resident ids, `steps` timed steps after a same-shape
    warm-up (whole step: index preparation + step kernels), then the step kernels alone on prepared chunks."""
    U, I, B = args.users, args.items, args.batch
    hbm_peak, _, peak_kind = peaks()
    P = torch.empty((U, d), dtype=torch.float32, device=dev)
    Q = torch.empty((I, d), dtype=torch.float32, device=dev)
    engine.init_truncated_normal(P, 0.01, 2019, 0)
    engine.init_truncated_normal(Q, 0.01, 2019, 1)
    aP, aQ = torch.full_like(P, 0.1), torch.full_like(Q, 0.1)
    CH = max(1, min(256, (1 << 22) // B))
    ws = engine.TrainWorkspace(CH, B, d, dev)
    chunks = [[torch.from_numpy(x).to(dev) for x in synth_triples(rng, CH, B, U, I)] for _ in range(max(1, steps // CH) + 1)]
    engine.train_steps(P, Q, aP, aQ, *chunks[0], *hp, ws, mode=args.mode)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for c in chunks[1:]:
        engine.train_steps(P, Q, aP, aQ, *c, *hp, ws, mode=args.mode)
    e1.record()
    torch.cuda.synchronize()
    n = (len(chunks) - 1) * CH
    ms_step = e0.elapsed_time(e1) / n
    c = chunks[-1]
    engine.train_prepare(P, Q, *c, ws)
    engine.train_run(P, Q, aP, aQ, *c, *hp, ws, mode=args.mode)
    k0, k1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    k0.record()
    engine.train_run(P, Q, aP, aQ, *c, *hp, ws, mode=args.mode)
    k1.record()
    torch.cuda.synchronize()
    ms_kernel = k0.elapsed_time(k1) / CH
    cnt = ws.unique_counts(CH).astype(np.int64)
    bytes_step = float(16 * d * cnt.sum() + 12 * B * CH) / CH
    res = {"d": d, "batch": B, "steps": n, "ms_per_step": ms_step, "triples_per_s": B / (ms_step * 1e-3),
           "roofline": {"bound": "hbm", "achieved": bytes_step / (ms_kernel * 1e-3) / 1e9, "peak": hbm_peak, "unit": "GB/s",
                        "frac": bytes_step / (ms_kernel * 1e-3) / 1e9 / hbm_peak, "frac_kind": "kernel-only (apr_train_run)",
                        "whole_step_achieved": bytes_step / (ms_step * 1e-3) / 1e9,
                        "whole_step_frac": bytes_step / (ms_step * 1e-3) / 1e9 / hbm_peak, "traffic": measured_traffic(B, d, args.mode),
                        "peak_kind": peak_kind, "bytes_per_step_model": bytes_step, "ms_per_step_kernel": ms_kernel}}
    del P, Q, aP, aQ, ws, chunks
    torch.cuda.empty_cache()
    return res


def bench_sampler(torch, engine, U, I, B, dev, pairs=1 << 26):
    """The epoch sampler (A8: APR.py:30-81) at the configs[3] scale: 2^26 training pairs of U users x I items, one
    uniform negative per pair rejected against the user's sorted CSR row.  Counter-based (Feistel permutation + Philox),
    so it is a pure streaming kernel: bytes = read one (u, i) pair (8) + write u, i, u_dns, j (16) + the CSR probes."""
    g = torch.Generator(device=dev)
    g.manual_seed(2019)
    pu = torch.randint(0, U, (pairs,), device=dev, dtype=torch.int32, generator=g)
    pu, _ = torch.sort(pu)
    pi = torch.randint(0, I, (pairs,), device=dev, dtype=torch.int32, generator=g)
    # CSR of the train lists: rows sorted by item id (duplicates are harmless for the membership test)
    key = pu.to(torch.int64) * I + pi.to(torch.int64)
    key, _ = torch.sort(key)
    cidx = (key % I).to(torch.int32)
    counts = torch.bincount(pu, minlength=U)
    cptr = torch.zeros(U + 1, dtype=torch.int64, device=dev)
    cptr[1:] = torch.cumsum(counts, 0)
    del key, counts
    for ep in range(2):
        out = engine.sample_epoch(pu, pi, B, I, cptr, cidx, 2019, ep, 1)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    reps = 5
    e0.record()
    for ep in range(reps):
        out = engine.sample_epoch(pu, pi, B, I, cptr, cidx, 2019, 2 + ep, 1)
    e1.record()
    torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / reps
    assert int(out[4].item()) == 0
    n = (pairs // B) * B
    return {"pairs": pairs, "batch": B, "ms_per_epoch": ms, "triples_per_s": n / (ms * 1e-3),
            "stream_gb_per_s": 24.0 * n / (ms * 1e-3) / 1e9,
            "note": "includes the output allocation of the call; 24 B/triple streamed + random CSR probes (HBM-latency bound)"}


def bench_eval_config5(torch, engine, dev, tc_peak, users_total, tile=16384, I=10_000_000, d=256, k_top=0, world=1, rank=0):
    """BASELINE.json configs[4] as SURVEY 8(d)/(e) ask for it: users_total users (in tiles of one library call each)
    against all I items, ITEM-SHARDED over `world` ranks: rank r holds only Q[lo_r:hi_r) (I/world rows), scores every user
    tile against it on the tensor cores, the held-out scores / positions are all_reduced and, with k_top, the per-shard
    top-k lists are all_gathered and merged by the apr_topk_merge kernel (apr_b200.distributed).  The item-operand
    image of the shard is built once and reused by every user tile (cache_q).  Timed on the device, max over ranks."""
    import torch.distributed as dist

    from apr_b200.distributed import evaluate_item_sharded_cuda, shard_bounds
    lo, hi = shard_bounds(I, world, rank, 128)
    g = torch.Generator(device=dev)
    g.manual_seed(2019)                                                    # same users / held-out items on every rank
    P = torch.randn((users_total, d), device=dev, generator=g) / d ** 0.5
    test = torch.randint(0, I, (users_total,), device=dev, dtype=torch.int32, generator=g)
    g.manual_seed(3000 + rank)                                             # this rank's rows of the item table
    Qs = torch.randn((hi - lo, d), device=dev, generator=g) / d ** 0.5
    users = torch.arange(tile, device=dev, dtype=torch.int32)
    ptr = torch.arange(0, tile + 1, device=dev, dtype=torch.int64)         # exclusion = the held-out item
    pos = torch.zeros(users_total, dtype=torch.int32, device=dev)
    n_tiles = users_total // tile
    last = {}

    def call(k):
        sl = slice(k * tile, (k + 1) * tile)
        r = evaluate_item_sharded_cuda(P[sl], Qs, users, test[sl], I, ptr, test[sl], k_top=k_top, q_row_offset=lo, cache_q=True)
        pos[sl] = r[0]
        last["ids"] = r[1]

    call(0)                                                                 # warm-up: builds the cached item image
    call(min(1, n_tiles - 1))
    torch.cuda.synchronize()
    if world > 1:
        dist.barrier()
    if world == 1:
        engine.eval_tc_timing(True)
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    kms, n_sampled = 0.0, 0
    e0.record()
    for k in range(n_tiles):
        call(k)
        if world == 1 and k % 8 == 7:                                       # GEMM-kernel time of every 8th tile (counting pass)
            kms += engine.eval_tc_timing(True)
            n_sampled += 1
    e1.record()
    torch.cuda.synchronize()
    if world == 1:
        engine.eval_tc_timing(False)
    ms = e0.elapsed_time(e1)
    if world > 1:
        t = torch.tensor([ms], device=dev)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        ms = float(t.item())
    passes = 2 if k_top else 1
    flops_run = 2.0 * users_total * I * 3 * d * passes
    sus_peak = sustained_tc_peak(tc_peak)
    res = {"users": users_total, "items": I, "d": d, "user_tile": tile, "k_top": k_top, "n_gpus": world,
           "items_per_gpu": hi - lo, "ms": ms, "users_per_s": users_total / (ms * 1e-3),
           "hr10": float((pos < 10).float().mean().item()),
           "issued_tflops_whole_run_per_gpu": flops_run / world / (ms * 1e-3) / 1e12,
           "whole_run_frac_of_sustained_peak": flops_run / world / (ms * 1e-3) / 1e12 / sus_peak,
           "q_image": "built once per table version, reused by every user tile",
           "note": "issued bf16 flops = 3 x useful (hi*hi + hi*lo + lo*hi split)%s; whole run = operand split of the user "
                   "tiles, GEMM pass(es), exact re-scoring, exclusion correction%s" %
                   (", two GEMM passes with k_top" if k_top else "", ", collectives" if world > 1 else "")}
    if world == 1 and n_sampled and not k_top:
        k_ms = kms / n_sampled
        flops_tile = 2.0 * tile * I * 3 * d
        res["gemm_kernel_ms_per_tile"] = k_ms
        res["roofline"] = {"bound": "tensor", "achieved": flops_tile / (k_ms * 1e-3) / 1e12, "peak": sus_peak,
                           "peak_kind": "sustained (long run); burst peak %.1f" % tc_peak, "unit": "TFLOP/s",
                           "frac": flops_tile / (k_ms * 1e-3) / 1e12 / sus_peak,
                           "frac_of_burst_peak": flops_tile / (k_ms * 1e-3) / 1e12 / tc_peak,
                           "whole_run_achieved": flops_run / (ms * 1e-3) / 1e12, "kernel": "tc_count_kernel<8,4,0>", "traffic": None,
                           "note": "frac = GEMM + counting kernel of sampled tiles (library CUDA events around its launch)"}
    if k_top and last.get("ids") is not None:
        res["top1_valid"] = bool((last["ids"][:, 0] >= 0).all().item())
    del P, Qs
    engine.release_eval_workspace()
    torch.cuda.empty_cache()
    return res


def bench_eval(torch, engine, dev, tc_peak, eval_users=1 << 20):
    """Full-rank leave-one-out evaluation (second headline metric): users/s with positions for HR@10/NDCG@10.
    Shapes: BASELINE.json configs[2] (yelp-sort shape, d=128) and a tile of configs[4] (10M items, d=256)."""
    out = {}

    def one(U, I, d, tag, exact_too, reps, topks=()):
        g = torch.Generator(device=dev)
        g.manual_seed(2019)
        P = torch.randn((U, d), device=dev, generator=g) / d ** 0.5
        Q = torch.randn((I + 1, d), device=dev, generator=g) / d ** 0.5
        test = torch.randint(0, I, (U,), device=dev, dtype=torch.int32, generator=g)
        # exclusion set = 16 random "train" items + the held-out item per user (sorted, unique)
        tr = torch.randint(0, I, (U, 16), device=dev, dtype=torch.int32, generator=g)
        rows = torch.cat([tr, test[:, None]], dim=1).cpu().numpy()
        lists = [np.unique(r) for r in rows]
        ptr = np.zeros(U + 1, np.int64)
        ptr[1:] = np.cumsum([x.size for x in lists])
        idx = np.concatenate(lists).astype(np.int32)
        t = lambda a, dt: torch.from_numpy(a).to(device=dev, dtype=dt)
        a = [P, Q, torch.arange(U, device=dev, dtype=torch.int32), test, 0, I, t(ptr, torch.int64), t(idx, torch.int32)]
        res = {"users": U, "items": I, "d": d}

        def timed(fn):
            fn()
            torch.cuda.synchronize()
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            for _ in range(reps):
                r = fn()
            e1.record()
            torch.cuda.synchronize()
            return e0.elapsed_time(e1) / reps, r

        ms, r = timed(lambda: engine.eval_fullrank_tc(*a, check=False)[0])
        pos_tc = r
        _, namb = engine.eval_fullrank_tc(*a)
        # the GEMM + counting kernel alone: CUDA events recorded by the library around that launch (its own stream)
        engine.eval_tc_timing(True)
        kms = []
        for _ in range(max(2, reps)):
            engine.eval_fullrank_tc(*a, check=False)
            torch.cuda.synchronize()
            kms.append(engine.eval_tc_timing(True))
        engine.eval_tc_timing(False)
        k_ms = float(np.mean(kms[1:]))
        flops = 2.0 * U * I * 3 * d
        issued_call = flops / (ms * 1e-3) / 1e12
        issued_kernel = flops / (k_ms * 1e-3) / 1e12
        res["tensor_core"] = {"users_per_s": U / (ms * 1e-3), "ms": ms, "ambiguous_pairs_rescored": namb,
                              "gemm_kernel_ms": k_ms,
                              "roofline": {"bound": "tensor", "achieved": issued_kernel, "peak": tc_peak, "unit": "TFLOP/s",
                                           "frac": issued_kernel / tc_peak, "traffic": None, "kernel": "tc_count_kernel",
                                           "whole_call_achieved": issued_call, "whole_call_frac": issued_call / tc_peak,
                                           "note": "issued bf16 flops = 3 x useful (hi*hi + hi*lo + lo*hi split); "
                                                   "achieved/frac = the tcgen05 GEMM + counting kernel alone (CUDA events "
                                                   "around its launch); whole_call_* also carries the operand split, the "
                                                   "exact re-scoring and the exclusion correction"}}
        if exact_too:
            ms_e, pos_e = timed(lambda: engine.eval_fullrank(*a, 0, exact=True)[0])
            res["exact_fp32"] = {"users_per_s": U / (ms_e * 1e-3), "ms": ms_e, "fp32_tflops": 2.0 * U * I * d / (ms_e * 1e-3) / 1e12}
            res["positions_identical"] = bool(torch.equal(pos_tc, pos_e))
            ms_k, _ = timed(lambda: engine.eval_fullrank(*a, 10, exact=True)[0])
            res["exact_fp32_with_top10"] = {"users_per_s": U / (ms_k * 1e-3), "ms": ms_k}
        for K in topks:
            # top-K ids on the tensor-core path (counting pass with group maxima -> threshold -> candidate pass -> exact
            # re-scoring -> selection): two GEMM passes, same ids as the exact kernel (tests/test_gpu_eval_tc.py)
            ms_t, _ = timed(lambda: engine.eval_fullrank_tc(*a, check=False, k_top=K)[0])
            _, ids_t, _, info = engine.eval_fullrank_tc(*a, k_top=K)
            ent = {"users_per_s": U / (ms_t * 1e-3), "ms": ms_t, "candidates_rescored": info["topk_candidates"],
                   "exact_fallback_users": info["exact_fallback_users"],
                   "issued_tflops_whole_call": 2 * flops / (ms_t * 1e-3) / 1e12}
            if exact_too:
                _, ids_e, _ = engine.eval_fullrank(*a, K, exact=True)
                ent["ids_identical_to_exact_kernel"] = bool(torch.equal(ids_t, ids_e))
            res["tensor_core_top%d" % K] = ent
        res["hr10"] = float((pos_tc < 10).float().mean().item())
        out[tag] = res

    one(25677, 25815, 128, "yelp_shape_d128", True, 3, topks=(10, 100))
    one(4096, 10_000_000, 256, "config5_tile_4096_users_x_10M_items_d256", False, 1, topks=(100,))

    if eval_users > 0:
        out["config5_%d_users_x_10M_items_d256" % eval_users] = bench_eval_config5(torch, engine, dev, tc_peak, eval_users)
        ku = min(eval_users, 1 << 18)
        out["config5_top100_%d_users_x_10M_items_d256" % ku] = bench_eval_config5(torch, engine, dev, tc_peak, ku, k_top=100)

    def sampled(U, I, d, C, tag):
        """configs[1] (pinterest shape): He protocol, 99 sampled negatives + the held-out item per user
        (utils.py:244-254).  Algorithmic bytes 4 d (1 + C) per user; the 2.5 MB item table is L2-resident,
        so the figure is a gather rate, not an HBM roofline claim."""
        g = torch.Generator(device=dev)
        g.manual_seed(2019)
        P = torch.randn((U, d), device=dev, generator=g) / d ** 0.5
        Q = torch.randn((I, d), device=dev, generator=g) / d ** 0.5
        cand = torch.randint(0, I, (U * C,), device=dev, dtype=torch.int32, generator=g)
        ptr = torch.arange(0, (U + 1) * C, C, device=dev, dtype=torch.int64)
        users = torch.arange(U, device=dev, dtype=torch.int32)
        for _ in range(3):
            pos, _ = engine.eval_candidates(P, Q, users, ptr, cand)
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        reps = 20
        e0.record()
        for _ in range(reps):
            pos, _ = engine.eval_candidates(P, Q, users, ptr, cand)
        e1.record()
        torch.cuda.synchronize()
        ms = e0.elapsed_time(e1) / reps
        out[tag] = {"users": U, "items": I, "d": d, "candidates_per_user": C, "ms": ms, "users_per_s": U / (ms * 1e-3),
                    "gather_gb_per_s": 4.0 * d * (1 + C) * U / (ms * 1e-3) / 1e9,
                    "hr10": float((pos < 10).float().mean().item()),
                    "note": "item table (2.5 MB) is L2-resident: launch/L2 bound, no HBM roofline claimed"}

    sampled(55187, 9916, 64, 100, "config2_pinterest_shape_sampled_d64")
    out["reference_logs"] = "yelp-sort full-rank eval 250-285 users/s (TF1 CPU, BASELINE.md 1.2)"
    return out


if __name__ == "__main__":
    main()
