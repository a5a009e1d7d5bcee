#!/usr/bin/env python
"""bench.py — whole-body QP solves/sec on N B200s (one process per GPU), next to the CPU path.

A step = one pass of the hot path (2-level cascade + output recovery, what one control_loop tick
does: ref:src/ForceAcc.cpp:184-219) over one batch of synthetic states.  Workload (BASELINE.json):
  * one GPU (no torchrun):  configs[2], "QPPVM batched, 65536 states, WALK-MAN-like 33-DoF, 4 contacts
    (feet+hands)" with friction cones + torque limits: the largest single-GPU configuration;
  * under torchrun (N > 1): configs[3], ONE batch of 2^20 such states split over the N ranks (strong scaling;
    the path shards with no data-path collective).
configs[0], [1], [4] are parity-test cases; `--config i` still runs any of them.

    python bench.py --gpus 1 --steps 1000 --warmup 10
    python -m torch.distributed.run --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1 \
        --master-port P bench.py --gpus N --steps K --warmup W
    python bench.py --impl reference ...     # the CPU restatement (oracle) on the host cores
"""
from __future__ import annotations

import argparse
import json
import os
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

METRIC = "whole-body QP solves/sec (2-level cascade, KKT<=1e-6)"
N_BUF = 4     # default; main() sizes it so that the rotated inputs exceed the 126 MB L2 (4 x 45 MB at config 1)
# SURVEY.md 8(d) minimum-work FP64 model (FLOP per solve)
F_ALG = {(29, 2): 182e3, (33, 4): 335e3}


def parse():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=100)
    ap.add_argument("--warmup", type=int, default=10)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--config", type=int, default=-1,
                    help="BASELINE.json config index (default: 2 on one GPU, 3 = 2^20 states sharded under torchrun)")
    ap.add_argument("--batch", type=int, default=0, help="override records per step per GPU")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-latency", action="store_true")
    ap.add_argument("--fast", action="store_true", help="kernel experiments: device-resident value only, no e2e / roofline legs")
    ap.add_argument("--multi", type=int, default=0,
                    help="single-process leg: configs[3] over G GPUs through qppvm_multi_* (C++ / NCCL behind the C-ABI); prints its own line")
    return ap.parse_args()


def measured_peaks():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        try:
            return json.load(open(p)), "measured (MEASURED_PEAKS.json)"
        except Exception:
            pass
    return {"hbm_gbs": 6650.0}, "fallback (B200_PROFILING.md)"


class ClockSampler(threading.Thread):
    """Samples SM clock and throttle reasons through NVML.  Started before the warm-up (NVML initialisation is
    slower than a short timed region); one sample is taken synchronously when the timed region is armed, the thread
    keeps sampling every 10 ms until it is disarmed."""

    NAMES = {"hw_slowdown": 0x8, "sw_power_cap": 0x4, "hw_thermal_slowdown": 0x40, "sw_thermal_slowdown": 0x20,
             "hw_power_brake_slowdown": 0x80, "applications_clocks_setting": 0x2}

    def __init__(self, index):
        super().__init__(daemon=True)
        self.index, self.stop_flag, self.armed = index, False, False
        self.samples, self.reasons, self.max_mhz, self.nv, self.h = [], set(), None, None, None
        try:
            import pynvml as nv
            nv.nvmlInit()
            self.nv, self.h = nv, nv.nvmlDeviceGetHandleByIndex(index)
            self.max_mhz = nv.nvmlDeviceGetMaxClockInfo(self.h, nv.NVML_CLOCK_SM)
        except Exception as e:  # NVML unavailable: report that rather than invent clocks
            self.reasons.add("nvml_unavailable:%s" % type(e).__name__)

    def sample(self):
        if self.nv is None:
            return
        nv = self.nv
        try:
            self.samples.append(nv.nvmlDeviceGetClockInfo(self.h, nv.NVML_CLOCK_SM))
            try:
                r = nv.nvmlDeviceGetCurrentClocksEventReasons(self.h)
            except Exception:
                r = nv.nvmlDeviceGetCurrentClocksThrottleReasons(self.h)
            for k, bit in self.NAMES.items():
                if r & bit:
                    self.reasons.add(k)
        except Exception as e:
            self.reasons.add("nvml_error:%s" % type(e).__name__)

    def arm(self):
        self.sample()            # the GPU is busy with the warm-up's tail / idle-to-busy edge: kept as sample 0
        self.armed = True

    def disarm(self):
        self.armed = False
        self.stop_flag = True

    def run(self):
        while not self.stop_flag:
            if self.armed:
                self.sample()
            time.sleep(0.01)

    def result(self):
        return {"sm_mhz": float(np.median(self.samples)) if self.samples else None,
                "sm_max_mhz": self.max_mhz, "reasons": sorted(self.reasons), "samples": len(self.samples)}


def host_threads():
    """All host threads this process may use (torchrun exports OMP_NUM_THREADS=1, so the count is passed explicitly)."""
    try:
        return max(1, len(os.sched_getaffinity(0)))
    except AttributeError:
        return max(1, os.cpu_count() or 1)


def cpu_reference(desc, recs, target_s, mode):
    """Times the oracle (restated active-set path, not the qpOASES binary) on a bounded sample."""
    from oracle import oracle
    oracle.use_native()                                    # -march=native build made on this box (falls back to the portable one)
    threads = host_threads()
    if len(recs) < 2048:                                   # tiny workloads (single-tick configs): tile to a useful sample
        recs = np.tile(recs, ((2048 + len(recs) - 1) // len(recs), 1))
    t0 = time.perf_counter(); oracle.solve_batch(desc, recs[:512], mode=mode, threads=threads); rate = 512 / (time.perf_counter() - t0)
    n = int(max(rate * target_s, 512))
    done, t0 = 0, time.perf_counter()
    while done < n and time.perf_counter() - t0 < 2.0 * target_s:     # bounded by work AND by wall clock
        m = min(len(recs), n - done)
        oracle.solve_batch(desc, recs[:m], mode=mode, threads=threads)
        done += m
    dt = time.perf_counter() - t0
    return done / dt, threads, done, dt


def config_dict(args, cfg_name, desc, L, batch, world):
    return {"workload": cfg_name, "records_per_step_per_gpu": batch, "n_a": desc.n_a,
            "contacts": desc.n_contacts, "flags": desc.flags,
            "l2": "inputs rotate over %d distinct batches per GPU (%.0f MB > 126 MB L2)" % (N_BUF, N_BUF * batch * L.rec_doubles * 8 / 1e6),
            "parallelism": "batch sharded over %d GPU(s), no data-path collective" % world}


def run_reference(args, desc, L, cfg_name, batch):
    """--impl reference: the reference's own CPU path cannot be built here (OpenSoT/qpOASES/XBotCore absent),
    so this arm times the oracle port with the reference's numerics (formed H + Cholesky) on all host threads."""
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    from qppvm_b200 import gen
    from oracle import oracle
    recs = gen.generate(desc, min(batch, 8192), gen.config_seed(args.config))   # a step samples from these
    native = oracle.use_native()                           # -march=native build made on this box
    threads = host_threads()
    ncal = min(256, len(recs))
    t0 = time.perf_counter(); oracle.solve_batch(desc, recs[:ncal], mode=oracle.FACTOR_CHOLESKY, threads=threads)
    rate = ncal / (time.perf_counter() - t0)
    budget = 120.0 / max(1, args.steps + args.warmup)            # whole run ~<= 2 min
    sample = int(max(1, min(len(recs), rate * budget)))
    for _ in range(args.warmup):
        oracle.solve_batch(desc, recs[:sample], mode=oracle.FACTOR_CHOLESKY, threads=threads)
    t0 = time.perf_counter()
    for _ in range(args.steps):
        out, _ = oracle.solve_batch(desc, recs[:sample], mode=oracle.FACTOR_CHOLESKY, threads=threads)
    dt = time.perf_counter() - t0
    o = oracle.split_out(desc, out)
    good = float(((o["status"] == 0) & (o["kkt"].max(axis=1) <= 1e-6)).mean())
    val = args.steps * sample * good / dt
    line = {"impl": "reference", "metric": METRIC, "value": val, "unit": "solves/s", "n_gpus": args.gpus,
            "steps": args.steps, "warmup": args.warmup, "ms_per_step": dt / args.steps * 1e3,
            "higher_is_better": True, "scaling": "strong" if args.config == 3 else "weak", "vs_baseline": None,
            "dtype": "f64", "data": "synthetic",
            "config": config_dict(args, cfg_name, desc, L, batch, args.gpus),
            "cpu_baseline": {"value": val, "unit": "solves/s", "cores": threads, "kind": "port",
                             "sample": "%d records/step of the %d-record batch, restated active-set path "
                                       "(qpOASES semantics, formed H + Cholesky), not the qpOASES binary; %s build"
                                       % (sample, batch, "-O3 -march=native" if native else "-O3 -march=x86-64-v3")},
            "e2e": {"value": val, "unit": "solves/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
            "gpu_launches": 0}
    print(json.dumps(line), flush=True)


def run_multi(args):
    """configs[3] (2^20 states, 33-DoF, 4 contacts, cones + tau limits) over G GPUs from ONE process through the C++ entry
    points: records resident on the root GPU (NCCL scatter -> solve -> gather, pipelined), and host buffers (records /
    compact states; every GPU pulls its own block over PCIe)."""
    import torch
    from qppvm_b200 import api, gen
    from qppvm_b200.layout import CONFIGS, layout
    G = args.multi
    desc = CONFIGS[3]["desc"]
    L = layout(desc)
    B = args.batch or CONFIGS[3]["batch"]
    steps = max(1, min(args.steps, 10))
    rob = gen.robot_for(desc.n_a)
    contacts = (rob.foot + rob.hand)[:desc.n_contacts]
    s0 = api.Solver(desc)
    s0.set_robot(rob, contacts)
    st = gen.generate_states(desc, B, gen.config_seed(3))
    recs = s0.records_from_states(torch.from_numpy(st).cuda())
    torch.cuda.synchronize()
    out = torch.empty((B, L.out_doubles), dtype=torch.float64, device="cuda:0")
    res = {}
    for g in sorted({1, G}):
        m = api.MultiSolver(desc, list(range(g)))
        m.set_robot(rob, contacts)
        for _ in range(2):
            m.solve_batch(recs, out=out)
        t0 = time.perf_counter()
        for _ in range(steps):
            m.solve_batch(recs, out=out)
        dt = (time.perf_counter() - t0) / steps
        o = api.split_out(L, out[:65536].cpu().numpy())
        good = float(((o["status"] == 0) & (o["kkt"].max(axis=1) <= 1e-6)).mean())
        entry = {"root_resident_solves_per_s": B * good / dt, "ms_per_step": dt * 1e3, "nccl_calls_per_step": m.nccl_calls // (steps + 2)}
        # the same with the compact states on the root GPU: the states travel, every GPU runs the front end on its chunks
        dst = torch.from_numpy(st).cuda()
        for _ in range(2):
            m.solve_states(dst, out=out)
        t0 = time.perf_counter()
        for _ in range(steps):
            m.solve_states(dst, out=out)
        dts = (time.perf_counter() - t0) / steps
        o = api.split_out(L, out[:65536].cpu().numpy())
        good_s = float(((o["status"] == 0) & (o["kkt"].max(axis=1) <= 1e-6)).mean())
        entry["root_states_solves_per_s"] = B * good_s / dts
        del dst
        hst = torch.from_numpy(st).pin_memory()
        hout = torch.empty((B, L.out_doubles), dtype=torch.float64).pin_memory()
        m.solve_states_host_ptr(hst.data_ptr(), hout.data_ptr(), B)
        t0 = time.perf_counter()
        for _ in range(steps):
            m.solve_states_host_ptr(hst.data_ptr(), hout.data_ptr(), B)
        entry["host_states_solves_per_s"] = B / ((time.perf_counter() - t0) / steps)
        if g == G:
            hrec = recs.cpu().pin_memory()
            m.solve_batch_host_ptr(hrec.data_ptr(), hout.data_ptr(), B)
            t0 = time.perf_counter()
            for _ in range(steps):
                m.solve_batch_host_ptr(hrec.data_ptr(), hout.data_ptr(), B)
            entry["host_records_solves_per_s"] = B / ((time.perf_counter() - t0) / steps)
            del hrec
        res["gpus_%d" % g] = entry
        m.close()
    line = {"leg": "multi_gpu_single_process", "api": "qppvm_multi_solve_batch / _solve_states / _solve_states_host / _solve_batch_host",
            "workload": "configs[3]: %d states, 33-DoF, 4 contacts, cones + tau limits" % B, "gpus": G, "steps": steps,
            "results": res}
    if G > 1:
        a, b = res["gpus_1"], res["gpus_%d" % G]
        line["scatter_gather_efficiency_vs_1gpu"] = b["root_resident_solves_per_s"] / (G * a["root_resident_solves_per_s"])
        line["host_states_efficiency_vs_1gpu"] = b["host_states_solves_per_s"] / (G * a["host_states_solves_per_s"])
        line["root_states_efficiency_vs_1gpu"] = b["root_states_solves_per_s"] / (G * a["root_states_solves_per_s"])
    print(json.dumps(line), flush=True)


def main():
    args = parse()
    if args.multi:
        return run_multi(args)
    from qppvm_b200.layout import CONFIGS, layout
    if args.config < 0:
        args.config = 3 if max(args.gpus, int(os.environ.get("WORLD_SIZE", "1"))) > 1 else 2
    cfg = CONFIGS[args.config]
    desc = cfg["desc"]
    L = layout(desc)
    world_env = int(os.environ.get("WORLD_SIZE", "1"))
    # configs[3] is ONE batch of 2^20 states sharded over the GPUs (strong scaling); every other config keeps the
    # per-GPU batch fixed as GPUs are added (weak scaling)
    strong = args.config == 3
    batch = args.batch or (cfg["batch"] // world_env if strong else min(cfg["batch"], 65536))
    cfg_name = "configs[%d]: %s" % (args.config, cfg["name"])
    global N_BUF
    N_BUF = int(min(4, max(1, -(-int(1.3 * 126e6) // (batch * L.rec_doubles * 8)))))
    if args.impl == "reference":
        return run_reference(args, desc, L, cfg_name, batch)

    import torch
    import torch.distributed as dist
    from qppvm_b200 import api, gen
    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if not torch.cuda.is_available():
        raise SystemExit("bench.py: no CUDA device (the product has no CPU path; use --impl reference for the CPU arm)")
    torch.cuda.set_device(local)
    if world > 1:
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    import dataclasses
    desc = dataclasses.replace(desc, device=local)
    solver = api.Solver(desc)
    dev = torch.device("cuda", local)

    # ---- synthetic inputs: N_BUF distinct batches per rank, resident in HBM and mirrored in pinned host memory.
    # The states come from the seeded generator (cheap); the rigid-body dynamics that turn them into records run on
    # the device front end (tests/test_rbd_frontend.py: equal to the numpy dynamics to 1e-11), so that the 65 536- and
    # 2^20-record configs start in seconds.  A sample is re-checked against the numpy path right here.
    n_in = N_BUF * batch
    if desc.kind == 1:
        rob = gen.robot_for(desc.n_a)
        solver.set_robot(rob, (rob.foot + rob.hand)[:desc.n_contacts])
        h_states_all = gen.generate_states(desc, n_in, gen.config_seed(args.config), start=rank * n_in)
        d_all = solver.records_from_states(torch.from_numpy(h_states_all).to(dev))
        torch.cuda.synchronize(dev)
        chk = gen.records_from_states(desc, h_states_all[:32])
        err = np.abs(d_all[:32].cpu().numpy() - chk).max() / max(1.0, np.abs(chk).max())
        assert err < 1e-10, "device front end disagrees with the numpy dynamics: %g" % err
        d_recs = [d_all[i * batch:(i + 1) * batch] for i in range(N_BUF)]
        # records in pinned host memory for the record-path e2e legs: whole batches up to 131 072 records (2.4 GB); the
        # 2^20-state workload runs those legs on a 131 072-record part of the shard (its headline e2e ships states)
        hb = min(batch, 131072)
        n_host = min(n_in, max(hb, 16384)) if batch <= 65536 else hb
        pinned = torch.empty((n_host, L.rec_doubles), dtype=torch.float64).pin_memory()
        pinned.copy_(d_all[:n_host])
        host_recs = pinned.numpy()
    else:
        hb = batch
        host_recs = gen.generate(desc, n_in, gen.config_seed(args.config), start=rank * n_in)
        pinned = torch.from_numpy(host_recs).pin_memory()
        d_recs = [pinned[i * batch:(i + 1) * batch].to(dev, non_blocking=False).contiguous() for i in range(N_BUF)]
    n_hbuf = max(1, pinned.shape[0] // hb)                    # host-side batches available for the e2e legs
    d_out = [torch.empty((batch, L.out_doubles), dtype=torch.float64, device=dev) for _ in range(N_BUF)]
    h_out = torch.empty((batch, L.out_doubles), dtype=torch.float64).pin_memory()
    stream = torch.cuda.current_stream(dev)

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize(dev)

    # ---- device-resident throughput ("value") ------------------------------------------------------
    sampler = ClockSampler(local); sampler.start()
    for i in range(args.warmup):
        solver.solve_batch(d_recs[i % N_BUF], out=d_out[i % N_BUF])
    barrier()
    # clocks: one NVML sample right here (GPU still busy with the warm-up's tail), then every 10 ms while the K timed
    # steps run (a configs[2] step is tens of milliseconds)
    sampler.arm()
    l0 = solver.kernel_launches
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record(stream)
    for i in range(args.steps):
        solver.solve_batch(d_recs[i % N_BUF], out=d_out[i % N_BUF])
    e1.record(stream)
    barrier()
    sampler.disarm(); sampler.join()
    launches = solver.kernel_launches - l0
    ms = torch.tensor([e0.elapsed_time(e1)], dtype=torch.float64, device=dev)
    # every counted solve must have converged with KKT <= 1e-6 (in-kernel certificate)
    good, total, kkt_max = 0, 0, 0.0
    for i in range(min(N_BUF, args.steps)):
        g = api.split_out(L, d_out[i].cpu().numpy())
        ok = (g["status"] == 0) & (g["kkt"].max(axis=1) <= 1e-6)
        good += int(ok.sum()); total += len(ok)
        kkt_max = max(kkt_max, float(g["kkt"][g["status"] == 0].max()) if (g["status"] == 0).any() else 0.0)
    frac = torch.tensor([good / max(1, total)], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(ms, op=dist.ReduceOp.MAX)
        dist.all_reduce(frac, op=dist.ReduceOp.MIN)
    t_s = ms.item() * 1e-3
    value = world * args.steps * batch * frac.item() / t_s

    # per-kernel share of a step (events around every launch; a separate short pass, not the timed region above)
    solver.kernel_timing(True)
    for i in range(min(3, args.steps)):
        solver.solve_batch(d_recs[i % N_BUF], out=d_out[i % N_BUF])
    kms, kn = solver.kernel_timing(False)
    ksplit = {"prepare_ms_per_step": kms[0] / max(1, min(3, args.steps)), "solve_ms_per_step": kms[1] / max(1, min(3, args.steps)),
              "certify_ms_per_step": kms[2] / max(1, min(3, args.steps)), "launches": [int(v) for v in kn],
              "solve_share": float(kms[1] / max(1e-12, kms.sum())),
              "what": "CUDA events on the launching stream around qp_factor_kernel / qp_solve_kernel / qp_certify_kernel"}
    if args.fast:
        if rank == 0:
            print(json.dumps({"value": value, "ms_per_step": t_s / args.steps * 1e3, "converged_frac": frac.item(),
                              "kkt_max": kkt_max, "n_gpus": world, "config": cfg_name, "fast": True}), flush=True)
        if world > 1:
            dist.destroy_process_group()
        return

    # ---- end to end through the reference-facing C-ABI call with HOST buffers ("e2e") ----------------
    e2e_steps = max(3, min(args.steps, 200))
    rec_ptr = [pinned[i * hb:(i + 1) * hb].data_ptr() for i in range(n_hbuf)]
    for i in range(3):
        solver.solve_batch_host_ptr(rec_ptr[i % n_hbuf], h_out.data_ptr(), hb)
    barrier()
    t0 = time.perf_counter()
    for i in range(e2e_steps):
        solver.solve_batch_host_ptr(rec_ptr[i % n_hbuf], h_out.data_ptr(), hb)   # H2D + solve + D2H, synchronous
    torch.cuda.synchronize(dev)
    t_e2e = torch.tensor([time.perf_counter() - t0], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(t_e2e, op=dist.ReduceOp.MAX)
    ge = api.split_out(L, h_out.numpy()[:hb])
    e2e_ok = float(((ge["status"] == 0) & (ge["kkt"].max(axis=1) <= 1e-6)).mean())
    e2e_val = world * e2e_steps * hb * e2e_ok / t_e2e.item()

    # pipelined variant: the async entry point, two output buffers in flight, one sync at the end; every step still
    # copies its own records in and its own results out inside the timed region.  Reported beside `e2e`, not as it.
    h_out2 = [h_out, torch.empty_like(h_out).pin_memory()]
    for i in range(3):
        solver.solve_batch_host_async_ptr(rec_ptr[i % n_hbuf], h_out2[i % 2].data_ptr(), hb)
    solver.host_sync()
    barrier()
    t0 = time.perf_counter()
    for i in range(e2e_steps):
        solver.solve_batch_host_async_ptr(rec_ptr[i % n_hbuf], h_out2[i % 2].data_ptr(), hb)
    solver.host_sync()
    t_pipe = torch.tensor([time.perf_counter() - t0], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(t_pipe, op=dist.ReduceOp.MAX)
    gp = api.split_out(L, h_out2[(e2e_steps - 1) % 2].numpy()[:hb])
    pipe_ok = float(((gp["status"] == 0) & (gp["kkt"].max(axis=1) <= 1e-6)).mean())
    e2e_pipe = {"value": world * e2e_steps * hb * pipe_ok / t_pipe.item(), "unit": "solves/s",
                "h2d_bytes_per_step": hb * L.rec_doubles * 8, "d2h_bytes_per_step": hb * L.out_bytes,
                "steps": e2e_steps, "api": "qppvm_solve_batch_host_async + qppvm_host_sync (steps overlap)"}

    # ---- end to end from compact STATES: rigid-body front end + solve on the device; the host
    # ships 1 KB states instead of 11-18 KB records.  This is the headline `e2e` at every N (see the end of main).
    e2e_states = None
    if desc.kind == 1:
        h_states = torch.from_numpy(h_states_all).pin_memory()
        sd = h_states.shape[1]
        st_ptr = [h_states[i * batch:(i + 1) * batch].data_ptr() for i in range(N_BUF)]
        for i in range(3):
            solver.solve_states_host_ptr(st_ptr[i % N_BUF], h_out.data_ptr(), batch)
        barrier()
        t0 = time.perf_counter()
        for i in range(e2e_steps):
            solver.solve_states_host_ptr(st_ptr[i % N_BUF], h_out.data_ptr(), batch)
        torch.cuda.synchronize(dev)
        t_st = torch.tensor([time.perf_counter() - t0], dtype=torch.float64, device=dev)
        if world > 1:
            dist.all_reduce(t_st, op=dist.ReduceOp.MAX)
        gs = api.split_out(L, h_out.numpy())
        st_ok = float(((gs["status"] == 0) & (gs["kkt"].max(axis=1) <= 1e-6)).mean())
        e2e_states = {"value": world * e2e_steps * batch * st_ok / t_st.item(), "unit": "solves/s",
                      "h2d_bytes_per_step": batch * sd * 8, "d2h_bytes_per_step": batch * L.out_bytes, "steps": e2e_steps,
                      "api": "qppvm_solve_states_host (states -> on-device rigid-body dynamics -> solve -> torques)"}
        for i in range(3):
            solver.solve_states_host_async_ptr(st_ptr[i % N_BUF], h_out2[i % 2].data_ptr(), batch)
        solver.host_sync()
        barrier()
        t0 = time.perf_counter()
        for i in range(e2e_steps):
            solver.solve_states_host_async_ptr(st_ptr[i % N_BUF], h_out2[i % 2].data_ptr(), batch)
        solver.host_sync()
        t_sp = torch.tensor([time.perf_counter() - t0], dtype=torch.float64, device=dev)
        if world > 1:
            dist.all_reduce(t_sp, op=dist.ReduceOp.MAX)
        gsp = api.split_out(L, h_out2[(e2e_steps - 1) % 2].numpy())
        sp_ok = float(((gsp["status"] == 0) & (gsp["kkt"].max(axis=1) <= 1e-6)).mean())
        e2e_states["pipelined_value"] = world * e2e_steps * batch * sp_ok / t_sp.item()

    # ---- closed-loop rollout on the device: front end -> solve -> integrate per control period,
    # nothing crossing PCIe.  Rollouts of ROLL_T periods restart from the pristine states (a 4 MB device copy, timed).
    rollout = None
    if desc.kind == 1:
        ROLL_T, dt = 10, 1e-3
        d_st0 = torch.from_numpy(h_states_all[:batch]).to(dev)
        d_st = d_st0.clone()
        n_roll = max(1, min(args.steps, 200) // ROLL_T)
        solver.rollout_states(d_st, 3, dt, out=d_out[0])
        barrier()
        r0, r1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        r0.record(stream)
        for i in range(n_roll):
            d_st.copy_(d_st0)
            solver.rollout_states(d_st, ROLL_T, dt, out=d_out[0])
        r1.record(stream)
        barrier()
        t_roll = torch.tensor([r0.elapsed_time(r1) * 1e-3], dtype=torch.float64, device=dev)
        if world > 1:
            dist.all_reduce(t_roll, op=dist.ReduceOp.MAX)
        gr = api.split_out(L, d_out[0].cpu().numpy())
        rollout = {"value": world * n_roll * ROLL_T * batch / t_roll.item(), "unit": "state-ticks/s",
                   "ticks_per_rollout": ROLL_T, "rollouts": n_roll, "dt": dt,
                   "last_tick_converged_frac": float(((gr["status"] == 0) & (gr["kkt"].max(axis=1) <= 1e-6)).mean()),
                   "api": "qppvm_rollout_states (device-resident states; 3 kernels per tick and lane, 4 lanes)"}

    # ---- N > 1: the same shards fed from rank 0 over NCCL (scatter records, gather outputs), device to device
    sg = None
    if world > 1:
        from qppvm_b200 import shard
        root_recs = torch.cat([d_recs[0]] * world) if rank == 0 else None      # world * batch records on the root GPU
        sg_steps = max(3, min(args.steps, 50))
        for i in range(2):
            shard.solve_sharded(lambda r: solver.solve_batch(r)[0], root_recs, world * batch, L.rec_doubles, dev)
        barrier()
        t0 = time.perf_counter()
        for i in range(sg_steps):
            shard.solve_sharded(lambda r: solver.solve_batch(r)[0], root_recs, world * batch, L.rec_doubles, dev)
        torch.cuda.synchronize(dev)
        t_sg = torch.tensor([time.perf_counter() - t0], dtype=torch.float64, device=dev)
        dist.all_reduce(t_sg, op=dist.ReduceOp.MAX)
        sg = {"value": world * batch * sg_steps / t_sg.item(), "unit": "solves/s", "steps": sg_steps,
              "what": "NCCL scatter of records from rank 0 + solve + gather of outputs to rank 0 (root-egress bound)"}

    if rank != 0:
        if world > 1:
            dist.destroy_process_group()
        return

    # ---- roofline of the step's kernel(s) ----------------------------------------------------------------
    # Unstaged shapes run two launches per step (prepare: factorisation + equality rows; solve: active set + KKT):
    # the roofline is quoted for the pair over the step time measured by the CUDA events, which bounds the dominant
    # kernel's own figure from below; profiles/ holds the per-kernel split (ncu).
    peaks, peak_src = measured_peaks()
    launch_s = t_s / max(1, args.steps)
    per_step = max(1, int(round(launches / max(1, args.steps))))
    kname = "qp_solve_kernel<ForceAcc<%d,%d,%d>>" % (desc.n_a, desc.n_contacts, desc.flags) if desc.kind == 1 else "qp_solve_kernel<Torque<%d>>" % desc.n_a
    if per_step >= 2:
        kname = ("qp_factor_kernel + " + kname + " + qp_certify_kernel (%d launches per step: the batch runs in workspace-sized "
                 "passes of one prepare, one solve and one certify launch, timed together; kernel_split has the shares)" % per_step)
    alg_bytes = L.algorithmic_bytes() * batch
    fp64_peak = solver.fp64_peak_tflops()
    f_alg = F_ALG.get((desc.n_a, desc.n_contacts))
    traffic = None                                         # ncu dram bytes per launch, scaled to this launch's record count
    tpath = os.path.join(ROOT, "profiles", "traffic_config%d.json" % args.config)
    if os.path.exists(tpath):
        try:
            tj = json.load(open(tpath))
            traffic = int(tj["dram_bytes_per_launch"] / tj["records_per_launch"] * batch)
        except Exception:
            traffic = None
    roofline_hbm = {"bound": "hbm", "achieved": alg_bytes / launch_s / 1e9, "peak": peaks["hbm_gbs"], "unit": "GB/s",
                    "frac": alg_bytes / launch_s / 1e9 / peaks["hbm_gbs"], "traffic": traffic, "peak_source": peak_src,
                    "bytes_per_solve": L.algorithmic_bytes(),
                    "note": "not the binding roof: arithmetic intensity 15-18 FLOP/B vs an FP64 ridge of ~5.5 FLOP/B"}
    f_alg = f_alg or 0.0
    ach = f_alg * batch / launch_s / 1e12
    # The binding roof of this path is the FP64 FMA pipe: algorithmic FLOPs per launch (minimum-work
    # model) / launch time, against the FP64 peak measured in this run (MEASURED_PEAKS.json has no FP64
    # figure).  `traffic` stays the ncu DRAM bytes of the same launch(es).
    roofline = {"bound": "fp64", "achieved": ach, "peak": fp64_peak, "unit": "TFLOP/s",
                "frac": ach / fp64_peak if fp64_peak else None, "traffic": traffic,
                "peak_source": "measured in this run (register-resident DFMA kernel; no FP64 figure in MEASURED_PEAKS.json)",
                "flop_per_solve": f_alg, "kernel": kname, "launches_per_step": per_step,
                "note": "FP64 CUDA-core path (tcgen05 has no FP64): `bound` names the binding roof; HBM roof in roofline_hbm"}

    e2e_records = {"value": e2e_val, "unit": "solves/s", "h2d_bytes_per_step": hb * L.rec_doubles * 8,
                   "d2h_bytes_per_step": hb * L.out_bytes, "steps": e2e_steps, "records_per_step_per_gpu": hb,
                   "api": "qppvm_solve_batch_host (records in pinned host buffers in, outputs out)"}
    # headline e2e, the same call at every N: compact states in, torques out (qppvm_solve_states_host) -- what one
    # control_loop tick does between sense() and move() (ref:src/ForceAcc.cpp:167-253: model update, stack update, solve,
    # torque recovery), batched.  The reference arm only times the QP part of that tick, from records: the comparison is
    # conservative.  The record path (the boundary after model->update(), 11-18 KB per problem over PCIe; host-DRAM
    # bound with 8 processes on one box) is reported beside it as `e2e_records`.  Shapes without a rigid-body front
    # end (Torque kind) ship records.
    if e2e_states:
        e2e_head = {k: e2e_states[k] for k in ("value", "unit", "h2d_bytes_per_step", "d2h_bytes_per_step", "steps", "api")}
    else:
        e2e_head = e2e_records
    line = {"metric": METRIC, "value": value, "unit": "solves/s", "n_gpus": world, "steps": args.steps,
            "warmup": args.warmup, "ms_per_step": t_s / args.steps * 1e3, "higher_is_better": True,
            "scaling": "strong" if strong else "weak", "vs_baseline": None, "dtype": "f64", "data": "synthetic",
            "config": config_dict(args, cfg_name, desc, L, batch, world),
            "clocks": sampler.result(),
            "e2e": e2e_head,
            "gpu_launches": int(launches),
            "roofline": roofline, "roofline_hbm": roofline_hbm,
            "converged_frac": frac.item(), "kkt_max": kkt_max}
    if sg:
        line["scatter_gather"] = sg
    line["e2e_records"] = e2e_records
    line["kernel_split"] = ksplit
    line["e2e_pipelined"] = e2e_pipe
    if rollout:
        line["rollout"] = rollout
    if e2e_states:
        line["e2e_states"] = e2e_states

    if world == 1 and not args.no_latency:
        # single-tick latency (the metric's second half): configs[4], one QP per 1 kHz control tick, host in / host out
        # per tick through qppvm_solve_one (resident prepare -> solve -> certify chain, hot start from the previous
        # tick).  Workload: ONE robot followed over 10 000 consecutive ticks (slow random walk of q, qdot), which is what
        # a control loop feeds the solver; the CPU port solves the same ticks on one host core.
        N_TICKS = 10000
        d4 = dataclasses.replace(CONFIGS[4]["desc"], device=local)
        s4 = api.Solver(d4)
        st0 = gen.generate_states(d4, 1, gen.config_seed(4))[0]
        rng = np.random.default_rng(gen.config_seed(4))
        st = np.repeat(st0[None], N_TICKS, axis=0)
        st[:, :d4.n_a] += np.cumsum(rng.normal(0.0, 2e-3, (N_TICKS, d4.n_a)), axis=0)
        st[:, d4.n_a:2 * d4.n_a] += np.cumsum(rng.normal(0.0, 5e-3, (N_TICKS, d4.n_a)), axis=0)
        r4 = gen.records_from_states(d4, st)
        L4 = layout(d4)
        o4 = np.empty(L4.out_doubles)
        for i in range(300):
            s4.solve_one(r4[i], o4)
        s4.reset_warm()
        lat = np.empty(N_TICKS); ok4 = 0
        for i in range(N_TICKS):
            t0 = time.perf_counter(); s4.solve_one(r4[i], o4); lat[i] = time.perf_counter() - t0
            tr = api.split_out(L4, o4[None])
            ok4 += int(tr["status"][0] == 0 and tr["kkt"].max() <= 1e-6)
        stages = None
        try:
            stg = np.zeros(6)
            for i in range(200):
                s4.solve_one(r4[i], o4)
                stg += np.diff(s4.tick_stamps().astype(np.int64))
            stg /= 200e3
            stages = {"copy_us": stg[0], "prepare_us": stg[1], "solve_us": stg[3], "certify_publish_us": stg[5],
                      "handoffs_us": stg[2] + stg[4], "chain_us": float(stg.sum())}
        except Exception:
            pass
        line["single_tick"] = {"p50_us": float(np.percentile(lat, 50) * 1e6), "p99_us": float(np.percentile(lat, 99) * 1e6),
                               "max_us": float(lat.max() * 1e6), "ticks": N_TICKS, "converged_frac": ok4 / N_TICKS,
                               "device_stages": stages,
                               "workload": "configs[4]: one robot over 10 000 consecutive ticks, one QP per tick through "
                                           "qppvm_solve_one (host in / host out, resident kernels, hot start)"}
        if not args.no_cpu_baseline:                       # the same ticks on one host core (restated active-set path)
            from oracle import oracle
            oracle.use_native()
            n_cpu = 2000
            clat, hlat = np.empty(n_cpu), np.empty(n_cpu)
            for i in range(n_cpu + 20):
                t0 = time.perf_counter(); oracle.solve_batch(d4, r4[i:i + 1], mode=oracle.FACTOR_CHOLESKY, threads=1)
                if i >= 20:
                    clat[i - 20] = time.perf_counter() - t0
            wcpu = np.zeros(8, dtype=np.uint32)
            for i in range(n_cpu + 20):                    # the same ticks, hot-started from the previous tick's working sets
                t0 = time.perf_counter(); oracle.solve_sequence(d4, r4[i:i + 1], wcpu, mode=oracle.FACTOR_CHOLESKY)
                if i >= 20:
                    hlat[i - 20] = time.perf_counter() - t0
            line["single_tick"]["cpu_port_p50_us"] = float(np.percentile(clat, 50) * 1e6)
            line["single_tick"]["cpu_port_p99_us"] = float(np.percentile(clat, 99) * 1e6)
            line["single_tick"]["cpu_port_hot_p50_us"] = float(np.percentile(hlat, 50) * 1e6)
            line["single_tick"]["cpu_port_hot_p99_us"] = float(np.percentile(hlat, 99) * 1e6)
            line["single_tick"]["cpu_port_ticks"] = n_cpu
            line["single_tick"]["cpu_port"] = ("oracle port built -O3 -march=native on this box, 1 core, called through ctypes like the GPU "
                                               "path: cold start per tick, and hot-started from the previous tick (oracle_solve_sequence)")
    if world == 1 and not args.no_cpu_baseline:
        from oracle import oracle
        v, threads, n_done, dt = cpu_reference(desc, host_recs[:65536], 12.0, oracle.FACTOR_CHOLESKY)
        t0 = time.perf_counter(); oracle.solve_batch(desc, host_recs[:2048], mode=oracle.FACTOR_CHOLESKY, threads=1)
        v1 = 2048 / (time.perf_counter() - t0)
        line["cpu_baseline"] = {"value": v, "unit": "solves/s", "cores": threads, "kind": "port",
                                "single_thread_value": v1,
                                "sample": "%d records of the same workload in %.1f s; restated active-set path "
                                          "(qpOASES semantics), not the qpOASES binary; built -O3 -march=native on this box; "
                                          "single_thread_value: 2048 records on one core" % (n_done, dt)}
    print(json.dumps(line), flush=True)
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
