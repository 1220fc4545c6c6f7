#!/usr/bin/env python
"""bench.py - headline benchmark of the B200 alignment engine (contract: see the task description).

    python bench.py --gpus N --steps K --warmup W          # our arm (under torchrun for N > 1)
    python bench.py --impl reference --gpus N --steps K --warmup W   # the reference's CPU cores on host threads

Workload (BASELINE.json configs[1]): synthetic paired-end soft-clip re-mapping,
StripedSmithWaterman(query = 1000/1500/2000 bp reference window, 2,-8,6,1)(target = 15-300 bp soft clip),
score + begin/end coordinates + cigar for every pair (flag = 1, what dysgu's wrapper always passes).
One step = one pass of the whole pipeline over one batch of pairs (per GPU).

Other BASELINE.json shapes run through the same code with --workload (parity-test shapes, reported in
profiles/; the driver's bench line is the default workload):
    pe_contigs     SSW (2,-3,10,1) paired-end contigs 100-500 bp           consensus.pyx:977-1018
    long_contigs   SSW (2,-3,10,1) long-read contigs 1-10 kb (configs 4/5) merge_svs.pyx:355-382
    remap_chain    edlib HW/locations clip vs 1000 bp window (config 2a)   re_map.py:209
    hifi_nw/ont_nw edlib NW/distance insertion sequences 50 bp-10 kb (configs 3/4)
    hifi_hw        edlib HW/locations on the same sequences

Metric: alignment GCUPS = sum(len(query) * len(target)) / time, forward matrix counted once.
  value : device-timed (CUDA events on the launching stream), inputs resident in HBM
  e2e   : the same metric through the host C-ABI call with pinned host buffers, H2D and D2H inside the timed region.
          Smith-Waterman workloads: db200_sw_batch_host_packed - the batch pre-encoded once, outside the timed region, in
          the engine's 4-bit pool format (the reference's C entry points take pre-encoded int8 codes as well, ssw.h:72-125);
          `e2e_ascii` is the same step through db200_sw_batch_host with raw ASCII CSR buffers (round 1's e2e), and
          `python_front_door` a 65,536-pair batch through dysgu_b200.batch.sw_align_batch(list, list).
"""
from __future__ import annotations

import argparse
import ctypes as C
import json
import os
import subprocess
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

METRIC = "alignment GCUPS (device-timed)"
UNIT = "GCUPS"
# ALU-pipe lane-instructions of the packed Smith-Waterman cell update: 6 s16x2 DPX instructions per 2 cells
# (U, H', H'-gapO, H, V, tagged running max); the profile merge is an IMAD on the FMA pipe and is not counted
# - see DESIGN.md "roofline"
SW_LANE_OPS_PER_CELL = 3.0
# measured on this pool's B200s by dysgu_b200/int_peak (profiles/r01_int_peak.json): VIADDMNMX.S16x2 lane-ops/s
INT_PEAK_FALLBACK = 1.81e13
# register-only Myers / Hyyro word steps per second (int_peak "myers_word_step", profiles/r02_int_peak.json): the roof of the
# edit-distance kernels - every word step of the kernels issues the same ~23 ALU-pipe instructions the benchmark issues
MYERS_PEAK_FALLBACK = 7.73e11


def load_measured_peaks():
    out = {"hbm_gbs": 6650.0, "source": "fallback"}
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        try:
            d = json.load(open(p))
            out["hbm_gbs"] = float(d.get("hbm_gbs", out["hbm_gbs"]))
            out["source"] = "measured"
        except Exception:
            pass
    out["int_lane_ops"] = INT_PEAK_FALLBACK
    out["myers_word_steps"] = MYERS_PEAK_FALLBACK
    out["int_source"] = "fallback constants"
    for name in ("r02_int_peak.json", "r01_int_peak.json"):
        ip = os.path.join(ROOT, "profiles", name)
        if os.path.exists(ip):
            try:
                d = json.load(open(ip))
                out["int_lane_ops"] = float(d["lane_ops_per_s"]["viaddmnmx_s16x2_relu"])
                out["myers_word_steps"] = float(d["lane_ops_per_s"].get("myers_word_step", MYERS_PEAK_FALLBACK))
                out["int_source"] = "measured: dysgu_b200/int_peak on this pool (profiles/%s)" % name
                break
            except Exception:
                pass
    return out


def measured_forward_traffic(pairs):
    """DRAM bytes of the forward scoring launches of one step, from this round's ncu pass over this very command
    (scripts/forward_traffic.sh -> profiles/r02_forward_traffic.json).  None unless the pass matches the batch size."""
    p = os.path.join(ROOT, "profiles", "r02_forward_traffic.json")
    try:
        d = json.load(open(p))
        if int(d["pairs_per_step"]) == int(pairs):
            return float(d["dram_bytes_per_step"]), "ncu dram__bytes_read.sum + dram__bytes_write.sum summed over the forward sw_pass_kernel launches of one step of this command (profiles/r02_forward_traffic.json)"
    except Exception:
        pass
    return None, "no ncu pass for this batch size (scripts/forward_traffic.sh)"


class ClockSampler:
    """nvidia-smi clocks / throttle reasons during the timed region."""
    Q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, index: int):
        self.index = index
        self.proc = None
        self.lines = []

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", "-i", str(self.index), "--query-gpu=" + self.Q, "--format=csv,noheader,nounits",
                                          "-lms", "200"], stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.t = threading.Thread(target=self._read, daemon=True)
            self.t.start()
        except Exception:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.lines.append(line.strip())

    def stop(self):
        if not self.proc:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        time.sleep(0.25)
        self.proc.terminate()
        try:
            self.proc.wait(timeout=2)
        except Exception:
            self.proc.kill()
        sm, mx, reasons = [], [], set()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for ln in self.lines:
            f = [x.strip() for x in ln.split(",")]
            if len(f) < 9:
                continue
            try:
                sm.append(float(f[1])); mx.append(float(f[2]))
            except ValueError:
                continue
            for name, v in zip(names, f[5:9]):
                if v.lower().startswith("active"):
                    reasons.add(name)
        return {"sm_mhz": float(np.median(sm)) if sm else None, "sm_max_mhz": max(mx) if mx else None,
                "reasons": sorted(reasons), "samples": len(sm)}


def bind_to_gpu_numa_node(index: int):
    """Pin this process to the CPUs of the NUMA node its GPU hangs off (before any pinned allocation), so that the
    page-locked staging buffers of the end-to-end arm are local to the GPU's PCIe root.  Returns a short note."""
    try:
        bus = subprocess.run(["nvidia-smi", "-i", str(index), "--query-gpu=pci.bus_id", "--format=csv,noheader"],
                             stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True, timeout=20).stdout.strip().lower()
        if bus.startswith("00000000:"):
            bus = bus[4:]
        node = int(open("/sys/bus/pci/devices/%s/numa_node" % bus).read().strip())
        if node < 0:
            return "numa: single node"
        cpus = set()
        for part in open("/sys/devices/system/node/node%d/cpulist" % node).read().strip().split(","):
            a, _, b = part.partition("-")
            cpus.update(range(int(a), int(b or a) + 1))
        allowed = cpus & os.sched_getaffinity(0)
        if not allowed:
            return "numa: node %d has no CPU in this process's cpuset" % node
        os.sched_setaffinity(0, allowed)
        return "numa: node %d, %d cpus" % (node, len(allowed))
    except Exception as exc:
        return "numa: not bound (%s)" % str(exc)[:60]


def pinned_array(L, shape, dtype):
    n = int(np.prod(shape)) * np.dtype(dtype).itemsize
    ptr = L.db200_host_alloc(max(n, 1))
    if not ptr:
        raise RuntimeError("pinned allocation failed")
    buf = (C.c_uint8 * max(n, 1)).from_address(ptr)
    return np.frombuffer(buf, dtype=dtype, count=int(np.prod(shape))).reshape(shape)


# Myers block step = 22 64-bit operations per 64-cell word step = 44 int32-equivalent lane-ops
ED_LANE_OPS_PER_WORD_STEP = 44.0

# name -> (kind, generator, default pairs per step per GPU, CPU reference sample, description)
WORKLOADS = {
    "remap_windows": ("sw", lambda W, n, s: W.remap_windows(n, seed=s), 1 << 19, 1 << 19,
                      "remap_windows: SSW(window 1000/1500/2000 bp; 2,-8,6,1)(soft clip 15-300 bp), score+coords+cigar"),
    "pe_contigs": ("sw", lambda W, n, s: W.pe_contig_pairs(n, seed=s), 1 << 18, 100000,
                   "pe_contigs: check_contig_match SSW (2,-3,10,1), paired-end contigs 100-500 bp, score+coords+cigar"),
    "long_contigs": ("sw", lambda W, n, s: W.contig_pairs(n, seed=s, platform="ont"), 1 << 14, 1600,
                     "long_contigs: check_contig_match SSW (2,-3,10,1), long-read contigs 1-10 kb (ONT error), score+coords+cigar"),
    "remap_chain": ("ed", lambda W, n, s: W.remap_chain(n, seed=s), 1 << 19, 200000,
                    "remap_chain: edlib HW/locations, clip 15-300 bp vs 1000 bp window"),
    "hifi_nw": ("ed", lambda W, n, s: W.insertion_linking(n, seed=s, platform="hifi"), 200000, 20000,
                "hifi_nw: edlib NW/distance k=-1, insertion sequences 50 bp-10 kb, HiFi error, 20 % unrelated"),
    "ont_nw": ("ed", lambda W, n, s: W.insertion_linking(n, seed=s, platform="ont"), 200000, 20000,
               "ont_nw: edlib NW/distance k=-1, insertion sequences 50 bp-10 kb, ONT r10 error, 20 % unrelated"),
    "hifi_hw": ("ed", lambda W, n, s: W.insertion_linking(n, seed=s, platform="hifi")[:2] + (dict(mode="HW", task="locations"),), 20000, 2000,
                "hifi_hw: edlib HW/locations, insertion sequences 50 bp-10 kb, HiFi error, 20 % unrelated"),
    # task "path" (no dysgu caller requests it): the locations run plus the alignment of the first location
    "remap_path": ("ed", lambda W, n, s: W.remap_chain(n, seed=s)[:2] + (dict(mode="HW", task="path"),), 1 << 18, 100000,
                   "remap_path: edlib HW/path, clip 15-300 bp vs 1000 bp window, distance+locations+edit operations"),
    "hifi_path": ("ed", lambda W, n, s: W.insertion_linking(n, seed=s, platform="hifi")[:2] + (dict(mode="NW", task="path"),), 20000, 4000,
                  "hifi_path: edlib NW/path, insertion sequences 50 bp-10 kb, HiFi error, 20 % unrelated, distance+edit operations"),
}


def make_workload(pairs: int, seed: int, name: str = "remap_windows"):
    from dysgu_b200 import workloads
    q, t, prm = WORKLOADS[name][1](workloads, pairs, seed)
    cells = float((q.lengths.astype(np.float64) * t.lengths.astype(np.float64)).sum())
    return q, t, prm, cells


def cpu_reference(O, name, q, t, prm, cores):
    """The reference's compiled cores (oracle/_ref) over a CSR batch on `cores` host threads -> result with .seconds / .fixed"""
    if WORKLOADS[name][0] == "sw":
        return O.ref_sw((q.buf, q.off), (t.buf, t.off), prm["match"], prm["mismatch"], prm["gap_open"], prm["gap_extend"],
                        nthreads=cores, csr=True)
    return O.ref_ed((q.buf, q.off), (t.buf, t.off), prm["mode"], prm["task"], nthreads=cores, csr=True)


# ---------------------------------------------------------------------------------------------------
def run_reference(args):
    """The reference's own CPU implementation (oracle/_ref, compiled unmodified from the reference checkout)
    on all host threads, same workload definition, bounded sample per step."""
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return 0
    try:
        from oracle import oracle as O
        O.driver_lib()
    except Exception as exc:  # the oracle always exists in this tier; report loudly if the prebuilt files are missing
        print(json.dumps({"impl": "reference", "unavailable": "oracle/_ref not loadable: %s" % str(exc)[:120]}))
        return 0
    cores = os.cpu_count() or 1
    kind = WORKLOADS[args.workload][0]
    sample = args.ref_pairs or WORKLOADS[args.workload][3]
    q, t, prm, cells = make_workload(sample, 20261021, args.workload)

    def step():
        return cpu_reference(O, args.workload, q, t, prm, cores).seconds

    for _ in range(args.warmup):
        step()
    t0 = time.perf_counter()
    inner = 0.0
    for _ in range(args.steps):
        inner += step()
    wall = time.perf_counter() - t0
    value = cells * args.steps / inner / 1e9
    line = {
        "impl": "reference", "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": args.gpus, "steps": args.steps,
        "warmup": args.warmup, "ms_per_step": 1e3 * inner / args.steps, "higher_is_better": True, "scaling": "weak",
        "vs_baseline": None, "dtype": "int16" if kind == "sw" else "u64", "data": "synthetic",
        "config": {"workload": WORKLOADS[args.workload][4], "pairs_per_step": sample, "cells_per_step": cells},
        "cpu_baseline": {"value": value, "unit": UNIT, "cores": cores, "kind": "reference",
                         "sample": "%d pairs per step of the same workload, %s per pair at C level" % (
                             sample, "ssw_init+ssw_align (ssw.c -O2)" if kind == "sw" else "edlibAlign (edlib.cpp -O3)")},
        "e2e": {"value": value, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "wall_s": wall,
    }
    print(json.dumps(line))
    return 0



def merge_shape_figure(L, B, torch, dev, dist, local_rank, rank, world, steps=3, pairs=8192):
    """BASELINE configs 4/5 beside the headline, so that the driver's 1/2/4/8-GPU sweep covers the merge shape as well: long-read
    contig pairs 1-10 kb (ONT error) through check_contig_match's scoring (2,-3,10,1), score + coordinates + cigar, pairs sharded
    by rank.  Device-timed GCUPS (inputs resident) and end to end through db200_sw_batch_host_packed (one synchronous call per
    step).  Max over ranks, whole-job aggregate."""
    from dysgu_b200 import _lib
    q, t, prm, cells = make_workload(pairs, 20261121 + rank, "long_contigs")
    n = len(q)
    prm_c = _lib.SwParams()
    prm_c.match, prm_c.mismatch, prm_c.gap_open, prm_c.gap_extend = prm["match"], prm["mismatch"], prm["gap_open"], prm["gap_extend"]
    prm_c.score_size, prm_c.flag = 2, 1
    qb, tb = int(q.off[-1]), int(t.off[-1])
    cap = qb + tb + 2 * n + 16
    d_q, d_t = torch.from_numpy(q.buf).to(dev), torch.from_numpy(t.buf).to(dev)
    d_qo, d_to = torch.from_numpy(q.off).to(dev), torch.from_numpy(t.off).to(dev)
    d_res = torch.zeros((n, 8), dtype=torch.int32, device=dev)
    d_cig = torch.zeros(cap, dtype=torch.int32, device=dev)
    d_co = torch.zeros(n + 1, dtype=torch.int64, device=dev)
    d_st = torch.zeros(n, dtype=torch.int32, device=dev)
    stream = torch.cuda.current_stream(dev)

    def dstep():
        _lib.check(L.db200_sw_batch_device(local_rank, C.byref(prm_c), d_q.data_ptr(), d_qo.data_ptr(), qb, d_t.data_ptr(), d_to.data_ptr(), tb, n,
                                           int(q.lengths.max()), int(t.lengths.max()), None, d_res.data_ptr(), d_cig.data_ptr(), cap,
                                           d_co.data_ptr(), d_st.data_ptr(), None, C.c_void_p(stream.cuda_stream)), "db200_sw_batch_device")

    def barrier():
        torch.cuda.synchronize(dev)
        if dist is not None:
            dist.barrier()
        torch.cuda.synchronize(dev)

    for _ in range(2):
        dstep()
    barrier()
    if int((d_st != 0).sum().item()):
        raise SystemExit("bench.py: merge shape reported a non-zero status")
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record(stream)
    for _ in range(steps):
        dstep()
    e1.record(stream)
    barrier()
    ms = torch.tensor([e0.elapsed_time(e1)], dtype=torch.float64, device=dev)
    tc = torch.tensor([cells], dtype=torch.float64, device=dev)
    if dist is not None:
        dist.all_reduce(ms, op=dist.ReduceOp.MAX)
        dist.all_reduce(tc, op=dist.ReduceOp.SUM)
    pq, pt = B.pack_seqs(q, pinned=True), B.pack_seqs(t, pinned=True)
    hres = pinned_array(L, (n, 8), np.int32); hcig = pinned_array(L, (cap,), np.uint32)
    hco = pinned_array(L, (n + 1,), np.int64); hst = pinned_array(L, (n,), np.int32)
    ql, tl = np.ascontiguousarray(pq.lengths), np.ascontiguousarray(pt.lengths)

    def hstep():
        _lib.check(L.db200_sw_batch_host_packed(local_rank, C.byref(prm_c), pq.units.ctypes.data, pq.unit_off.ctypes.data, ql.ctypes.data,
                                                pt.units.ctypes.data, pt.unit_off.ctypes.data, tl.ctypes.data, n, None, hres.ctypes.data,
                                                hcig.ctypes.data, cap, hco.ctypes.data, hst.ctypes.data), "db200_sw_batch_host_packed")
    hstep()
    barrier()
    t0 = time.perf_counter()
    for _ in range(steps):
        hstep()
    dt = torch.tensor([time.perf_counter() - t0], dtype=torch.float64, device=dev)
    if dist is not None:
        dist.all_reduce(dt, op=dist.ReduceOp.MAX)
    if not np.array_equal(hres[:, :7], d_res.cpu().numpy()[:, :7]):
        raise SystemExit("bench.py: merge shape: host and device arms disagree")
    return {"workload": WORKLOADS["long_contigs"][4], "pairs_per_step_per_gpu": n, "value": float(tc.item()) * steps / (float(ms.item()) * 1e-3) / 1e9,
            "unit": UNIT, "ms_per_step": float(ms.item()) / steps,
            "e2e": {"value": float(tc.item()) * steps / float(dt.item()) / 1e9, "unit": UNIT, "ms_per_step": 1e3 * float(dt.item()) / steps,
                    "h2d_bytes_per_step": pq.nbytes + pt.nbytes + 8 * n, "d2h_bytes_per_step": int(32 * n + 12 * n + 4 * int(hco[n]))}}


# ---------------------------------------------------------------------------------------------------
def run_ours(args):
    import torch
    from dysgu_b200 import _lib
    from dysgu_b200 import batch as B

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    if not torch.cuda.is_available():
        raise SystemExit("bench.py: no CUDA device - the alignment engine has no CPU fallback")
    numa_note = bind_to_gpu_numa_node(local_rank) if world > 1 else "numa: not bound (single GPU)"
    torch.cuda.set_device(local_rank)
    dev = torch.device("cuda", local_rank)
    dist = None
    if world > 1:
        import torch.distributed as dist_mod
        dist = dist_mod
        dist.init_process_group(backend="nccl", device_id=dev)
    L = _lib.lib()
    _lib.check(L.db200_init(local_rank), "db200_init")
    name = args.workload
    kind = WORKLOADS[name][0]

    # ---- per-rank shard of the workload (pairs are independent: weak scaling, no collective on the data path)
    pairs = args.pairs or WORKLOADS[name][2]
    q, t, prm, cells = make_workload(pairs, 20261021 + rank, name)
    n = len(q)
    q_bytes, t_bytes = int(q.off[-1]), int(t.off[-1])
    max_q = int(q.lengths.max())
    max_t = int(t.lengths.max())

    if kind == "sw":
        prm_c = _lib.SwParams()
        prm_c.match, prm_c.mismatch, prm_c.gap_open, prm_c.gap_extend = prm["match"], prm["mismatch"], prm["gap_open"], prm["gap_extend"]
        prm_c.score_size, prm_c.flag = 2, 1
        var_cap = int(4 * n + 1024) if name == "remap_windows" else int(q_bytes + t_bytes + 2 * n + 16)   # cigar words
        var_dtype, var_np, res_cols = torch.int32, np.uint32, 8
    else:
        prm_c = _lib.EdParams()
        prm_c.mode, prm_c.task, prm_c.k = B.ED_MODES[prm["mode"]], B.ED_TASKS[prm["task"]], -1
        var_cap = int(t_bytes + 2 * n + 16) if prm["mode"] != "NW" else int(n + 16)                                       # locations
        var_dtype, var_np, res_cols = torch.int32, np.int32, 4
    var_words = var_cap * (1 if kind == "sw" else 2)
    want_path = kind == "ed" and prm["task"] == "path"
    path_cap = q_bytes + t_bytes + 16

    # ---- host-side batch preparation for the Smith-Waterman workloads, once, outside every timed region: pairs ordered by target
    #      length class and both sides encoded in the 4-bit pool format (db200_sw_pack_host)
    order = pq = pt = None
    if kind == "sw":
        order = np.argsort(-((t.lengths + 31) // 32), kind="stable") if os.environ.get("DB200_BENCH_NO_BUCKETING") is None else np.arange(n)
        pq, pt = B.pack_seqs(q.take(order), pinned=True), B.pack_seqs(t.take(order), pinned=True)
        d_units = torch.from_numpy(np.concatenate([pq.units[:int(pq.unit_off[-1]) * 4], pt.units[:int(pt.unit_off[-1]) * 4]]).view(np.int32)).to(dev)
        d_qlen = torch.from_numpy(np.ascontiguousarray(pq.lengths)).to(dev)
        d_tlen = torch.from_numpy(np.ascontiguousarray(pt.lengths)).to(dev)
        n_units = int(pq.unit_off[-1] + pt.unit_off[-1])

    # ---- device-resident arm --------------------------------------------------------------------------
    d_q = torch.from_numpy(q.buf).to(dev)
    d_t = torch.from_numpy(t.buf).to(dev)
    d_qoff = torch.from_numpy(q.off).to(dev)
    d_toff = torch.from_numpy(t.off).to(dev)
    d_res = torch.zeros((n, res_cols), dtype=torch.int32, device=dev)
    d_var = torch.zeros(var_words, dtype=var_dtype, device=dev)
    d_voff = torch.zeros(n + 1, dtype=torch.int64, device=dev)
    d_status = torch.zeros(n, dtype=torch.int32, device=dev)
    d_path = torch.zeros(path_cap if want_path else 1, dtype=torch.uint8, device=dev)
    d_poff = torch.zeros(n + 1 if want_path else 1, dtype=torch.int64, device=dev)
    stream = torch.cuda.current_stream(dev)

    def device_step():
        if kind == "sw":   # the batch resident in HBM in pool format (north_star: "device-resident packed batches")
            rc = L.db200_sw_batch_device_packed(local_rank, C.byref(prm_c), d_units.data_ptr(), n_units, d_qlen.data_ptr(), d_tlen.data_ptr(),
                                                q_bytes, t_bytes, n, max_q, max_t, None, d_res.data_ptr(), d_var.data_ptr(), var_cap,
                                                d_voff.data_ptr(), d_status.data_ptr(), None, C.c_void_p(stream.cuda_stream))
        elif want_path:
            rc = L.db200_ed_path_batch_device(local_rank, C.byref(prm_c), d_q.data_ptr(), d_qoff.data_ptr(), q_bytes, d_t.data_ptr(),
                                              d_toff.data_ptr(), t_bytes, n, max_q, max_t, None, d_res.data_ptr(), d_var.data_ptr(), var_cap,
                                              d_voff.data_ptr(), None, d_path.data_ptr(), path_cap, d_poff.data_ptr(), None,
                                              C.c_void_p(stream.cuda_stream))
        else:
            rc = L.db200_ed_batch_device(local_rank, C.byref(prm_c), d_q.data_ptr(), d_qoff.data_ptr(), q_bytes, d_t.data_ptr(),
                                         d_toff.data_ptr(), t_bytes, n, max_q, max_t, None, d_res.data_ptr(), d_var.data_ptr(), var_cap,
                                         d_voff.data_ptr(), None, C.c_void_p(stream.cuda_stream))
        _lib.check(rc, "db200_%s_batch_device" % kind)

    def barrier():
        torch.cuda.synchronize(dev)
        if dist is not None:
            dist.barrier()
        torch.cuda.synchronize(dev)

    for _ in range(max(args.warmup, 3)):
        device_step()
    barrier()
    bad = int((d_status != 0).sum().item()) if kind == "sw" else int((d_res[:, 0] != 0).sum().item())
    if bad:
        raise SystemExit("bench.py: %d pairs reported a non-zero status" % bad)
    if want_path and int(d_poff[-1].item()) <= 0:
        raise SystemExit("bench.py: no alignment paths")
    checksum = int(d_res[:, 0 if kind == "sw" else 1].sum().item())

    L.db200_profile_enable(local_rank, 1)
    L.db200_kernel_launches(local_rank, 1)
    L.db200_ed_word_steps(local_rank, 1)
    sampler = ClockSampler(local_rank)
    if rank == 0:
        sampler.start()
    barrier()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record(stream)
    for _ in range(args.steps):
        device_step()
    e1.record(stream)
    barrier()
    ms = e0.elapsed_time(e1)
    clocks = sampler.stop() if rank == 0 else None
    launches = int(L.db200_kernel_launches(local_rank, 1))
    word_steps = int(L.db200_ed_word_steps(local_rank, 1)) / args.steps
    stage_ms = (C.c_float * 16)()
    L.db200_profile_collect(local_rank, stage_ms, 16)
    L.db200_profile_enable(local_rank, 0)
    stage_ms = [float(x) for x in stage_ms]
    tms = torch.tensor([ms], dtype=torch.float64, device=dev)
    if dist is not None:
        dist.all_reduce(tms, op=dist.ReduceOp.MAX)
    ms_max = float(tms.item())
    total_cells = cells * world  # every rank processes a shard of the same size and shape distribution
    if dist is not None:
        tc = torch.tensor([cells], dtype=torch.float64, device=dev)
        dist.all_reduce(tc, op=dist.ReduceOp.SUM)
        total_cells = float(tc.item())
    value = total_cells * args.steps / (ms_max * 1e-3) / 1e9

    # ---- end-to-end arm: host C-ABI call, pinned host buffers, H2D + D2H inside the timed region --------------
    hq = pinned_array(L, (q_bytes,), np.uint8); hq[:] = q.buf
    ht = pinned_array(L, (t_bytes,), np.uint8); ht[:] = t.buf
    hqo = pinned_array(L, (n + 1,), np.int64); hqo[:] = q.off
    hto = pinned_array(L, (n + 1,), np.int64); hto[:] = t.off
    hres = pinned_array(L, (n, res_cols), np.int32)
    hvar = pinned_array(L, (var_words,), var_np)
    hvoff = pinned_array(L, (n + 1,), np.int64)
    hstat = pinned_array(L, (n,), np.int32)
    hpath = pinned_array(L, (path_cap if want_path else 1,), np.uint8); hpath[:] = 0
    hpoff = pinned_array(L, (n + 1 if want_path else 1,), np.int64)

    def host_step():
        if kind == "sw":
            rc = L.db200_sw_batch_host(local_rank, C.byref(prm_c), hq.ctypes.data, hqo.ctypes.data, ht.ctypes.data, hto.ctypes.data, n,
                                       None, hres.ctypes.data, hvar.ctypes.data, var_cap, hvoff.ctypes.data, hstat.ctypes.data)
        elif want_path:
            rc = L.db200_ed_path_batch_host(local_rank, C.byref(prm_c), hq.ctypes.data, hqo.ctypes.data, ht.ctypes.data, hto.ctypes.data, n,
                                            None, hres.ctypes.data, hvar.ctypes.data, var_cap, hvoff.ctypes.data, hpath.ctypes.data,
                                            path_cap, hpoff.ctypes.data)
        else:
            rc = L.db200_ed_batch_host(local_rank, C.byref(prm_c), hq.ctypes.data, hqo.ctypes.data, ht.ctypes.data, hto.ctypes.data, n,
                                       None, hres.ctypes.data, hvar.ctypes.data, var_cap, hvoff.ctypes.data)
        _lib.check(rc, "db200_%s_batch_host" % kind)

    e2e_steps = max(1, min(args.steps, args.e2e_steps))
    for _ in range(2):
        host_step()
    barrier()
    t0 = time.perf_counter()
    for _ in range(e2e_steps):
        host_step()
    torch.cuda.synchronize(dev)
    dt = time.perf_counter() - t0
    tdt = torch.tensor([dt], dtype=torch.float64, device=dev)
    if dist is not None:
        dist.all_reduce(tdt, op=dist.ReduceOp.MAX)
    e2e_value = total_cells * e2e_steps / float(tdt.item()) / 1e9
    if int(hres[:, 0 if kind == "sw" else 1].astype(np.int64).sum()) != checksum:
        raise SystemExit("bench.py: host and device arms disagree")
    nvar = int(hvoff[n])
    h2d = q_bytes + t_bytes + 2 * 8 * (n + 1)
    d2h = 4 * res_cols * n + 8 * (n + 1) + (4 * n + 4 * nvar if kind == "sw" else 8 * nvar) + (int(hpoff[n]) + 8 * (n + 1) if want_path else 0)

    # ---- end-to-end arm with the batch pre-encoded on the host (db200_sw_batch_host_packed): the engine's 4-bit pool format,
    #      written once by db200_sw_pack_host outside the timed region - the counterpart of the int8 codes the reference's
    #      C entry points take (ssw.h:72-125; the wrapper encodes before ssw_init / ssw_align) at half a byte per base
    e2e_packed = e2e_packed_sync = None
    if kind == "sw":
        # host-side batch preparation, once, outside the timed region ("queues pairs into packed batches with length bucketing",
        # BASELINE.json north_star): pairs ordered by target length class, both sides encoded.  With the pairs of a length bucket
        # next to each other every chunk of the host pipeline feeds one or two bucket kernels with as many pairs as the whole batch
        # would - chunks in submission order spread every bucket thin over all chunks.  Results come back in this order.
        # (order, pq, pt: prepared before the device arm, see above)
        hql = pinned_array(L, (n,), np.int32); hql[:] = pq.lengths
        htl = pinned_array(L, (n,), np.int32); htl[:] = pt.lengths
        hquo = pinned_array(L, (n + 1,), np.int64); hquo[:] = pq.unit_off
        htuo = pinned_array(L, (n + 1,), np.int64); htuo[:] = pt.unit_off
        # three engine instances on the device, one host thread each: consecutive steps overlap (H2D of step k + 1 and D2H of
        # step k - 1 under the kernels of step k; the kernels of different calls run one after the other, engine.cu), the way a
        # long job - config 2's ~10 M pairs - feeds its batches.  Every step still copies its own inputs in and its own results
        # out inside the timed region; each thread has its own output buffers.  (Two threads fall into lock step now and then:
        # 23.7 / 30.3 ms per step in two runs; three: 22.6.)
        NT = int(os.environ.get("DB200_BENCH_STREAM_THREADS", "3"))
        outs = [(pinned_array(L, (n, res_cols), np.int32), pinned_array(L, (var_words,), var_np), pinned_array(L, (n + 1,), np.int64),
                 pinned_array(L, (n,), np.int32)) for _ in range(NT)]

        def packed_step(j=0):
            r_, v_, o_, s_ = outs[j]
            rc = L.db200_sw_batch_host_packed(local_rank | (j << 8), C.byref(prm_c), pq.units.ctypes.data, hquo.ctypes.data, hql.ctypes.data,
                                              pt.units.ctypes.data, htuo.ctypes.data, htl.ctypes.data, n, None, r_.ctypes.data,
                                              v_.ctypes.data, var_cap, o_.ctypes.data, s_.ctypes.data)
            _lib.check(rc, "db200_sw_batch_host_packed")

        def timed(fn_per_thread, nthreads, steps):
            errs = []

            def run(j):
                try:
                    for _ in range(j, steps, nthreads):
                        fn_per_thread(j)
                except Exception as exc:   # noqa: BLE001
                    errs.append(exc)
            barrier()
            t0 = time.perf_counter()
            if nthreads == 1:
                run(0)
            else:
                ths = [threading.Thread(target=run, args=(j,)) for j in range(nthreads)]
                for th in ths:
                    th.start()
                for th in ths:
                    th.join()
            torch.cuda.synchronize(dev)
            dt_ = time.perf_counter() - t0
            if errs:
                raise errs[0]
            td = torch.tensor([dt_], dtype=torch.float64, device=dev)
            if dist is not None:
                dist.all_reduce(td, op=dist.ReduceOp.MAX)
            return float(td.item())

        for j in range(NT):
            packed_step(j); packed_step(j)
        sync_s = timed(packed_step, 1, e2e_steps)                        # one synchronous call after the other
        stream_steps = NT * e2e_steps
        for j in range(NT):   # the overlap happens between the calls: each call runs its kernels once over the whole batch
            L.db200_set_host_chunk_pairs(local_rank | (j << 8), C.c_int64(int(os.environ.get("DB200_BENCH_STREAM_CHUNK", n))))
        for j in range(NT):
            packed_step(j)
        stream_s = timed(packed_step, NT, stream_steps)                  # two host threads alternate over the steps
        for j in range(NT):
            L.db200_set_host_chunk_pairs(local_rank | (j << 8), C.c_int64(0))
        for j in range(NT):
            if not np.array_equal(outs[j][0], hres[order]) or int(outs[j][2][n]) != nvar:
                raise SystemExit("bench.py: packed and ASCII host arms disagree")
        pk_bytes = pq.nbytes + pt.nbytes + 2 * 4 * n
        note = ("pairs grouped by target length class and pre-encoded in the 4-bit pool format in page-locked memory (numpy argsort + "
                "db200_sw_pack_host, outside the timed region - the reference's C entry points take pre-encoded int8 codes as well), "
                "db200_sw_batch_host_packed")
        e2e_packed_sync = {"value": total_cells * e2e_steps / sync_s / 1e9, "unit": UNIT, "h2d_bytes_per_step": pk_bytes, "d2h_bytes_per_step": d2h,
                           "steps": e2e_steps, "ms_per_step": 1e3 * sync_s / e2e_steps, "input": note, "mode": "one synchronous call per step"}
        e2e_packed = {"value": total_cells * stream_steps / stream_s / 1e9, "unit": UNIT, "h2d_bytes_per_step": pk_bytes, "d2h_bytes_per_step": d2h,
                      "steps": stream_steps, "ms_per_step": 1e3 * stream_s / stream_steps, "input": note,
                      "mode": "%d host threads, each with its own engine instance on the GPU (DB200_INSTANCE) and its own output buffers, alternate "
                              "over the steps: the copies of one step run under the kernels of another; one chunk per call "
                              "(db200_set_host_chunk_pairs)" % NT}

    # ---- the Python front door on a smaller batch: list of bytes in, numpy results out (dysgu_b200.batch.sw_align_batch)
    front_door = None
    if kind == "sw" and rank == 0:
        nf = min(n, 1 << 16)
        ql, tl_ = q.take(np.arange(nf)).tolist(), t.take(np.arange(nf)).tolist()
        fcells = float((q.lengths[:nf].astype(np.float64) * t.lengths[:nf].astype(np.float64)).sum())
        B.sw_align_batch(ql, tl_, device=local_rank, **prm)
        t0 = time.perf_counter()
        rr = B.sw_align_batch(ql, tl_, device=local_rank, **prm)
        dtf = time.perf_counter() - t0
        if not np.array_equal(rr.fixed[:, :7], hres[:nf, :7]):
            raise SystemExit("bench.py: Python front door and C ABI disagree")
        front_door = {"value": fcells / dtf / 1e9, "unit": UNIT, "pairs": nf, "ms": 1e3 * dtf,
                      "call": "dysgu_b200.batch.sw_align_batch(list[bytes], list[bytes]) -> numpy (pageable buffers, CSR built in Python)"}

    merge_shape = None
    if name == "remap_windows" and not args.no_merge_shape:
        merge_shape = merge_shape_figure(L, B, torch, dev, dist, local_rank, rank, world)

    if rank != 0:
        if dist is not None:
            dist.destroy_process_group()
        return 0

    # ---- roofline of the dominant kernel -----------------------------------------------------------------------
    peaks = load_measured_peaks()
    peak = peaks["int_lane_ops"] / 1e12
    stage_names = ["begin", "pack", "bucket", "forward", "forward_rerun", "reverse_prep", "reverse", "reverse_rerun", "finalize",
                   "traceback", "cigar", "ed_prep", "ed_forward", "ed_starts", "ed_output", "ed_path"]
    stages = {k: v / args.steps for k, v in zip(stage_names, stage_ms[:16]) if v > 0 or k == "begin"}
    if kind == "sw":
        k_ms = stage_ms[3] / args.steps          # forward sw_pass_kernel launches of one step (this rank)
        lane_ops = cells * SW_LANE_OPS_PER_CELL
        achieved = lane_ops / (k_ms * 1e-3) / 1e12 if k_ms > 0 else None
        alg_bytes = (q_bytes + t_bytes) / 2.0 + 16.0 * n   # packed 4-bit codes in, one PassOut record out
        roofline = {
            "bound": "int_alu", "kernel": "sw_pass_kernel (forward scoring pass, all length buckets)",
            "achieved": achieved, "peak": peak, "unit": "Tlaneop/s (s16x2 DPX lane-instructions)",
            "frac": (achieved / peak) if achieved else None, "peak_source": peaks["int_source"],
            "frac_whole_step": lane_ops / (ms_max / args.steps * 1e-3) / 1e12 / peak,
            "ops_per_cell": SW_LANE_OPS_PER_CELL, "kernel_ms_per_step": k_ms, "kernel_gcups": cells / (k_ms * 1e-3) / 1e9 if k_ms > 0 else None,
            "hbm_algorithmic_gbs": alg_bytes / (k_ms * 1e-3) / 1e9 if k_ms > 0 else None, "hbm_peak_gbs": peaks["hbm_gbs"],
            "hbm_peak_source": peaks["source"],
            "traffic": measured_forward_traffic(n)[0] if name == "remap_windows" else None, "algorithmic_bytes": alg_bytes,
            "traffic_source": measured_forward_traffic(n)[1],
            "stage_ms_per_step": stages,
        }
        roofline["hbm"] = {"bound": "hbm", "achieved": roofline["hbm_algorithmic_gbs"], "peak": peaks["hbm_gbs"], "unit": "GB/s",
                           "frac": (roofline["hbm_algorithmic_gbs"] / peaks["hbm_gbs"]) if roofline["hbm_algorithmic_gbs"] else None,
                           "note": "integer dynamic programming: the kernel is bound by the ALU pipe (frac above), not by HBM"}
    else:
        k_ms = stage_ms[12] / args.steps         # banded rounds + ed_myers_kernel forward launches of one step
        # word steps actually evaluated per second against the measured register-only Myers word-step rate (int_peak
        # "myers_word_step"): both sides count the same unit, so the fraction is what the ALU pipe could still give.
        peak_ws = peaks["myers_word_steps"] / 1e12
        achieved = word_steps / (k_ms * 1e-3) / 1e12 if k_ms > 0 else None
        alg_bytes = float(q_bytes + t_bytes) + 16.0 * n
        roofline = {
            "bound": "int_alu", "kernel": "ed_banded_nw_kernel + ed_myers_kernel (forward pass, all configurations)" if prm["mode"] == "NW"
            else "ed_myers_kernel (forward pass, all configurations)",
            "achieved": achieved, "peak": peak_ws, "unit": "T word steps/s (one Myers / Hyyro update of a 64-row word, edlib.cpp:411-443)",
            "frac": (achieved / peak_ws) if achieved else None, "peak_source": peaks["int_source"] + " myers_word_step",
            "frac_algorithmic_44_ops": (word_steps * ED_LANE_OPS_PER_WORD_STEP / (k_ms * 1e-3) / 1e12 / peak) if k_ms > 0 else None,
            "ops_per_word_step": ED_LANE_OPS_PER_WORD_STEP, "word_steps_per_step": word_steps,
            "word_steps_full_matrix": float(((q.lengths + 63) // 64 * t.lengths).sum()),
            "kernel_ms_per_step": k_ms, "kernel_gcups": cells / (k_ms * 1e-3) / 1e9 if k_ms > 0 else None,
            "hbm_algorithmic_gbs": alg_bytes / (k_ms * 1e-3) / 1e9 if k_ms > 0 else None, "hbm_peak_gbs": peaks["hbm_gbs"],
            "hbm_peak_source": peaks["source"], "traffic": None, "algorithmic_bytes": alg_bytes,
            "stage_ms_per_step": stages,
        }
        roofline["hbm"] = {"bound": "hbm", "achieved": roofline["hbm_algorithmic_gbs"], "peak": peaks["hbm_gbs"], "unit": "GB/s",
                           "frac": (roofline["hbm_algorithmic_gbs"] / peaks["hbm_gbs"]) if roofline["hbm_algorithmic_gbs"] else None,
                           "note": "bit-parallel dynamic programming: the kernels are bound by the ALU pipe (frac above), not by HBM"}

    # ---- CPU baseline (rank 0, N = 1 only): the reference cores on all host threads, bounded sample ------------
    cpu_baseline = None
    if world == 1 and not args.no_cpu_baseline:
        try:
            from oracle import oracle as O
            cores = os.cpu_count() or 1
            sample = min(n, args.ref_pairs or WORKLOADS[name][3])
            qs, ts = q.take(np.arange(sample)), t.take(np.arange(sample))
            scell = float((qs.lengths.astype(np.float64) * ts.lengths.astype(np.float64)).sum())
            cpu_reference(O, name, qs, ts, prm, cores)  # warm
            r = cpu_reference(O, name, qs, ts, prm, cores)
            cpu_baseline = {"value": scell / r.seconds / 1e9, "unit": UNIT, "cores": cores, "kind": "reference",
                            "sample": "first %d pairs of the step's batch, %s per pair at C level, %.1f s" % (
                                sample, "ssw_init+ssw_align (ssw.c -O2)" if kind == "sw" else "edlibAlign (edlib.cpp -O3)", r.seconds)}
            # spot parity on the sample against the reference itself
            ncmp = 7 if kind == "sw" else 4
            mism = int((r.fixed[:, :ncmp] != hres[:sample, :ncmp]).any(axis=1).sum())
            cpu_baseline["parity_mismatches_on_sample"] = mism
            if kind == "sw":   # cigars: same CSR offsets and the same BAM-encoded operations
                same_off = bool(np.array_equal(np.asarray(r.cigar_off[:sample + 1]), hvoff[:sample + 1]))
                nops = int(r.cigar_off[sample])
                cpu_baseline["cigar_ops_on_sample"] = nops
                cpu_baseline["cigar_mismatches_on_sample"] = int((np.asarray(r.cigar[:nops]) != hvar[:nops]).sum()) if same_off else \
                    int((np.diff(np.asarray(r.cigar_off[:sample + 1])) != np.diff(hvoff[:sample + 1])).sum())
            if want_path:   # the reference's alignment arrays (one slot per pair) against ours (back to back), operation by operation
                ln = r.path_len.astype(np.int64)
                idx = np.repeat(np.asarray(r.path_off, dtype=np.int64) - (np.cumsum(ln) - ln), ln) + np.arange(int(ln.sum()), dtype=np.int64)
                same_len = bool(np.array_equal(np.diff(hpoff[:sample + 1]), ln))
                cpu_baseline["path_mismatches_on_sample"] = int((r.path[idx] != hpath[:int(ln.sum())]).sum()) if same_len else int((np.diff(hpoff[:sample + 1]) != ln).sum())
                cpu_baseline["path_operations_on_sample"] = int(ln.sum())
        except Exception as exc:
            cpu_baseline = {"value": None, "unit": UNIT, "cores": 0, "kind": "reference", "sample": "unavailable: %s" % str(exc)[:100]}

    e2e_ascii = {"value": e2e_value, "unit": UNIT, "h2d_bytes_per_step": h2d, "d2h_bytes_per_step": d2h, "steps": e2e_steps,
                 "ms_per_step": 1e3 * float(tdt.item()) / e2e_steps,
                 "input": "ASCII CSR batches in page-locked memory, db200_%s_batch_host" % kind}
    line = {
        "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps, "warmup": max(args.warmup, 3),
        "ms_per_step": ms_max / args.steps, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
        "dtype": "int16" if kind == "sw" else "u64", "data": "synthetic",
        "config": {"workload": WORKLOADS[name][4],
                   "pairs_per_step_per_gpu": n, "cells_per_step_per_gpu": cells, "input_bytes_per_step_per_gpu": q_bytes + t_bytes,
                   "l2": "inputs larger than L2 (%.0f MB per step%s)" % (
                       ((pq.nbytes + pt.nbytes) if kind == "sw" else (q_bytes + t_bytes)) / 1e6, ", 4-bit pool format" if kind == "sw" else ""),
                   "device_arm_input": "length-bucketed batch resident in HBM in the 4-bit pool format (db200_sw_batch_device_packed)" if kind == "sw"
                   else "ASCII CSR batch resident in HBM", "sharding": "pairs by rank, no collective",
                   "host": numa_note},
        "clocks": clocks,
        "e2e": e2e_packed if e2e_packed else e2e_ascii,
        "e2e_single_call": e2e_packed_sync,
        "e2e_ascii": e2e_ascii,
        "python_front_door": front_door,
        "merge_shape": merge_shape,
        "gpu_launches": launches,
        "roofline": roofline,
        "cpu_baseline": cpu_baseline,
        "checksum_%s" % ("score1" if kind == "sw" else "edit_distance"): checksum,
    }
    print(json.dumps(line))
    if dist is not None:
        dist.destroy_process_group()
    return 0


def main():
    # keep stdout clean for the single JSON line: libraries (NCCL prints its version) write to fd 1 as well
    real_stdout = os.dup(1)
    os.dup2(2, 1)
    sys.stdout = os.fdopen(real_stdout, "w")
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=10)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--workload", default="remap_windows", choices=sorted(WORKLOADS), help="default: the headline (BASELINE configs[1])")
    ap.add_argument("--pairs", type=int, default=0, help="pairs per step per GPU (0: the workload's default)")
    ap.add_argument("--ref-pairs", type=int, default=0, help="pairs per step of the CPU reference sample (0: the workload's default)")
    ap.add_argument("--e2e-steps", type=int, default=10, help="steps per host thread of the end-to-end arms (the streaming arm runs 3 threads)")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-merge-shape", action="store_true", help="skip the long-contig (configs 4/5) figure that rides on the default line")
    args = ap.parse_args()
    if args.impl == "reference":
        return run_reference(args)
    return run_ours(args)


if __name__ == "__main__":
    sys.exit(main())
