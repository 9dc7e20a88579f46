#!/usr/bin/env python
"""bench.py — the intersect stage of mcDispatch (BVH build + BVH x BVH traversal + exact edge/face narrowphase)
on the BASELINE.json workloads, one process per GPU.

    python bench.py --gpus N --steps K --warmup W                 # this repo's CUDA path, BASELINE configs[1] (C2)
    python bench.py --workload c5|c3|c4|c5dense|c3batch|c4batch    # the other configs
    python bench.py --impl reference --gpus N --steps K ...        # the reference's own CPU path (oracle/_ref)

A "step" is one whole intersect stage of one mcDispatch.  One JSON line is printed by rank 0:

  value        candidate face pairs ("tri-pairs") pushed through build + traversal + narrowphase per second, inputs
               resident in HBM, device time from CUDA events, max over ranks; ms_per_step = intersect-stage ms per
               dispatch (the other half of BASELINE.json's metric)
  e2e          the same through the C-ABI with HOST buffers: every step uploads both meshes from pinned memory and reads
               pairs, registry records and status back
  roofline     the dominant kernel's algorithmic bytes / its CUDA-event duration vs the measured HBM peak
  cpu_baseline the reference's own CPU implementation of the same stage on this box's host cores

N > 1 ("weak"): every rank runs its own dispatch of the same workload (the MultipleContextsInParallel pattern, one
context per GPU, no data-path collective); value = all ranks' pairs / max-over-ranks time.  The sharded single-dispatch
mode (leaf-range split + NCCL exchange inside the C-ABI, mcb200_intersect_stage_sharded) is measured in the same run — on
the workload itself and on C5, where the stage is long enough for a split to pay — and reported under "sharded".

Batch workloads (c3batch: 256 planar sections of one terrain; c4batch: many small CSG pairs): a step is still one
dispatch; they are issued through mcb200_batch_intersect_host over several context lanes of the GPU.
"""
from __future__ import annotations

import argparse
import ctypes
import json
import os
import re
import subprocess
import sys
import tempfile
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

from mcut_b200 import meshgen  # noqa: E402

METRIC = "intersect_stage_tri_pairs_per_s"
UNIT = "candidate face pairs/s (BVH build + traversal + exact narrowphase per mcDispatch)"
DBL = meshgen.MC_DISPATCH_VERTEX_ARRAY_DOUBLE | meshgen.MC_DISPATCH_ENFORCE_GENERAL_POSITION


def load_peaks():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        with open(p) as fp:
            return float(json.load(fp)["hbm_gbs"]), "measured (MEASURED_PEAKS.json)"
    return 6650.0, "fallback (B200_PROFILING.md)"


WORKLOADS = {
    "c2": "C2: CSG of two cube-spheres, 1,002,252 triangles / 501,128 vertices each, R=20",
    "c2small": "C2-small: two cube-spheres, 101,568 triangles each (smoke size)",
    "c5": "C5: dense overlap of two cube-spheres, 2,007,372 triangles each, with near-coplanar regions (1,037,662 tests need "
          "the exact orient3d stages)",
    "c5dense": "C5-dense: SURVEY's original recipe (dense overlap, no exact tests), 2,007,372 triangles each",
    "c3": "C3 (one of its 256 dispatches): 3,998,792-triangle terrain cut by the section supertriangle of plane 0",
    "c4": "C4 (one of its dispatches): two icospheres of 5,120 triangles",
    "c3batch": "C3: planar sectioning of a 3,998,792-triangle terrain by 256 planes (mcEnqueueDispatchPlanarSection's supertriangle)",
    "c4batch": "C4: independent small CSG pairs (two icospheres of 5,120 triangles each), MultipleContextsInParallel pattern",
}


def workload(name: str):
    if name == "c2":
        return meshgen.c2_two_spheres(k=289)
    if name == "c2small":
        return meshgen.c2_two_spheres(k=92)
    if name == "c5":
        return meshgen.c5_coplanar_regions(k=409)
    if name == "c5dense":
        return meshgen.c5_near_coplanar(k=409)
    if name == "c3":
        ter = meshgen.terrain()
        nrm, _ = meshgen.c3_plane(0)
        return ter, meshgen.c3_supertriangle(ter[0], nrm), DBL
    if name == "c4":
        return meshgen.c4_pair(0)
    raise SystemExit(f"unknown workload {name}")


# The reference's own stage is QUADRATIC in the number of faces one face overlaps: on C3's section (one supertriangle over the
# whole terrain) it takes 94 s at 250,632 triangles and 372 s at 500,000 (measured, 8 cores), i.e. hours at the workload's
# 3,998,792.  The CPU legs of C3 therefore run a bounded sample — the same terrain recipe at 63,368 triangles — and say so.
C3_CPU_SAMPLE_N = 179
C3_CPU_SAMPLE_NOTE = ("BOUNDED SAMPLE: the same terrain recipe at 63,368 triangles (n = 179), because the reference's stage is quadratic "
                      "here (94 s at 250,632 triangles, 372 s at 500,000: hours at the workload's 3,998,792); ")


def cpu_workload(name: str):
    """what the CPU legs (cpu_baseline, --impl reference) run: the workload itself, or a bounded sample of it + a note"""
    if name == "c3":
        ter = meshgen.terrain(n=C3_CPU_SAMPLE_N)
        nrm, _ = meshgen.c3_plane(0)
        return (ter, meshgen.c3_supertriangle(ter[0], nrm), DBL), C3_CPU_SAMPLE_NOTE
    return workload(name), ""


def base_config(name: str, world: int):
    """the keys both arms (ours and --impl reference) print"""
    return {"workload": WORKLOADS[name], "workload_id": name,
            "parallelism": f"{world} independent dispatch stream(s)" if world > 1 else "1 dispatch stream"}


# ----------------------------------------------------------------------------------------------------------------------
# clocks sampler (nvidia-smi during the timed region)
# ----------------------------------------------------------------------------------------------------------------------
class ClockSampler:
    Q = ("clocks.sm,clocks.max.sm,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
         "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, index: int):
        self.index = index
        self.samples = []
        self.proc = None

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", f"--id={self.index}", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits",
                                          "-lms", "100"], stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.thread = threading.Thread(target=self._read, daemon=True)
            self.thread.start()
        except Exception:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.samples.append(line.strip())

    def stop(self):
        if not self.proc:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        time.sleep(0.15)
        self.proc.terminate()
        try:
            self.proc.wait(timeout=2)
        except Exception:
            self.proc.kill()
        sm, mx, reasons = [], [], set()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for s in self.samples:
            parts = [p.strip() for p in s.split(",")]
            if len(parts) < 6:
                continue
            try:
                sm.append(float(parts[0]))
                mx.append(float(parts[1]))
            except ValueError:
                continue
            for n, v in zip(names, parts[2:6]):
                if v.lower().startswith("active"):
                    reasons.add(n)
        return {"sm_mhz": float(np.median(sm)) if sm else None, "sm_max_mhz": max(mx) if mx else None,
                "reasons": sorted(reasons), "samples": len(sm)}


# ----------------------------------------------------------------------------------------------------------------------
# reference arm / cpu baseline: the reference's own intersect stage, timed by its own stage timers
# ----------------------------------------------------------------------------------------------------------------------
REF_STAGES = ["build_oibvh", "intersectOIBVHs", "Prepare edge-to-face pairs", "Build edge bounding boxes",
              "Cull redundant edge-face pairs", "Compute intersecting face properties",
              "Calculate intersection points (edge-to-face)"]


def have_reference_binary() -> bool:
    return os.path.exists(os.path.join(ROOT, "oracle", "_ref", "stage_harness_prof"))


def write_input(src, cut, flags, path):
    from mcut_b200.mcbio import write_mcb
    d = {"src_xyz": src[0], "src_faces": src[1], "cut_xyz": cut[0], "cut_faces": cut[1], "flags": np.array([flags], dtype=np.uint32)}
    if src[2] is not None:
        d["src_sizes"] = src[2]
    if cut[2] is not None:
        d["cut_sizes"] = cut[2]
    write_mcb(path, d)


def reference_stage_start(in_path: str, out_path: str, helpers: int, extra=()):
    exe = os.path.join(ROOT, "oracle", "_ref", "stage_harness_prof")
    return subprocess.Popen([exe, in_path, out_path, "--helpers", str(helpers), "--no-events", "--abort-after-narrowphase", *extra],
                            stdout=subprocess.DEVNULL, stderr=subprocess.PIPE, text=True, cwd=os.path.dirname(out_path))


def reference_stage_finish(proc, out_path: str):
    """(stage ms = sum of the reference's own timers, candidate pairs, intersection points) of one finished harness run"""
    from mcut_b200.mcbio import read_mcb
    _, err = proc.communicate()
    ms = 0.0
    seen = {}
    for line in err.splitlines():
        m = re.search(r'\[MCUT\]\[PROF:\d+\]: "(.*)" \((\d+)ms\)', line)
        if m and m.group(1) in REF_STAGES:
            ms += float(m.group(2))
            seen[m.group(1)] = seen.get(m.group(1), 0.0) + float(m.group(2))
    o = read_mcb(out_path)
    # sub-millisecond precision where the harness measured the call itself (build x2, traversal)
    fine = o["timings_ms"]
    fine_bt = float(fine[fine[:, 0] < 2, 1].sum())
    coarse_bt = seen.get("build_oibvh", 0.0) + seen.get("intersectOIBVHs", 0.0)
    ms = ms - coarse_bt + fine_bt
    n_pairs = int(o["isect0_map_entries"].shape[0] // 2) if "isect0_map_entries" in o else 0
    n_points = int(o["dispatch0_ipoints"].shape[0]) if "dispatch0_ipoints" in o else 0
    return ms, n_pairs, n_points


def reference_streams(inputs, streams: int, cores: int, td: str, extra_of=None):
    """Runs the harness on `inputs` (paths), `streams` processes at a time, each with cores/streams - 1 helper threads.
    Returns (wall seconds, [stage ms], total pairs, helper threads per process)."""
    helpers = max(cores // max(streams, 1) - 1, 0)
    t0 = time.perf_counter()
    stage_ms, pairs = [], 0
    pending = list(enumerate(inputs))
    running = []
    while pending or running:
        while pending and len(running) < streams:
            i, ip = pending.pop(0)
            op = os.path.join(td, f"out{i}.mcb")
            running.append((reference_stage_start(ip, op, helpers, extra_of(i) if extra_of else ()), op))
        proc, op = running.pop(0)
        ms, npairs, _ = reference_stage_finish(proc, op)
        stage_ms.append(ms)
        pairs += npairs
    return time.perf_counter() - t0, stage_ms, pairs, helpers


def oracle_port_stage_once(src, cut, flags):
    from oracle import pyoracle
    t0 = time.perf_counter()
    r = pyoracle.intersect_stage(src, cut, flags)
    return (time.perf_counter() - t0) * 1e3, len(r["pairs"]), len(r["records"])


def c3_plane_extra(k):
    nrm, off = meshgen.c3_plane(k)
    return ["--planar", repr(float(nrm[0])), repr(float(nrm[1])), repr(float(nrm[2])), repr(float(off))]


def run_reference_arm(args):
    """The reference's own CPU path on the same workload.  At --gpus N it runs N dispatch streams side by side on the host
    (N harness processes, cores/N - 1 helper threads each) so that the N-GPU figure is compared like for like."""
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    cores = os.cpu_count() or 1
    streams = max(args.gpus, 1)
    name = args.workload
    cfg = base_config(name, streams)
    with tempfile.TemporaryDirectory() as td:
        kind = "reference" if have_reference_binary() else "port"
        if name in ("c3batch", "c4batch"):
            if kind != "reference":
                print(json.dumps({"impl": "reference", "unavailable": "batch workloads need oracle/_ref (the reference harness)"}))
                return
            total = args.steps
            inputs, extra_of = [], None
            if name == "c4batch":
                for j in range(args.warmup + total):
                    s, c, f = meshgen.c4_pair(j)
                    ip = os.path.join(td, f"in{j}.mcb")
                    write_input(s, c, f, ip)
                    inputs.append(ip)
                streams_eff = min(cores, max(streams * 8, 8))  # the tutorial pattern: many contexts in flight, no helpers
            else:
                ter = meshgen.terrain(n=C3_CPU_SAMPLE_N)  # (see C3_CPU_SAMPLE_NOTE)
                dummy = (np.zeros((3, 3)), np.array([0, 1, 2], dtype=np.uint32), None)
                ip = os.path.join(td, "in.mcb")
                write_input(ter, dummy, DBL, ip)
                inputs = [ip] * (args.warmup + total)
                extra_of = c3_plane_extra
                streams_eff = streams
            if args.warmup:
                reference_streams(inputs[:args.warmup], streams_eff, cores, td, extra_of)
            wall, stage_ms, pairs, helpers = reference_streams(inputs[args.warmup:], streams_eff, cores, td,
                                                               (lambda i: extra_of(i + args.warmup)) if extra_of else None)
            ms_per_step = wall * 1e3 / total
            value = pairs / wall
            sample = (C3_CPU_SAMPLE_NOTE if name == "c3batch" else "") + (f"{total} dispatches, {streams_eff} harness processes in flight with {helpers} helper threads each; value = pairs / "
                      "wall time of the batch (each process = one mcDispatch of the unmodified reference cut short after its narrowphase)")
            cfg.update({"dispatches": total, "pairs_per_dispatch": pairs / max(total, 1)})
        else:
            (src, cut, flags), sample_note = cpu_workload(name)
            times, n_pairs = [], 0
            if kind == "reference":
                write_input(src, cut, flags, os.path.join(td, "in.mcb"))
            for i in range(args.warmup + args.steps):
                if kind == "reference":
                    wall, stage_ms, pairs, helpers = reference_streams([os.path.join(td, "in.mcb")] * streams, streams, cores, td)
                    ms, n_pairs = max(stage_ms), pairs // streams
                else:
                    ms, n_pairs, _ = oracle_port_stage_once(src, cut, flags)
                if i >= args.warmup:
                    times.append(ms)
            ms_per_step = float(np.mean(times))
            value = streams * n_pairs / (ms_per_step * 1e-3) if kind == "reference" else n_pairs / (ms_per_step * 1e-3)
            sample = sample_note + (f"{args.steps} whole intersect stages of the {'sample' if sample_note else 'full workload'} on {streams} stream(s); each = one mcDispatch of the unmodified "
                      "reference cut short after its narrowphase, stage ms = sum of the reference's own timers (build_oibvh x2, "
                      "intersectOIBVHs, the five kernel.cpp:1781-3231 stages), max over the streams") if kind == "reference" else \
                f"{args.steps} runs of the oracle port (single thread)"
            cfg.update({"pairs_per_dispatch": n_pairs})
    line = {
        "impl": "reference", "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": args.gpus, "steps": args.steps,
        "warmup": args.warmup, "ms_per_step": ms_per_step, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
        "dtype": "f64", "data": "synthetic", "config": cfg,
        "cpu_baseline": {"value": value, "unit": UNIT, "cores": cores if kind == "reference" else 1, "kind": kind, "sample": sample},
        "e2e": {"value": value, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
    }
    print(json.dumps(line))


# ----------------------------------------------------------------------------------------------------------------------
# this repo's arm
# ----------------------------------------------------------------------------------------------------------------------
def pinned_copy(torch, a: np.ndarray):
    t = torch.from_numpy(np.ascontiguousarray(a)).pin_memory()
    return t, t.numpy()


def profile_facts(workload_id: str):
    """per-kernel facts from the committed ncu captures of THIS workload (profiles/r02_<workload>_kernels.json, written by
    tools/ncu_to_profiles.py): DRAM bytes per launch, FP64-pipe utilisation; {} when the workload was not captured"""
    p = os.path.join(ROOT, "profiles", f"r02_{workload_id}_kernels.json")
    try:
        return json.load(open(p))
    except Exception:
        return {}


def algorithmic_bytes(kname: str, src_nv, src_nf, cut_nv, cut_nf, counts, vbytes=24):
    """Algorithmic bytes per LAUNCH of a kernel (SURVEY.md §8-d figures).  Build kernels run once per mesh, so
    their per-launch figure uses the mean mesh size."""
    F = (src_nf + cut_nf) / 2.0
    V = (src_nv + cut_nv) / 2.0
    n_pairs, n_tests = counts["n_pairs"], counts["n_tests"]
    table = {
        "k_vertex_bbox": vbytes * V,
        # coords gathered once per vertex + 12 B of indices per face in; 48 B box + code + sort key out
        "k_face_codes<true>": vbytes * V + 12 * F + 48 * F + 8 * F,
        "k_face_codes<false>": vbytes * V + 12 * F + 48 * F + 8 * F,
        "k_face_bbox<true>": vbytes * V + 12 * F + 48 * F,
        "k_morton": 48 * F + 8 * F,
        "onesweep_pass_u32_kv": 16 * F,  # one radix pass: key + value read once, written once
        # sorted keys + face ids + exact boxes (gathered) in; exact boxes in leaf order, single-precision leaf boxes, the levels
        # above (1/32 + 1/1024 ...), groups (about F/16 of 40 B) out
        "k_leaves": 4 * F + 4 * F + 48 * F + 48 * F + 24 * F * (1 + 1 / 32 + 1 / 1024) + 40 * F / 16,
        "k_traverse": 24.0 * counts["n_node_tests"] + 8.0 * n_pairs,
        "k_pair_scatter": 16.0 * n_pairs,
        "k_pair_segsort": 16.0 * n_pairs + 8.0 * src_nf,
        # triangle narrowphase: per pair the pair word, six vertex ids, six vertices, two candidate flags and
        # 8 B per queued pair; per queued pair the same again + two face boxes, six edge ids, six edge rows, about three owner
        # boxes; per queued test the entry, the pair, five vertex ids + vertices, the edge id, 32 B per record
        "k_tri_prefilter": (8 + 24 + 6 * vbytes + 2) * n_pairs + 8.0 * counts.get("n_mid", 0),
        "k_tri_classify": (8 + 8 + 24 + 6 * vbytes + 96 + 24 + 48 + 144) * counts.get("n_mid", 0) + 8.0 * (counts.get("n_open", 0) + counts.get("n_cross", 0)),
        "k_tri_resolve": (8 + 8 + 20 + 5 * vbytes + 4) * (counts.get("n_open", 0) + counts.get("n_cross", 0)) + 32.0 * counts.get("n_records", 0),
        "k_tests_filter_poly": 8.0 * n_pairs + 128.0 * n_tests,
    }
    return table.get(kname)


def run_ours(args):
    import torch
    import torch.distributed as dist

    from mcut_b200 import stage
    from mcut_b200._lib import BatchItem, Counts, HostMesh, HostSoup

    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    # libraries (NCCL's version banner, ...) may write to fd 1; the contract is ONE JSON line on stdout, so everything
    # else goes to stderr until the line is printed
    real_stdout = os.dup(1)
    os.dup2(2, 1)
    if world > 1:
        dist.init_process_group("nccl", device_id=torch.device(f"cuda:{local_rank}"))
    torch.cuda.set_device(local_rank)
    dev = torch.device(f"cuda:{local_rank}")
    # torch's legacy default stream has handle 0, which the C-ABI reads as "make your own stream"; use an explicit
    # stream for everything so torch.cuda.Event (which sees torch's CURRENT stream only) brackets our kernels
    tstream = torch.cuda.Stream(device=dev)
    torch.cuda.set_stream(tstream)
    ctx = stage.Context(local_rank, tstream.cuda_stream)
    L = ctx.L
    vp = ctypes.c_void_p
    flush_buf = torch.empty(256 * 1024 * 1024, dtype=torch.uint8, device=dev)  # > 126 MB L2
    peak, peak_src = load_peaks()

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def max_over_ranks(x: float) -> float:
        if world > 1:
            t = torch.tensor([x], dtype=torch.float64, device=dev)
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
            return float(t.item())
        return x

    def sum_over_ranks(x: float) -> float:
        if world > 1:
            t = torch.tensor([x], dtype=torch.float64, device=dev)
            dist.all_reduce(t, op=dist.ReduceOp.SUM)
            return float(t.item())
        return x

    def timed_loop(fn, steps, warmup, flush=True):
        for _ in range(warmup):
            fn()
        torch.cuda.synchronize()
        evs = []
        barrier()
        for _ in range(steps):
            if flush:
                flush_buf.zero_()  # evict the previous step's lines from L2 (untimed)
            a = torch.cuda.Event(enable_timing=True)
            b = torch.cuda.Event(enable_timing=True)
            a.record()
            fn()
            b.record()
            evs.append((a, b))
        barrier()
        return max_over_ranks(sum(a.elapsed_time(b) for a, b in evs))

    def settle_capacity(step, result):
        """One untimed run: a dispatch whose pair count exceeds the default buffer reports MCB200_ERR_CAPACITY from
        mcb200_result_counts, which also raises the capacity; the timed loops then run with buffers that fit."""
        for _ in range(5):
            try:
                step()
                result.counts()
                return
            except stage.Mcb200Error as e:
                if e.code != stage.ERR_CAPACITY:
                    raise
        raise RuntimeError("pair buffer did not settle")

    def make_comm():
        idbuf = ctypes.create_string_buffer(128)
        if rank == 0:
            ctx.check(L.mcb200_comm_unique_id(idbuf))
        t = torch.frombuffer(bytearray(idbuf.raw), dtype=torch.uint8).to(dev)
        dist.broadcast(t, 0)
        raw = bytes(t.cpu().numpy().tobytes())
        h = vp()
        ctx.check(L.mcb200_comm_create(ctx.h, world, rank, raw, ctypes.byref(h)))
        return h

    class Resident:
        """both meshes, their frame and the polygon soup resident on the device (the `value` leg's inputs)"""

        def __init__(self, src, cut):
            sx, sf, _ = src
            cx, cf, _ = cut
            self.src_nv, self.src_nf = meshgen.mesh_counts(src)
            self.cut_nv, self.cut_nf = meshgen.mesh_counts(cut)
            self.com, self.shift, sbb, cbb = stage.vertex_parameters(sx, cx)
            self.eps = stage.cut_bbox_eps(cbb, 1e-4, False)
            soff = np.arange(0, sf.size + 1, 3, dtype=np.uint32)
            coff = np.arange(0, cf.size + 1, 3, dtype=np.uint32)
            fv, fe, ev, ef = stage.soup_ids(self.src_nv, soff, sf, coff, cf)
            self.ne, self.nh = ev.shape[0], fv.size
            self.keep, self.host = [], {}
            for nm, arr in (("sx", sx), ("sf", sf), ("cx", cx), ("cf", cf), ("fv", fv), ("fe", fe), ("ef", ef)):
                t, a = pinned_copy(torch, arr)
                self.keep.append(t)
                self.host[nm] = a
            h = self.host
            self.m_src, self.m_cut, self.soup = vp(), vp(), vp()
            ctx.check(L.mcb200_mesh_create(ctx.h, 0, h["sx"].ctypes.data, self.src_nv, h["sf"].ctypes.data_as(stage.c_u32p), None, self.src_nf,
                                           ctypes.byref(self.m_src)))
            ctx.check(L.mcb200_mesh_create(ctx.h, 0, h["cx"].ctypes.data, self.cut_nv, h["cf"].ctypes.data_as(stage.c_u32p), None, self.cut_nf,
                                           ctypes.byref(self.m_cut)))
            for m in (self.m_src, self.m_cut):
                ctx.check(L.mcb200_mesh_set_frame(ctx.h, m, self.com.ctypes.data_as(stage.c_dp), self.shift.ctypes.data_as(stage.c_dp), None))
            ctx.check(L.mcb200_soup_create(ctx.h, self.src_nf, self.cut_nf, self.nh, self.ne, h["fv"].ctypes.data_as(stage.c_u32p),
                                           h["fe"].ctypes.data_as(stage.c_u32p), h["ef"].ctypes.data_as(stage.c_u32p), ctypes.byref(self.soup)))
            self.hm_src = HostMesh(0, h["sx"].ctypes.data, self.src_nv, h["sf"].ctypes.data, None, self.src_nf)
            self.hm_cut = HostMesh(0, h["cx"].ctypes.data, self.cut_nv, h["cf"].ctypes.data, None, self.cut_nf)
            self.h_soup = HostSoup(self.nh, self.ne, h["fe"].ctypes.data, h["ef"].ctypes.data)

        def stage(self, res):
            ctx.check(L.mcb200_intersect_stage(ctx.h, self.m_src, self.m_cut, self.eps, self.soup, res.h, 0))

        def stage_host(self, res, soup_arg=None):
            ctx.check(L.mcb200_intersect_stage_host(ctx.h, ctypes.byref(self.hm_src), ctypes.byref(self.hm_cut),
                                                    self.com.ctypes.data_as(stage.c_dp), self.shift.ctypes.data_as(stage.c_dp), None, self.eps,
                                                    soup_arg, res.h, 0))

        def free(self):
            L.mcb200_soup_free(ctx.h, self.soup)
            L.mcb200_mesh_free(ctx.h, self.m_src)
            L.mcb200_mesh_free(ctx.h, self.m_cut)

    def measure_sharded(R, comm, steps, warmup, reference_res):
        """one dispatch split over all ranks; returns the section for the JSON line.  The merged pairs and records of every
        rank are compared byte for byte with that rank's own single-GPU result."""
        res3 = stage.Result(ctx)

        def step():
            for _ in range(5):
                rc = L.mcb200_intersect_stage_sharded(ctx.h, comm, R.m_src, R.m_cut, R.eps, R.soup, res3.h, 0)
                if rc != stage.ERR_CAPACITY:
                    ctx.check(rc)
                    return
            raise RuntimeError("sharded stage: capacities did not settle")

        step()
        res3.counts()
        total_ms = timed_loop(step, steps, warmup)
        c3 = res3.counts()
        same = None
        if reference_res is not None:
            c1 = reference_res.counts()
            same = bool(c1.n_pairs == c3.n_pairs and c1.n_records == c3.n_records and c1.n_tests == c3.n_tests
                        and reference_res.pairs().tobytes() == res3.pairs().tobytes()
                        and reference_res.records().tobytes() == res3.records().tobytes())
            same = bool(sum_over_ranks(1.0 if same else 0.0) == world)
        out = {"ms_per_dispatch": total_ms / steps, "pairs": int(c3.n_pairs), "records": int(c3.n_records),
               "pairs_per_s": int(c3.n_pairs) / (total_ms / steps * 1e-3),
               "identical_to_single_gpu_result_on_every_rank": same,
               "scheme": "replicated meshes + BVHs, 4096-leaf chunks of the Morton order dealt round-robin, exchange inside the C-ABI: "
                         "ncclAllGather of the counter blocks, grouped ncclBroadcast of pairs and records, ncclAllReduce of per-face "
                         "pair counts and candidate flags, canonical orders on every rank"}
        res3.free()
        return out

    name = args.workload
    cfg = base_config(name, world)

    # ==================================================================================================================
    # batch workloads
    # ==================================================================================================================
    if name in ("c3batch", "c4batch"):
        nlanes = args.lanes if args.lanes > 0 else (8 if name == "c4batch" else 2)
        n_items = args.batch if args.batch > 0 else (256 if name == "c3batch" else 2000)
        my_items = list(range(rank, n_items, world))  # round-robin over the GPUs
        lanes = [ctx] + [stage.Context(local_rank) for _ in range(nlanes - 1)]
        lane_res = [stage.Result(c) for c in lanes]
        ctx_arr = (vp * nlanes)(*[c.h for c in lanes])
        res_arr = (vp * nlanes)(*[r.h for r in lane_res])
        keep = []
        items = (BatchItem * len(my_items))()
        h2d = 0
        if name == "c4batch":
            for k, j in enumerate(my_items):
                s, c, f = meshgen.c4_pair(j)
                ts, sx = pinned_copy(torch, s[0])
                tc, cx = pinned_copy(torch, c[0])
                if k == 0:
                    tf, fa = pinned_copy(torch, s[1])
                    keep.append(tf)
                keep += [ts, tc]
                nv, nf = sx.shape[0], fa.size // 3
                items[k].src = HostMesh(0, sx.ctypes.data, nv, fa.ctypes.data, None, nf)
                items[k].cut = HostMesh(0, cx.ctypes.data, nv, fa.ctypes.data, None, nf)
                items[k].com = None
                items[k].gp_constant = 1e-4
                items[k].flags = 0
                h2d = sx.nbytes + cx.nbytes + 2 * fa.nbytes
            src_faces = cut_faces = nf
            cut_nv = nv
        else:
            ter = meshgen.terrain()
            tt, tx = pinned_copy(torch, ter[0])
            tf, tfa = pinned_copy(torch, ter[1])
            keep += [tt, tf]
            nv, nf = tx.shape[0], tfa.size // 3
            for k, j in enumerate(my_items):
                nrm, _ = meshgen.c3_plane(j)
                tri = meshgen.c3_supertriangle(ter[0], nrm)
                tc, cx = pinned_copy(torch, tri[0])
                tcf, cfa = pinned_copy(torch, tri[1])
                keep += [tc, tcf]
                items[k].src = HostMesh(0, tx.ctypes.data, nv, tfa.ctypes.data, None, nf)
                items[k].cut = HostMesh(0, cx.ctypes.data, 3, cfa.ctypes.data, None, 1)
                items[k].com = None
                items[k].gp_constant = 1e-4
                # the terrain arrays are the same in every dispatch: after a lane's first item only the plane travels
                items[k].flags = stage.STAGE_SRC_RESIDENT if k >= nlanes else 0
            src_faces, cut_faces, cut_nv = nf, 1, 3
            h2d = 3 * 24 + 12  # per dispatch once the terrain is resident on the lane (first item of a lane: + 96 MB)
        counts = (Counts * len(my_items))()

        def run_batch(lo, hi):
            n = hi - lo
            if n <= 0:
                return
            sub = ctypes.cast(ctypes.addressof(items) + lo * ctypes.sizeof(BatchItem), ctypes.POINTER(BatchItem))
            csub = ctypes.cast(ctypes.addressof(counts) + lo * ctypes.sizeof(Counts), ctypes.POINTER(Counts))
            rc = L.mcb200_batch_intersect_host(ctx_arr, res_arr, nlanes, sub, n, csub)
            if rc:
                ctx.check(rc)

        warm = min(max(args.warmup * nlanes, 2 * nlanes), len(my_items) // 2)
        warm -= warm % nlanes  # the timed part starts on lane 0 again
        run_batch(0, warm)  # allocations, graph capture, (c3batch) the terrain lands on every lane
        torch.cuda.synchronize()
        sampler = ClockSampler(local_rank)
        if rank == 0:
            sampler.start()
        barrier()
        launches0 = sum(c.launches for c in lanes)
        t0 = time.perf_counter()
        run_batch(warm, len(my_items))
        torch.cuda.synchronize()
        wall = time.perf_counter() - t0
        launches = sum(c.launches for c in lanes) - launches0
        barrier()
        clocks = sampler.stop() if rank == 0 else None
        timed = len(my_items) - warm
        pairs = sum(int(counts[i].n_pairs) for i in range(warm, len(my_items)))
        bad = sum(1 for i in range(len(my_items)) if counts[i].status not in (0, 1))
        wall = max_over_ranks(wall)
        pairs_all = sum_over_ranks(float(pairs))
        timed_all = sum_over_ranks(float(timed))
        e2e_value = pairs_all / wall
        # per-kernel device times of a few dispatches (lane 0, profiling on: no graph) for the roofline entry
        ctx.set_profiling(True)
        prof_n = min(4, len(my_items))
        sub_ctx = (vp * 1)(ctx.h)
        sub_res = (vp * 1)(lane_res[0].h)
        c_tmp = (Counts * prof_n)()
        for it in range(prof_n):
            items[it].flags = 0
        L.mcb200_batch_intersect_host(sub_ctx, sub_res, 1, items, prof_n, c_tmp)
        prof = ctx.profile_read()
        ctx.set_profiling(False)
        kern = {k: {"launches_per_step": cnt / prof_n, "ms_per_step": ms / prof_n, "ms_per_launch": ms / cnt} for k, (cnt, ms) in prof.items()}
        top = max(kern, key=lambda k: kern[k]["ms_per_step"])
        counts1 = {"n_pairs": int(c_tmp[0].n_pairs), "n_tests": int(c_tmp[0].n_tests), "n_node_tests": int(c_tmp[0].n_node_tests)}
        roofline, rooflines = None, []
        facts = profile_facts(name)
        for kname in sorted(kern, key=lambda k: -kern[k]["ms_per_step"]):
            ab = algorithmic_bytes(kname, nv, src_faces, cut_nv, cut_faces, counts1)
            if ab is None:
                continue
            ach = ab / (kern[kname]["ms_per_launch"] * 1e-3) / 1e9
            if ach / peak > 1.2:
                continue  # the mean-size byte model does not describe this launch (e.g. a one-face mesh)
            e = {"kernel": kname, "bound": "hbm", "achieved": ach, "peak": peak, "unit": "GB/s", "frac": ach / peak,
                 "traffic": facts.get(kname, {}).get("dram_bytes_per_launch"), "peak_source": peak_src,
                 "algorithmic_bytes_per_launch": ab, "ms_per_launch": kern[kname]["ms_per_launch"]}
            rooflines.append(e)
            roofline = roofline or e
        cpu = None
        if rank == 0 and world == 1 and not args.no_cpu_baseline and have_reference_binary():
            cores = os.cpu_count() or 1
            with tempfile.TemporaryDirectory() as td:
                if name == "c4batch":
                    sample_n = min(64, n_items)
                    inputs = []
                    for j in range(sample_n):
                        s, c, f = meshgen.c4_pair(j)
                        ip = os.path.join(td, f"in{j}.mcb")
                        write_input(s, c, f, ip)
                        inputs.append(ip)
                    w, sms, prs, helpers = reference_streams(inputs, cores, cores, td)
                    what = f"{sample_n} pairs, {cores} reference processes in flight with 0 helper threads (the tutorial's pattern)"
                else:
                    sample_n = 2
                    ip = os.path.join(td, "in.mcb")
                    write_input(ter, (np.zeros((3, 3)), np.array([0, 1, 2], dtype=np.uint32), None), DBL, ip)
                    w, sms, prs, helpers = reference_streams([ip] * sample_n, 1, cores, td, c3_plane_extra)
                    what = f"{sample_n} planar sections through mcEnqueueDispatchPlanarSection, one at a time with {helpers} helper threads"
                cpu = {"value": prs / w, "unit": UNIT, "cores": cores, "kind": "reference", "dispatches_per_s": sample_n / w,
                       "stage_ms_mean": float(np.mean(sms)),
                       "sample": what + "; value = candidate pairs / wall time of the sample (process start-up and mesh conversion included: "
                                        "the reference has no batched entry point)"}
        if rank == 0:
            cfg.update({"dispatches": int(timed_all), "lanes_per_gpu": nlanes, "src_faces": src_faces, "cut_faces": cut_faces,
                        "pairs_per_dispatch": pairs_all / max(timed_all, 1), "dispatches_per_s": timed_all / wall,
                        "failed_dispatches": bad,
                        "timing": "host wall clock around mcb200_batch_intersect_host between device synchronisations (several "
                                  "streams are in flight: no single stream's events bracket the batch); L2 is not flushed between "
                                  "dispatches (every dispatch brings new inputs from the host)"})
            line = {"metric": METRIC, "value": e2e_value, "unit": UNIT, "n_gpus": world, "steps": int(timed_all), "warmup": warm,
                    "ms_per_step": wall * 1e3 / max(timed_all / world, 1), "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
                    "dtype": "f64", "data": "synthetic", "config": cfg, "clocks": clocks,
                    "e2e": {"value": e2e_value, "unit": UNIT, "h2d_bytes_per_step": int(h2d), "d2h_bytes_per_step": 56,
                            "call": "mcb200_batch_intersect_host",
                            "note": "value == e2e for batch workloads: the batch entry point takes host arrays by design; inputs come "
                                    "from pinned host memory every dispatch, the counts/status of every dispatch are read back"},
                    "gpu_launches": int(launches), "roofline": roofline, "cpu_baseline": cpu, "rooflines": rooflines, "kernels": kern,
                    "top_kernel": top}
            sys.stdout.flush()
            os.dup2(real_stdout, 1)
            print(json.dumps(line), flush=True)
            os.dup2(2, 1)
        if world > 1:
            dist.barrier()
            dist.destroy_process_group()
        return

    # ==================================================================================================================
    # single-dispatch workloads
    # ==================================================================================================================
    src, cut, flags = workload(name)
    R = Resident(src, cut)
    res = stage.Result(ctx)
    settle_capacity(lambda: R.stage(res), res)

    # ---- value: resident inputs ----
    launches0 = ctx.launches
    sampler = ClockSampler(local_rank)
    if rank == 0:
        sampler.start()
    total_ms = timed_loop(lambda: R.stage(res), args.steps, args.warmup)
    clocks = sampler.stop() if rank == 0 else None
    launches = (ctx.launches - launches0) // (args.steps + args.warmup) * args.steps
    c = res.counts()
    counts = {k: int(getattr(c, k)) for k in ("n_pairs", "n_node_tests", "n_tests", "n_exact", "n_records", "n_cand_faces")}
    status = int(c.status)
    qc = (ctypes.c_uint64 * 4)()
    ctx.check(L.mcb200_result_queue_counts(ctx.h, res.h, qc))
    counts.update({"n_mid": int(qc[0]), "n_open": int(qc[1]), "n_cross": int(qc[2]), "n_full": int(qc[3])})
    # the number of edge/face tests the REFERENCE runs on this input (the side prefilter dismisses most of them before the
    # culls; MCB200_NARROW_COUNT_TESTS = 8 puts them through the culls to be counted): one untimed run, then the plain one again
    ctx.check(L.mcb200_intersect_stage(ctx.h, R.m_src, R.m_cut, R.eps, R.soup, res.h, 8))
    counts["n_tests_reference"] = int(res.counts().n_tests)
    R.stage(res)
    assert int(res.counts().n_tests) == counts["n_tests"]
    ms_per_step = total_ms / args.steps
    value = world * counts["n_pairs"] / (ms_per_step * 1e-3)

    # ---- e2e: host buffers through the C-ABI every step ----
    pairs_host_t = torch.empty(max(counts["n_pairs"] * 2, 1 << 16), dtype=torch.int64).pin_memory()
    rec_host_t = torch.empty(max(counts["n_records"] * 2, 1 << 12) * 4, dtype=torch.float64).pin_memory()
    pairs_host, rec_host = pairs_host_t.numpy(), rec_host_t.numpy()
    res2 = stage.Result(ctx)
    d2h = {"bytes": 0}

    def step_e2e(soup_arg=None):
        # ONE reference-facing call with host arrays: uploads are pipelined with the builds inside it; soup == NULL: the
        # polygon soup is numbered on the device, only the two meshes travel
        R.stage_host(res2, soup_arg)
        cc = res2.counts()
        ctx.check(L.mcb200_result_read_pairs(ctx.h, res2.h, pairs_host.ctypes.data_as(stage.c_u64p), pairs_host.size))
        ctx.check(L.mcb200_result_read_records(ctx.h, res2.h, ctypes.cast(rec_host.ctypes.data, ctypes.POINTER(stage.Record)),
                                               rec_host.size // 4))
        d2h["bytes"] = int(cc.n_pairs) * 8 + int(cc.n_records) * 32 + 128
        d2h["pairs"], d2h["records"] = int(cc.n_pairs), int(cc.n_records)

    settle_capacity(step_e2e, res2)
    e2e_steps = max(3, min(args.steps, 10))
    e2e_ms = timed_loop(step_e2e, e2e_steps, max(args.warmup, 3), flush=False) / e2e_steps
    assert d2h["pairs"] == counts["n_pairs"] and d2h["records"] == counts["n_records"], "host-array path disagrees with the resident path"
    h2d = sum(R.host[k].nbytes for k in ("sx", "sf", "cx", "cf"))
    e2e_hs_ms = timed_loop(lambda: step_e2e(ctypes.byref(R.h_soup)), e2e_steps, 3, flush=False) / e2e_steps
    e2e = {"value": world * counts["n_pairs"] / (e2e_ms * 1e-3), "unit": UNIT, "ms_per_step": e2e_ms,
           "h2d_bytes_per_step": int(h2d), "d2h_bytes_per_step": int(d2h["bytes"]), "call": "mcb200_intersect_stage_host",
           "ms_per_step_with_host_soup_ids": e2e_hs_ms,
           "h2d_bytes_with_host_soup_ids": int(h2d + R.host["fe"].nbytes + R.host["ef"].nbytes),
           "note": "inputs = both meshes from pinned host memory (uploads pipelined with the builds on a copy stream, polygon soup "
                   "numbered on the device); outputs = sorted pairs, registry records, status"}

    # ---- per-kernel device times (separate pass, event pair around every launch) -> roofline of the dominant kernel ----
    prof_steps = max(3, min(args.steps, 10))
    ctx.set_profiling(True)
    for _ in range(prof_steps):
        flush_buf.zero_()
        R.stage(res)
    prof = ctx.profile_read()
    ctx.set_profiling(False)
    kern = {}
    for kname, (cnt, ms) in prof.items():
        kern[kname] = {"launches_per_step": cnt / prof_steps, "ms_per_step": ms / prof_steps, "ms_per_launch": ms / cnt}
    step_kernel_ms = sum(v["ms_per_step"] for v in kern.values())
    top = max(kern, key=lambda k: kern[k]["ms_per_step"])
    facts = profile_facts(name)
    # `roofline` = the kernel with the largest share of the step (among those with an algorithmic-bytes model);
    # `rooflines` = the same figures for every modelled kernel, largest share first
    roofline, rooflines = None, []
    for kname in sorted(kern, key=lambda k: -kern[k]["ms_per_step"]):
        ab = algorithmic_bytes(kname, R.src_nv, R.src_nf, R.cut_nv, R.cut_nf, counts)
        fact = facts.get(kname, {})
        if kname.startswith("k_tests_exact") and (counts["n_full"] > 0 or (not kname.endswith("_tri") and counts["n_exact"] > 0)):
            # the exact-expansion kernel is compute / local-memory bound: what is reported is the FP64 pipe's busy share from
            # the committed ncu capture of this workload (sm__pipe_fp64_cycles_active.avg.pct_of_peak_sustained_active)
            pct = fact.get("fp64_pipe_pct")
            rooflines.append({"kernel": kname, "bound": "fp64", "achieved": pct, "peak": 100.0, "unit": "% of FP64 pipe cycles",
                              "frac": (pct / 100.0) if pct is not None else None, "traffic": fact.get("dram_bytes_per_launch"),
                              "peak_source": "ncu capture of this workload under profiles/" if pct is not None else "not captured",
                              "tests_per_launch": counts["n_exact"], "ms_per_launch": kern[kname]["ms_per_launch"],
                              "share_of_step": kern[kname]["ms_per_step"] / step_kernel_ms})
            continue
        if kname == "k_tri_resolve" and counts["n_exact"] > 0:
            # the kernel that settles the stage-A failures (exact sign from a 24-term error-free sum, in registers): besides its
            # byte model below, the FP64 pipe's busy share from the committed ncu capture of this workload
            pct = fact.get("fp64_pipe_pct")
            rooflines.append({"kernel": kname, "bound": "fp64", "achieved": pct, "peak": 100.0, "unit": "% of FP64 pipe cycles",
                              "frac": (pct / 100.0) if pct is not None else None, "traffic": fact.get("dram_bytes_per_launch"),
                              "peak_source": "ncu capture of this workload under profiles/" if pct is not None else "not captured",
                              "tests_per_launch": counts["n_exact"], "ms_per_launch": kern[kname]["ms_per_launch"],
                              "share_of_step": kern[kname]["ms_per_step"] / step_kernel_ms})
        if ab is None:
            continue
        achieved = ab / (kern[kname]["ms_per_launch"] * 1e-3) / 1e9
        if achieved / peak > 1.2:
            continue  # the mean-size byte model does not describe this launch (e.g. the one-face side of a planar section)
        entry = {"kernel": kname, "bound": "hbm", "achieved": achieved, "peak": peak, "unit": "GB/s", "frac": achieved / peak,
                 "traffic": fact.get("dram_bytes_per_launch"), "peak_source": peak_src, "algorithmic_bytes_per_launch": ab,
                 "ms_per_launch": kern[kname]["ms_per_launch"], "share_of_step": kern[kname]["ms_per_step"] / step_kernel_ms,
                 "fp64_pipe_pct": fact.get("fp64_pipe_pct")}
        rooflines.append(entry)
        if roofline is None:
            roofline = entry
    build_kernels = ("k_vertex_bbox", "k_face_codes", "k_face_bbox", "k_morton", "k_leaves")
    stage_ms = {
        "build_ms": sum(kern[k]["ms_per_step"] for k in kern if k.startswith(build_kernels) or k == "onesweep_pass_u32_kv"),
        "traverse_ms": sum(kern[k]["ms_per_step"] for k in kern if k == "k_traverse"),
        "narrowphase_ms": sum(kern[k]["ms_per_step"] for k in kern if "k_tests" in k or "k_planes" in k or "k_tri_" in k),
        "pair_and_record_order_ms": sum(kern[k]["ms_per_step"] for k in kern if "u64" in k or k.startswith("k_pair") or "k_rank_sort" in k
                                        or "k_make_keys" in k or "k_gather" in k),
        "sum_of_kernels_ms": step_kernel_ms,
    }
    # whole-stage and whole-build figures against the HBM roofline by SURVEY §8-d's byte model (B_build = 24V + 296F per mesh)
    b_build = 24.0 * (R.src_nv + R.cut_nv) + 296.0 * (R.src_nf + R.cut_nf)
    b_rest = 48.0 * counts["n_node_tests"] + 8.0 * counts["n_pairs"] + 8.0 * counts["n_pairs"] + 128.0 * counts["n_tests_reference"]
    stage_roofline = {"bytes_model": "SURVEY 8-d: build 24V+296F per mesh, traversal 48/test + 8/pair, cull+predicates 8/pair + 128/test",
                      "build_frac_of_hbm_by_kernel_sum": b_build / (stage_ms["build_ms"] * 1e-3) / 1e9 / peak if stage_ms["build_ms"] else None,
                      "stage_frac_of_hbm_by_step_time": (b_build + b_rest) / (ms_per_step * 1e-3) / 1e9 / peak}

    # ---- sharded single dispatch (N > 1) ----
    sharded = None
    if world > 1:
        comm = make_comm()
        sh_steps = max(3, min(args.steps, 10))
        sharded = {"this_workload": measure_sharded(R, comm, sh_steps, max(args.warmup, 3), res)}
        sharded["this_workload"]["single_gpu_ms"] = ms_per_step
        if name != "c5" and not args.no_sharded_c5:
            # the split pays where traversal + narrowphase dominate: BASELINE config 5
            res.free()
            res2.free()
            R.free()
            s5, c5, f5 = workload("c5")
            R5 = Resident(s5, c5)
            r5 = stage.Result(ctx)
            settle_capacity(lambda: R5.stage(r5), r5)
            one = timed_loop(lambda: R5.stage(r5), 3, 3) / 3
            sec = measure_sharded(R5, comm, 3, 3, r5)
            sec["single_gpu_ms"] = one
            sec["speedup_vs_single_gpu"] = one / sec["ms_per_dispatch"]
            sec["workload"] = WORKLOADS["c5"]
            sharded["c5"] = sec
        L.mcb200_comm_destroy(comm)

    # ---- cpu baseline (rank 0, N == 1 only) ----
    cpu = None
    if rank == 0 and world == 1 and not args.no_cpu_baseline:
        cores = os.cpu_count() or 1
        with tempfile.TemporaryDirectory() as td:
            (csrc, ccut, cflags), sample_note = cpu_workload(name)
            if have_reference_binary():
                write_input(csrc, ccut, cflags, os.path.join(td, "in.mcb"))
                runs, npairs_ref = [], 0
                for _ in range(2):
                    w, sms, npairs_ref, helpers = reference_streams([os.path.join(td, "in.mcb")], 1, cores, td)
                    runs.append(sms[0])
                best = min(runs)
                cpu = {"value": npairs_ref / (best * 1e-3), "unit": UNIT, "cores": cores, "kind": "reference",
                       "ms_per_step": best, "pairs": npairs_ref,
                       "sample": sample_note + "2 whole intersect stages (best of 2): one mcDispatch of the unmodified "
                                 "reference each, cut short after its narrowphase; stage ms = the reference's own timers "
                                 "(build_oibvh x2, intersectOIBVHs, kernel.cpp:1781-3231), helper pool = cores-1 threads"}
            else:
                ms, npairs_ref, nrec = oracle_port_stage_once(csrc, ccut, cflags)
                cpu = {"value": npairs_ref / (ms * 1e-3), "unit": UNIT, "cores": 1, "kind": "port", "ms_per_step": ms,
                       "pairs": npairs_ref, "sample": sample_note + "1 run of the single-threaded oracle port"}

    # ---- e2e_dropin (--dropin): what a LIVE mcDispatch pays for the stage through the adapter + hooked kernel ----
    dropin = None
    if rank == 0 and args.dropin:
        dropin = measure_dropin(src, cut, flags)

    if rank == 0:
        cfg.update({"pairs_per_dispatch": counts["n_pairs"], "src_faces": R.src_nf, "cut_faces": R.cut_nf,
                    "l2": "256 MiB write between timed steps", **counts, "status": status})
        line = {
            "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps, "warmup": args.warmup,
            "ms_per_step": ms_per_step, "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f64",
            "data": "synthetic", "config": cfg, "clocks": clocks, "e2e": e2e, "gpu_launches": int(launches), "roofline": roofline,
            "cpu_baseline": cpu, "rooflines": rooflines, "stage_roofline": stage_roofline, "stage_ms": stage_ms, "kernels": kern,
            "top_kernel": top,
        }
        if sharded:
            line["sharded"] = sharded
        if dropin:
            line["e2e_dropin"] = dropin
        sys.stdout.flush()
        os.dup2(real_stdout, 1)
        print(json.dumps(line), flush=True)
        os.dup2(2, 1)
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()


def measure_dropin(src, cut, flags):
    """Two whole mcDispatch calls of the workload through the public C API of the HOOKED reference build
    (oracle/_ref/api_driver_hooked: unmodified objects + kernel.cpp with its narrowphase region replaced by one call) with
    the adapter loaded; the adapter's own wall-clock timers (MCB200_SHIM_TIMING) of the SECOND dispatch are summed.  The
    rest of mcDispatch is the reference's host code and takes tens of seconds at BASELINE sizes, hence opt-in."""
    drv = os.path.join(ROOT, "oracle", "_ref", "api_driver_hooked")
    nodump = os.path.join(ROOT, "oracle", "_ref", "libnodump.so")
    if not (os.path.exists(drv) and os.path.exists(nodump)):
        return {"unavailable": "oracle/_ref/api_driver_hooked is not built (needs /root/reference at build time)"}
    with tempfile.TemporaryDirectory() as td:
        write_input(src, cut, flags, os.path.join(td, "in.mcb"))
        env = dict(os.environ, LD_PRELOAD=nodump, MCB200_SHIM_TIMING="1")
        t0 = time.perf_counter()
        r = subprocess.run([drv, os.path.join(td, "in.mcb"), os.path.join(td, "out.mcb"), "--repeat", "2"], capture_output=True, text=True,
                           cwd=td, env=env)
        wall = time.perf_counter() - t0
        if r.returncode != 0:
            return {"unavailable": "api_driver_hooked failed: " + r.stderr[-300:]}
    entries = []  # (name, ms) of the top-level adapter entries (sub-steps are indented in the log)
    for ln in r.stderr.splitlines():
        if ln.startswith("[mcut_b200 shim] ") and not ln.startswith("[mcut_b200 shim]   "):
            name, _, ms = ln[len("[mcut_b200 shim] "):].rpartition(": ")
            entries.append((name, float(ms.split()[0])))
    half = len(entries) // 2
    second = entries[half:]
    by_name = {}
    for name, ms in second:
        by_name[name] = by_name.get(name, 0.0) + ms
    return {"stage_ms_per_mcdispatch": sum(ms for _, ms in second), "adapter_entries_ms": by_name,
            "first_dispatch_stage_ms": sum(ms for _, ms in entries[:half]), "process_wall_s_for_2_dispatches": wall,
            "call": "mcDispatch (public C API) -> hooked libmcut -> libmcut_b200_shim -> C-ABI",
            "note": "user arrays are pageable host memory; includes uploads, device work, read-backs and filling the reference's own containers"}


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=5)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--workload", default="c2", choices=sorted(WORKLOADS))
    ap.add_argument("--batch", type=int, default=0, help="batch workloads: number of dispatches (default 256 / 2000)")
    ap.add_argument("--lanes", type=int, default=0, help="batch workloads: contexts in flight per GPU")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-sharded-c5", action="store_true")
    ap.add_argument("--dropin", action="store_true", help="also time a live mcDispatch through the adapter (adds about a minute)")
    args = ap.parse_args()
    if args.warmup < 3 and args.impl == "ours":
        args.warmup = 3
    if args.impl == "reference":
        run_reference_arm(args)
    else:
        run_ours(args)


if __name__ == "__main__":
    main()
