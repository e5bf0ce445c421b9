#!/usr/bin/env python
"""bench.py -- mesh -> SSVDAG build on B200 (BASELINE.json metric), one JSON line on stdout.

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference] [--workload NAME]

One "step" = one full pass of the hot path over the workload: svb_build (voxelize + per-level DAG
reduction) followed by svb_to_sdag (mirror-symmetry reduction), i.e. triangle soup -> SSVDAG node
arrays.  `value` is measured with the triangle soup already resident in HBM; `e2e` goes through the
public C ABI with HOST buffers every step (H2D of the triangles from pinned memory, build, toSDAG, the
.ssvdag image written on the GPU and copied D2H into pinned host memory by rank 0).  `--impl reference` times the UNMODIFIED reference svbuilder
(oracle/_ref/svbuilder_ref, compiled from /root/reference by oracle/Makefile) on the host cores, on
a bounded sample of the same workload (see WORKLOADS[*]["cpu_sample"]).

Default workload: BASELINE.json configs[2], the configuration its metric is quoted on -- the
procedural city (256x256 lots, ~11 M triangles) at 16384^3 (levels 14, step 4); it fits one B200
(the build streams it through ~20 tile batches).
"""
from __future__ import annotations

import argparse
import json
import os
import re
import subprocess
import sys
import tempfile
import threading
import time
from pathlib import Path

ROOT = Path(__file__).resolve().parent
sys.path.insert(0, str(ROOT))

WORKLOADS = {
    # name: mesh generator + args, octree levels, build step, and the bounded CPU sample
    "city_16k": dict(mesh="city", kw=dict(lots=256), levels=14, step=4, grid="16384^3",
                     cpu_sample=dict(mesh="city", kw=dict(lots=32), levels=11, step=2,
                                     note="same generator at lots=32, 2048^3 (levels 11, step 2): every lot spans 64 voxels as in the full workload; 1/64 of its ground area")),
    # BASELINE.json configs[3]: the same city, CSVDAG (svbuilder -c): build + mergeAcrossAllLevels -> <base>-multi.svdag
    "city_16k_c": dict(mesh="city", kw=dict(lots=256), levels=14, step=4, grid="16384^3", cross=True,
                       cpu_sample=dict(mesh="city", kw=dict(lots=32), levels=11, step=2, cross=True,
                                       note="same generator at lots=32, 2048^3 (levels 11, step 2), svbuilder -c: every lot spans 64 voxels as in the full workload; 1/64 of its ground area")),
    "terrain_4k": dict(mesh="terrain", kw=dict(n=1024), levels=12, step=3, grid="4096^3",
                       cpu_sample=dict(mesh="terrain", kw=dict(n=257), levels=10, step=2,
                                       note="same generator at n=257, 1024^3 (levels 10, step 2): 4 voxels per grid cell as in the full workload")),
    "spongeball_1k": dict(mesh="sphere_menger", kw=dict(), levels=10, step=1, grid="1024^3",
                          cpu_sample=dict(mesh="sphere_menger", kw=dict(), levels=10, step=1, note="the full workload")),
    "composite_64k": dict(mesh="composite", kw=dict(n_terrain=1024, lots=256), levels=16, step=7, grid="65536^3",
                          cpu_sample=dict(mesh="composite", kw=dict(n_terrain=129, lots=32), levels=11, step=2,
                                          note="same generator at n_terrain=129, lots=32, 2048^3 (levels 11, step 2)")),
    "city_small": dict(mesh="city", kw=dict(lots=32), levels=11, step=2, grid="2048^3",
                       cpu_sample=dict(mesh="city", kw=dict(lots=16), levels=10, step=2, note="lots=16 at 1024^3")),
}
# reference pins (tests/golden/*.json, minted by tests/golden/make_fullsize.py from the UNMODIFIED reference svbuilder):
# the bench asserts the SHA-256 of its .ssvdag image and the node counts against them at every N
GOLDEN = {"city_16k": "fullsize_city16k.json", "city_16k_c": "fullsize_city16k.json", "terrain_4k": "size_terrain4k.json", "spongeball_1k": "size_spongeball1k.json"}
# committed wall times of the unmodified reference on the same generator at growing sizes (8 cores of the build container)
REF_SCALING = ["midsize_city4k.json", "bigsize_city8k.json", "fullsize_city16k.json"]
METRIC = "mesh->SSVDAG build throughput (Gvoxel/s; BASELINE.json: build time at 16K^3 + dedup HBM GB/s vs peak)"


def load_pkg():
    import __graft_entry__ as g
    return g._pkg()


def measured_peak_gbs():
    p = ROOT / "MEASURED_PEAKS.json"
    if p.exists():
        try:
            return float(json.loads(p.read_text())["hbm_gbs"]), "measured (MEASURED_PEAKS.json)"
        except Exception:
            pass
    return 6650.0, "fallback (B200_PROFILING.md)"


class ClockSampler:
    """nvidia-smi clocks / throttle reasons during the timed region (B200_PROFILING.md recipe)."""
    Q = "index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap"

    def __init__(self, devices=(0,)):
        self.dev = ",".join(str(d) for d in devices)
        self.rows = []
        self.proc = None
        self.first = 0

    def mark(self):
        """The timed region starts here: rows read so far (warm-up) are not counted."""
        self.first = len(self.rows)

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits", "-lms", "200", "-i", self.dev],
                                         stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.t = threading.Thread(target=self._read, daemon=True)
            self.t.start()
        except Exception:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.rows.append([x.strip() for x in line.split(",")])

    def stop(self):
        if not self.proc:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        self.proc.terminate()
        try:
            self.proc.wait(timeout=2)
        except Exception:
            self.proc.kill()
        self.rows = self.rows[self.first:] or self.rows[-1:]     # (a timed region shorter than one polling interval: the last row before it)
        sm = sorted(float(r[1]) for r in self.rows if len(r) > 2 and r[1].replace(".", "").isdigit())
        mx = [float(r[2]) for r in self.rows if len(r) > 2 and r[2].replace(".", "").isdigit()]
        reasons = set()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for r in self.rows:
            for k, nm in enumerate(names):
                if len(r) > 5 + k and r[5 + k].lower().startswith("active"):
                    reasons.add(nm)
        return {"sm_mhz": sm[len(sm) // 2] if sm else None, "sm_max_mhz": max(mx) if mx else None,
                "reasons": sorted(reasons), "samples": len(self.rows)}


def golden_pin(workload):
    p = ROOT / "tests" / "golden" / GOLDEN.get(workload, "-")
    if not p.exists():
        return None
    try:
        return json.loads(p.read_text())
    except Exception:
        return None


def reference_scaling():
    """Committed timings of the unmodified reference (same city generator, every lot spans 64 voxels) at growing sizes."""
    rows = []
    for name in REF_SCALING:
        p = ROOT / "tests" / "golden" / name
        if not p.exists():
            continue
        g = json.loads(p.read_text())
        rows.append({"golden": name, "workload": g.get("workload"), "triangles": g.get("triangles"), "voxels": g.get("Voxels"),
                     "reference_seconds": g.get("reference_seconds"), "cores": g.get("cores", 8),
                     "Gvoxel_per_s": (g["Voxels"] / g["reference_seconds"] / 1e9) if g.get("reference_seconds") else None})
    return rows


# ------------------------------------------------------------------------------------- reference arm
def run_reference_once(orc, tris, levels, step, threads, cross=False):
    with tempfile.TemporaryDirectory() as td:
        r = orc.run_reference(td, tris, levels, step, threads=threads, cross=cross)
    vox = int(re.search(r"Voxels:\s+.*\((\d+)\)", r["log"]).group(1))
    return vox, r["seconds"]


def run_port_once(orc, tris, levels, step):
    t0 = time.time()
    o = orc.OracleOctree(tris)
    o.build(levels, step)
    o.to_sdag()
    o.encode("ssvdag")
    return o.stat("nTotalVoxels"), time.time() - t0


def cpu_arm(pkg, wl, steps, warmup):
    """Times the reference's own CPU svbuilder (or, if its binary did not travel, the oracle port)."""
    from oracle import oracle as orc
    cs = wl["cpu_sample"]
    tris = pkg.meshgen.make_mesh(cs["mesh"], **cs["kw"])
    cores = os.cpu_count() or 1
    use_ref = orc.REF_BIN.exists()
    kind = "reference" if use_ref else "port"
    if not use_ref:
        cores = 1
    times, vox = [], 0
    for i in range(warmup + steps):
        if use_ref:
            vox, dt = run_reference_once(orc, tris, cs["levels"], cs["step"], cores, cross=bool(cs.get("cross")))
        else:
            vox, dt = run_port_once(orc, tris, cs["levels"], cs["step"])
        if i >= warmup:
            times.append(dt)
    avg = sum(times) / len(times)
    return {"value": vox / avg / 1e9, "unit": "Gvoxel/s", "cores": cores, "kind": kind,
            "sample": f"{cs['note']}; {tris.shape[0]} triangles, {vox} voxels, svbuilder wall {avg:.2f} s/step "
                      f"(bincache load + buildDAG + toSDAG + encoders + file writes), OMP_NUM_THREADS={cores}",
            "seconds_per_step": avg, "voxels": vox}


# ------------------------------------------------------------------------------------- our arm
def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=5)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--workload", default=os.environ.get("SVB_BENCH_WORKLOAD", "city_16k"), choices=sorted(WORKLOADS))
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-e2e", action="store_true", help="skip the end-to-end leg (profiling runs only; the line then carries e2e = null)")
    args = ap.parse_args()
    wl = WORKLOADS[args.workload]
    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    cross = bool(wl.get("cross"))
    config = {"workload": f"{args.workload}: procedural {wl['mesh']} mesh {wl['kw']} at {wl['grid']} (levels {wl['levels']}, step {wl['step']}) -> SVDAG -> " + ("CSVDAG (cross-level merge)" if cross else "SSVDAG"),
              "l2_policy": "inputs larger than L2 (every pass streams GBs of freshly written pair/node arrays)"}

    if args.impl == "reference":
        if rank != 0:
            return 0
        pkg = load_pkg()
        # Every step runs the bounded SAMPLE of the workload (the full 16K^3 job is hours of reference CPU time), and the
        # line says so: config.workload names what was actually built; `sample_of` names the workload it stands in for;
        # `reference_scaling` carries the committed full-size timings of the same unmodified binary, so the cost of the
        # reference at the real size is a measured number, not an extrapolation of this line.
        steps, warmup = max(1, args.steps), max(0, args.warmup)
        cb = cpu_arm(pkg, wl, steps, warmup)
        cs = wl["cpu_sample"]
        config = {"workload": f"BOUNDED SAMPLE of {args.workload}: procedural {cs['mesh']} mesh {cs['kw']} (levels {cs['levels']}, step {cs['step']}) -> SVDAG -> SSVDAG; {cs['note']}",
                  "sample_of": args.workload, "full_workload": config["workload"]}
        line = {"impl": "reference", "metric": METRIC, "value": cb["value"], "unit": "Gvoxel/s", "n_gpus": args.gpus,
                "steps": steps, "warmup": warmup, "ms_per_step": cb["seconds_per_step"] * 1e3, "higher_is_better": True,
                "scaling": "strong", "vs_baseline": None, "dtype": "f64", "data": "synthetic", "config": config,
                "cpu_baseline": {k: cb[k] for k in ("value", "unit", "cores", "kind", "sample")},
                "e2e": {"value": cb["value"], "unit": "Gvoxel/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
                "reference_scaling": reference_scaling(),
                "gpu_launches": 0}
        print(json.dumps(line))
        return 0

    import numpy as np
    import torch
    import torch.distributed as dist
    pkg = load_pkg()
    if not torch.cuda.is_available():
        print(json.dumps({"error": "no CUDA device: this framework has no CPU fallback"}))
        return 2
    torch.cuda.set_device(local_rank)
    if world > 1:
        dist.init_process_group("nccl", device_id=torch.device("cuda", local_rank))

    tris = pkg.meshgen.make_mesh(wl["mesh"], **wl["kw"])
    T = tris.shape[0]
    v = tris.reshape(-1, 3)
    bbox = (v.min(axis=0).astype(np.float64), v.max(axis=0).astype(np.float64))
    L, S = wl["levels"], wl["step"]
    pinned = torch.from_numpy(tris.reshape(-1)).pin_memory()

    # N > 1: the context works on a torch-owned stream, so the exchange buffers (torch tensors), the NCCL collectives and the
    # library's kernels are ordered by one stream that outlives all of them
    oct_ = pkg.GeomOctree(device=local_rank, stream=torch.cuda.Stream(device=local_rank) if world > 1 else None)
    shard = dict(sharded=True) if world > 1 else {}

    out_kind = "svdag" if cross else "ssvdag"     # the file the step ends in: <base>-multi.svdag or <base>.ssvdag

    def post(o):
        if not cross:
            return o.to_sdag()
        cm = o.cross_merge()                      # mergeAcrossAllLevels (geom_octree_extension.cpp:1192-1544)
        cm["msSdag"] = cm["msCrossMerge"]
        cm["nNodesSDAG"] = 0
        return cm

    def step_resident():
        st = oct_.build(L, S, bbox=bbox, **shard)
        sd = post(oct_)
        return st, sd

    h2d_bytes = [int(T) * 36]

    def step_e2e():
        # H2D from pinned host memory: the whole soup on one GPU; with N ranks every rank copies its 1/N slice over its own
        # PCIe link and NCCL all-gathers the rest over NVLink (sharded.set_triangles_sharded)
        if world > 1:
            h2d_bytes[0] = pkg.sharded.set_triangles_sharded(oct_, pinned, T)
        else:
            oct_.set_triangles_ptr(pinned.data_ptr(), T)
        st = oct_.build(L, S, bbox=bbox, **shard)
        sd = post(oct_)
        # the file image is written on the GPU (svb_encode.cu) and lands in pinned host memory: D2H of the finished file,
        # once, by rank 0 (the file has one writer)
        img = pkg.encoders.encode_view(oct_, out_kind) if rank == 0 else b""
        return st, sd, img

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    # ---- resident-input timing (value)
    oct_.set_triangles_ptr(pinned.data_ptr(), T)
    # one nvidia-smi poller for the whole job (rank 0, all N GPUs), started BEFORE the warm-up steps: its start-up (NVML
    # initialisation, hundreds of ms of driver traffic) and N pollers side by side used to land in the first timed steps --
    # at N = 8 the resident figure came out 7 ms per step above the end-to-end one that ran right after it
    sampler = ClockSampler(range(world) if world > 1 else (local_rank,))
    if rank == 0:
        sampler.start()
    for _ in range(args.warmup):
        step_resident()
    # CUDA events around every launch of the roofline kernel (the "emit" family) accumulate inside the library and are fetched
    # after the timed region; the other families are bracketed in ONE extra step after the timed regions (`kernels`,
    # `roofline_dedup`): ~1400 event records per build between the launches cost the timed figure 4 % (resident 623 ms next to an
    # end-to-end step of 598 ms on the same box)
    oct_.set_profiling(3)
    prof, launches = [], 0
    barrier()
    sampler.mark()
    t0 = time.perf_counter()
    dev_ms = 0.0
    for _ in range(args.steps):
        st, sd = step_resident()
        launches += st["nKernelLaunches"] + sd["nKernelLaunches"]
        dev_ms += st["msTotal"] + sd["msSdag"]
    barrier()
    elapsed = time.perf_counter() - t0
    clocks = sampler.stop()
    prof = oct_.profile()
    oct_.set_profiling(False)

    # ---- end-to-end timing through the public API with host buffers
    elapsed_e2e, d2h, img = float("nan"), 0, b""
    if not args.no_e2e:
        step_e2e()
        barrier()
        t1 = time.perf_counter()
        for _ in range(args.steps):
            st_e, sd_e, img = step_e2e()
        barrier()
        elapsed_e2e = time.perf_counter() - t1
        d2h = len(img) + 4 * sum(oct_.level_sizes()[:-1])       # the image + the per-level reference counts of the SSVDAG node order

    # ---- one more resident step with every kernel family bracketed (not timed: the breakdown below)
    oct_.set_profiling(2)
    barrier()
    step_resident()
    barrier()
    prof_all = oct_.profile()
    oct_.set_profiling(False)

    per_rank = None
    if world > 1:
        tmax = torch.tensor([elapsed, elapsed_e2e], dtype=torch.float64, device="cuda")
        dist.all_reduce(tmax, op=dist.ReduceOp.MAX)
        elapsed, elapsed_e2e = float(tmax[0]), float(tmax[1])
        # where every rank spent its last timed step: device times of its share, host wall clock of the two phases
        mine = torch.tensor([st["msVoxelize"], st["msDedup"], st["msFinalize"], st["msTotal"], st.get("wallLocalMs", 0.0), st.get("wallMergeMs", 0.0),
                             float(st["nBatches"]), sd["msSdag"]], dtype=torch.float64, device="cuda")
        allr = torch.empty(world * mine.numel(), dtype=torch.float64, device="cuda")
        dist.all_gather_into_tensor(allr, mine)
        per_rank = {k: [round(float(x), 2) for x in allr.view(world, -1)[:, i]] for i, k in enumerate(
            ["ms_voxelize", "ms_dedup", "ms_finalize", "ms_total_device", "wall_local_ms", "wall_merge_ms", "batches", "ms_sdag"])}

    # ---- parity pin: the .ssvdag image this very run produced (N ranks, NCCL merge and all) against the unmodified
    #      reference's file for the same input (tests/golden/*.json) -- asserted at every N, never assumed
    import hashlib
    if rank == 0 and not len(img):
        img = pkg.encoders.encode_view(oct_, out_kind)
    img = bytes(img)
    sha = hashlib.sha256(img).hexdigest() if rank == 0 else None
    pin = golden_pin(args.workload)
    parity = {"ssvdag_sha256": sha, "ssvdag_bytes": len(img), "reference_pin": None, "ok": None}
    if cross:
        parity["file"] = "-multi.svdag"
        parity["cross_merge"] = {"ms_per_step": sd["msCrossMerge"], "nodes_eliminated": sd["nCrossLevelMerged"], "dag_nodes_after": sd["nNodesDAG"]}
        pin = None                                # (reference pins of -c exist at 4096^3: tests/golden/midsize_city4k.json["cross"], checked by the GPU suite)
    if pin and rank == 0:
        want = pin["files"]["ssvdag"]
        checks = {"ssvdag_sha256": sha == want["sha256"], "ssvdag_bytes": len(img) == want["bytes"], "voxels": st["nTotalVoxels"] == pin["Voxels"],
                  "svo_nodes": st["nNodesSVO"] == pin["SVO Nodes"], "dag_nodes": st["nNodesDAG"] == pin["DAG Nodes"], "sdag_nodes": sd["nNodesSDAG"] == pin["SDAG Nodes"]}
        parity.update(reference_pin=f"tests/golden/{GOLDEN[args.workload]} (unmodified reference svbuilder, {pin.get('reference_seconds', 0):.0f} s on {pin.get('cores', 8)} cores)",
                      reference_sha256=want["sha256"], checks=checks, ok=all(checks.values()))

    vox = st["nTotalVoxels"]
    sec = elapsed / args.steps
    sec_e2e = elapsed_e2e / args.steps
    # ---- roofline.  The largest kernel that is a plain HBM stream is k_emit_warp (child-pair emission of the voxelizer, ~15 % of the
    #      GPU time; the larger classify kernels are issue bound): achieved = algorithmic bytes of all its launches
    #      (recorded per launch by the library: 20 B per parent pair read, 10 B (+ 4 B first touch where tracked) per child pair
    #      written) / their CUDA-event time on the build stream.  `traffic` = DRAM bytes per launch from the ncu capture
    #      (profiles/ncu_traffic.json).  The dedup family BASELINE.json's metric names gets its own block, `roofline_dedup`, with
    #      REAL DRAM bytes per node from ncu (not the "effective" figures of round 1).
    peak, peak_src = measured_peak_gbs()
    def families(records, nsteps, skip=()):   # per-step totals of every kernel family
        out = {}
        for r in records:
            if r["name"] in skip:
                continue
            f = out.setdefault(r["name"], {"launches": 0, "ms": 0.0, "units": 0, "out": 0, "bytes_survey": 0.0})
            f["launches"] += 1; f["ms"] += r["ms"]; f["units"] += r["n_in"]; f["out"] += r["n_out"]; f["bytes_survey"] += r["bytes"]
        for f in out.values():
            for k in f:
                f[k] = f[k] / nsteps
        return out
    fam = families(prof, args.steps)                         # the timed region: the roofline kernel
    fam.update(families(prof_all, 1, skip=tuple(fam)))       # the extra instrumented step: everything else
    emit = fam.get("emit")
    roof = None
    tj_all = {}
    tp = ROOT / "profiles" / "ncu_traffic.json"     # dram__bytes_read+write of one `ncu --set full` capture per kernel
    if tp.exists():
        try:
            tj_all = json.loads(tp.read_text())
        except Exception:
            tj_all = {}
    if emit and emit["ms"] > 0:
        alg = emit["bytes_survey"]
        ach = alg / (emit["ms"] * 1e-3) / 1e9
        traffic, traffic_src = None, None
        tj = tj_all.get("k_emit_warp")
        if tj:
            traffic = float(tj["dram_bytes_per_algorithmic_byte"]) * alg / emit["launches"]
            traffic_src = tj.get("source")
        roof = {"kernel": "k_emit_warp (voxelizer: child-pair emission; the largest plain HBM stream of the build)", "bound": "hbm", "achieved": ach, "peak": peak,
                "unit": "GB/s", "frac": ach / peak, "traffic": traffic, "traffic_source": traffic_src,
                "algorithmic_bytes_per_launch": alg / emit["launches"], "peak_source": peak_src,
                "bytes_per_unit": "20 B per parent pair (11 B pair + 9 B node fields) + 10 B per child pair (+ 4 B first touch where tracked)",
                "units_per_launch": {"parent_pairs": emit["units"] / emit["launches"], "child_pairs": emit["out"] / emit["launches"]},
                "avg_launch_ms": emit["ms"] / emit["launches"], "share_of_step": emit["ms"] / (dev_ms / args.steps if dev_ms else 1.0),
                "note": "all launches of the step, the many small upper-level ones included (the deepest launches alone run at 0.65 of the peak, issue bound: profiles/r2h_ncu_emit_dedup_city16k.md)"}
        ch = fam.get("children")
        if ch and ch["ms"] > 0:
            a2 = ch["bytes_survey"] / (ch["ms"] * 1e-3) / 1e9
            roof["k_children"] = {"achieved": a2, "frac": a2 / peak, "bytes_per_unit": "13 B per node + 8 B per child node", "avg_launch_ms": ch["ms"] / ch["launches"]}
    # per-level dedup kernels (BASELINE.json: "dedup HBM GB/s vs peak"): nodes x DRAM bytes per node (ncu) / CUDA-event time of the
    # family's launch groups (which also hold the flag read-backs between the passes: a lower bound of the kernels' own rate)
    dd = tj_all.get("dedup", {})
    roofline_dedup = {}
    for famname, key, label in (("dedup_leaf", "k_leaf_known", "leaf level (k_leaf_known / k_leaf_lazy)"),
                                ("dedup_k64", "k_insert_k64", "4^3 level (k_insert_k64)"),
                                ("dedup_inner", "k_insert_inner", "inner levels (k_insert + k_winner / k_convert on new entries)")):
        f = fam.get(famname)
        if not f or f["ms"] <= 0:
            continue
        bpn = float(dd.get(key, {}).get("dram_bytes_per_node", 0.0))
        achd = bpn * f["units"] / (f["ms"] * 1e-3) / 1e9
        roofline_dedup[famname] = {"kernels": label, "nodes_per_step": int(f["units"]), "ms_per_step": f["ms"],
                                   "dram_bytes_per_node": bpn, "achieved": achd, "peak": peak, "unit": "GB/s", "frac": achd / peak,
                                   "source": dd.get(key, {}).get("source")}
    dedup_eff = roofline_dedup or None
    kernels = {k: {"launches": int(f["launches"]), "ms_per_step": f["ms"], "units_per_step": int(f["units"]),
                   "measured": "timed region" if k == "emit" else "one extra instrumented step"}
               for k, f in sorted(fam.items())}

    line = {"metric": METRIC, "value": vox / sec / 1e9, "unit": "Gvoxel/s", "n_gpus": world, "steps": args.steps, "warmup": args.warmup,
            "ms_per_step": sec * 1e3, "higher_is_better": True, "scaling": "strong", "vs_baseline": None, "dtype": "f64",
            "data": "synthetic", "config": config, "build_s": sec, "device_ms_per_step": dev_ms / args.steps,
            "voxels": vox, "triangles": int(T), "nodes": {"svo": st["nNodesSVO"], "dag": st["nNodesDAG"], "sdag": sd["nNodesSDAG"]},
            "tiles": st["nTiles"], "batches": st["nBatches"], "pairs": st["nPairsTotal"], "exact_retests": st["nExactTests"],
            "e2e": None if args.no_e2e else {"value": vox / sec_e2e / 1e9, "unit": "Gvoxel/s", "h2d_bytes_per_step": h2d_bytes[0] + (4 * sum(oct_.level_sizes()[:-1]) if rank == 0 else 0),
                                             "d2h_bytes_per_step": int(d2h), "seconds_per_step": sec_e2e, "ssvdag_bytes": len(img)},
            "gpu_launches": int(launches), "clocks": clocks, "roofline": roof, "roofline_dedup": dedup_eff, "kernels": kernels, "parity": parity, "per_rank": per_rank}
    if rank == 0:
        if not args.no_cpu_baseline and world == 1:
            try:
                cb = cpu_arm(pkg, wl, 1, 0)
                line["cpu_baseline"] = {k: cb[k] for k in ("value", "unit", "cores", "kind", "sample")}
            except Exception as e:  # the baseline is reported, never required
                line["cpu_baseline"] = {"value": None, "unit": "Gvoxel/s", "cores": 0, "kind": "reference", "sample": f"failed: {e}"}
        print(json.dumps(line))
    # orderly teardown: the context (and its CUDA stream, which torch's allocator knows as the allocation stream of the exchange
    # buffers) goes before the process group and the interpreter
    torch.cuda.synchronize()
    if hasattr(oct_, "_dev_tris"):
        del oct_._dev_tris
    torch.cuda.empty_cache()
    oct_.close()
    if world > 1:
        dist.destroy_process_group()
    if rank == 0 and parity["ok"] is False:
        print(f"PARITY FAILURE: result differs from the reference pin: {parity['checks']}", file=sys.stderr)
        return 3
    return 0


if __name__ == "__main__":
    sys.exit(main())
