#!/usr/bin/env python
"""bench.py — headline benchmark of the voxelization hot path (BASELINE.json metric).

    python bench.py --gpus N --steps K --warmup W            # this repo's CUDA path (one rank per GPU)
    python bench.py --impl reference --steps K --warmup W    # the reference's own CPU path on host cores

Workload (config.workload): BASELINE.json configs[3] — the 10,025,280-triangle geodesic icosphere
(nu=708, radius 1024, one world unit per voxel) surface-voxelized at 2048^3.  It is the configuration
the metric ("Mtriangles/s ... 2048^3 ... 10M-triangle mesh") is quoted on and it fits one GPU.  With N
GPUs the SAME mesh is voxelized once: rank r owns z-slab r of the bit table (disjoint slices, no
reduction, no data-path collective), every rank reads the replicated triangle soup.  Total work is
fixed, so scaling is "strong".

A step = one full voxelization of the rank's slab.  Surface workloads run through the prepared-mesh interface
(voxb200_mesh_*, the resident / per-frame API): the TIMED REGION is one voxb200_mesh_update — the re-ordering of the
caller's triangle soup into the tile records, from the caller's order, buffers reused — followed by K voxb200_mesh_voxelize
calls (one captured CUDA graph replayed K times), so the preparation is inside the headline number, amortised over the K
steps the command line asks for.  The line also carries "resident" (the K steps alone), "prepare_ms", and "one_shot"
(voxb200_surface on the caller's order: what a single voxelization without preparation costs).
  value : Mtri/s with triangles resident in HBM (device time, CUDA events, max over ranks)
  e2e   : Mtri/s through voxb200_voxelize_host_indexed — pinned host mesh -> H2D -> tile records -> voxelize -> the table in pinned
          host memory (dense D2H, or the non-zero words + host-thread expansion: voxb200_download_table); N > 1: voxb200_voxelize_host_multi
"""
import argparse
import json
import os
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

import numpy as np  # noqa: E402

WORKLOADS = {
    # name: (mesh generator args, gridsize, solid, description)
    "config4": dict(nu=708, radius=1024.0, G=2048, solid=False,
                    desc="icosphere nu=708 r=1024 (10,025,280 tris) surface @2048^3"),
    "config3": dict(nu=224, radius=512.0, G=1024, solid=True,
                    desc="icosphere nu=224 r=512 (1,003,520 tris) solid @1024^3"),
    "config2": dict(nu=0, radius=0.0, G=1024, solid=False, desc="bunny (5,110 tris) surface @1024^3"),
    # the regime of the reference README's timing table (70k-triangle bunny): ~8-voxel triangles at 1024^3
    "readme1024": dict(nu=59, radius=512.0, G=1024, solid=False, desc="icosphere nu=59 r=512 (69,620 tris) surface @1024^3"),
    "readme2048": dict(nu=59, radius=1024.0, G=2048, solid=False, desc="icosphere nu=59 r=1024 (69,620 tris) surface @2048^3"),
}
METRIC = "Mtriangles/s"


def load_mesh(name):
    from cuda_voxelizer_b200 import meshgen
    w = WORKLOADS[name]
    if w["nu"] == 0:
        d = np.load(os.path.join(ROOT, "tests", "golden", "bunny.npz"))
        return d["verts"], d["faces"]
    return meshgen.icosphere(w["nu"], radius=w["radius"])


def expand_soup(verts, faces):
    return np.ascontiguousarray(verts[faces.reshape(-1)].reshape(-1, 9))


class ClockSampler:
    """Samples nvidia-smi clocks / throttle reasons during the timed region (B200_PROFILING.md recipe)."""
    Q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, index):
        self.index, self.rows, self.proc = index, [], None

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", "-i", str(self.index), "--query-gpu=" + self.Q, "--format=csv,noheader,nounits", "-lms", "100"],
                                         stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            threading.Thread(target=self._pump, daemon=True).start()
        except Exception:
            self.proc = None

    def _pump(self):
        for line in self.proc.stdout:
            self.rows.append([c.strip() for c in line.split(",")])

    def stop(self):
        if self.proc:
            time.sleep(0.15)
            self.proc.terminate()
            try:
                self.proc.wait(timeout=2)
            except Exception:
                self.proc.kill()
        sm, reasons, mx = [], set(), None
        for r in self.rows:
            try:
                sm.append(float(r[1])); mx = float(r[2])
                for name, col in (("hw_slowdown", 5), ("hw_thermal_slowdown", 6), ("sw_thermal_slowdown", 7), ("sw_power_cap", 8)):
                    if r[col].lower().startswith("active"):
                        reasons.add(name)
            except Exception:
                pass
        return {"sm_mhz": float(np.median(sm)) if sm else None, "sm_max_mhz": mx, "reasons": sorted(reasons), "samples": len(sm)}


def measured_peak_gbs():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        try:
            return float(json.load(open(p))["hbm_gbs"]), "measured (MEASURED_PEAKS.json hbm_gbs)"
        except Exception:
            pass
    return 6650.0, "fallback (B200_PROFILING.md 6.65 TB/s)"


def cpu_reference_run(verts, faces, G, solid, n_sample, threads):
    """One bounded run of the reference CPU path (oracle/_ref) or, if it was not built, the oracle port.
    Returns (ms, kind)."""
    import oracle
    f = np.ascontiguousarray(faces[:n_sample])
    if oracle.have_ref():
        _, ms = oracle.ref_voxelize(verts, f, G, solid=solid, morton=False, threads=threads, return_ms=True)
        return ms, "reference"
    os.environ["OMP_NUM_THREADS"] = str(threads)
    mn, mx, unit = oracle.voxinfo(verts, G)
    soup = expand_soup(verts, f)
    t0 = time.perf_counter()
    (oracle.solid if solid else oracle.surface)(soup, mn, unit, G)
    return (time.perf_counter() - t0) * 1e3, "port"


def pick_cpu_threads(verts, faces, G, solid):
    """The reference serialises every bit write in a global omp critical (cpu_voxelizer.cpp:11-14), so more
    threads can be slower.  Calibrate on a small sample and keep the faster setting."""
    import oracle
    max_thr = oracle.ref_max_threads() if oracle.have_ref() else (os.cpu_count() or 1)
    n_cal = max(1, min(len(faces), 200000))
    t_one, kind = cpu_reference_run(verts, faces, G, solid, n_cal, 1)
    if max_thr <= 1:
        return 1, max_thr, {"1": t_one}, kind, n_cal
    t_all, _ = cpu_reference_run(verts, faces, G, solid, n_cal, max_thr)
    return (1 if t_one <= t_all else max_thr), max_thr, {"1": round(t_one, 1), str(max_thr): round(t_all, 1)}, kind, n_cal


def bench_config(w, n_tris):
    """The `config` object both arms print (identical, so the driver can tell they ran the same work)."""
    return {"workload": w["desc"], "gridsize": w["G"], "triangles": int(n_tris), "mode": "solid" if w["solid"] else "surface"}


def run_reference(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    wname = args.workload
    w = WORKLOADS[wname]
    verts, faces = load_mesh(wname)
    threads, max_thr, cal, kind, n_cal = pick_cpu_threads(verts, faces, w["G"], w["solid"])
    n_sample = len(faces)              # the whole mesh every step: the same work as the GPU arm
    if args.reference_sample:
        n_sample = min(len(faces), args.reference_sample)
    for _ in range(args.warmup):
        cpu_reference_run(verts, faces, w["G"], w["solid"], n_sample, threads)
    total_ms = 0.0
    for _ in range(args.steps):
        ms, _ = cpu_reference_run(verts, faces, w["G"], w["solid"], n_sample, threads)
        total_ms += ms
    ms_per_step = total_ms / args.steps
    value = n_sample / ms_per_step / 1e3
    sample = ("all %d triangles of the workload per step, whole %d^3 grid; %d thread(s) chosen by calibration %s ms on %d tris "
              "(host has %d hardware threads; the reference's global omp critical makes more threads slower)"
              % (n_sample, w["G"], threads, cal, n_cal, max_thr))
    if n_sample != len(faces):
        sample = "first %d of %d triangles (--reference-sample); " % (n_sample, len(faces)) + sample
    line = {
        "impl": "reference", "metric": METRIC, "value": round(value, 4), "unit": "Mtri/s", "n_gpus": args.gpus, "steps": args.steps,
        "warmup": args.warmup, "ms_per_step": round(ms_per_step, 3), "higher_is_better": True, "scaling": "strong", "vs_baseline": None,
        "dtype": "f32", "data": "synthetic",
        "config": bench_config(w, len(faces)),
        "cpu_baseline": {"value": round(value, 4), "unit": "Mtri/s", "cores": threads, "kind": kind, "sample": sample},
        "e2e": {"value": round(value, 4), "unit": "Mtri/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
    }
    print(json.dumps(line), flush=True)


def run_ours(args):
    import torch
    import torch.distributed as dist
    import cuda_voxelizer_b200 as vb

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    if world != args.gpus and world > 1:
        raise SystemExit("--gpus %d but WORLD_SIZE=%d" % (args.gpus, world))
    torch.cuda.set_device(local_rank)
    vb.init(local_rank)
    if world > 1:
        dist.init_process_group("nccl", device_id=torch.device("cuda", local_rank))

    cpu_group = dist.new_group(backend="gloo") if world > 1 else None

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def cpu_barrier():
        """A wait that leaves the GPUs alone: an NCCL barrier parks a spinning kernel on every waiting rank's device, which is in the
        way when ONE rank drives all devices (voxb200_voxelize_host_multi) while the others wait."""
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier(group=cpu_group)

    wname = args.workload
    w = WORKLOADS[wname]
    G, solid = w["G"], w["solid"]
    verts, faces = load_mesh(wname)
    n_tris = len(faces)
    soup = expand_soup(verts, faces)
    grid = vb.grid_from_verts(verts, G, n_tris)
    region, region_bytes = vb.partition(G, False, rank, world)
    region_arg = None if world == 1 else region
    d_all = torch.from_numpy(soup).cuda()
    table = torch.empty(region_bytes // 4, dtype=torch.int32, device="cuda")
    fn = vb.voxelize_solid if solid else vb.voxelize
    stream = torch.cuda.current_stream()
    # N > 1: the resident input of a rank is the soup ROUTED to its slab (part of the upload path, done once,
    # outside the timed region; the e2e number below pays for it every step).
    import copy
    step_grid, d_tris, routed, route_ms = grid, d_all, n_tris, 0.0
    if world > 1:
        torch.cuda.synchronize()
        warm, _ = vb.route_triangles(grid, d_all, region, solid=solid)
        warm.close()
        torch.cuda.synchronize()
        t_route = time.perf_counter()
        d_tris, routed = vb.route_triangles(grid, d_all, region, solid=solid)      # synchronous (returns the count)
        route_ms = (time.perf_counter() - t_route) * 1e3
        step_grid = copy.copy(grid)
        step_grid.n_triangles = routed
        del d_all
        torch.cuda.empty_cache()

    def one_shot_step():
        fn(step_grid, d_tris, table=table, region=region_arg)          # on torch's current stream (the capture stream during capture)

    def timed(run, n):
        barrier()
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record(stream)
        for _ in range(n):
            run()
        b.record(stream)
        barrier()
        tt = torch.tensor([a.elapsed_time(b) / n], dtype=torch.float64, device="cuda")
        if world > 1:
            dist.all_reduce(tt, op=dist.ReduceOp.MAX)
        return float(tt.item())

    # ---- what one voxelization costs WITHOUT preparation: voxb200_surface / voxb200_solid on the soup in the caller's order ----
    for _ in range(max(args.warmup, 3)):
        one_shot_step()
    n_u = max(1, min(args.steps, 10))
    t_one = timed(one_shot_step, n_u)
    one_shot = {"ms_per_step": round(t_one, 4), "value": round(n_tris / t_one / 1e3, 2), "unit": "Mtri/s",
                "note": "voxb200_%s on the device soup in the caller's triangle order, direct launches, %d steps" % ("solid" if solid else "surface", n_u)}

    # ---- the prepared mesh (surface workloads): created once (allocations), re-prepared inside the timed region ----
    use_mesh = not solid and not args.one_shot
    mesh, mesh_info, prepare_ms = None, None, None
    if use_mesh:
        mesh = vb.Mesh(step_grid, tris=d_tris, region=region_arg)
        mesh_info = mesh.info()

        def step():
            mesh.voxelize(table=table)

        def prepare():
            mesh.update(tris=d_tris)
        for _ in range(2):
            prepare()
        prepare_ms = timed(prepare, 3)
    else:
        step, prepare = one_shot_step, None

    # ---- device-resident timing -----------------------------------------------------------------
    for _ in range(max(args.warmup, 3)):
        step()
    barrier()
    sampler = ClockSampler(local_rank)
    if rank == 0:
        sampler.start()
        time.sleep(0.25)
    # The K steps replay ONE captured CUDA graph of the step (the library only enqueues on the caller's stream, so a whole
    # voxelization is capturable): the kernels are the same, the host-side launch gaps between them are not paid K times.
    graph = None
    if not args.no_graph:
        try:
            torch.cuda.synchronize()
            graph = torch.cuda.CUDAGraph()
            with torch.cuda.graph(graph):
                step()
            graph.replay()
            torch.cuda.synchronize()
        except Exception as exc:      # capture not possible: time the direct calls
            sys.stderr.write("bench: CUDA graph capture failed (%s); timing direct launches\n" % exc)
            graph = None

    def run_steps():
        for _ in range(args.steps):
            if graph is not None:
                graph.replay()
            else:
                step()
    ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    barrier()
    ev0.record(stream)
    if prepare is not None:
        prepare()                     # the caller's order -> tile records: inside the timed region, amortised over the K steps
    run_steps()
    ev1.record(stream)
    barrier()
    elapsed_ms = ev0.elapsed_time(ev1)
    resident_ms = timed(run_steps, 1) / args.steps        # the K steps alone
    # per-kernel times and the launch count come from a separate, untimed pass of direct calls with the library's event ring on
    vb.set_profiling(True)
    launches0 = vb.launch_count()
    n_prof = max(1, min(args.steps, 16))
    for _ in range(n_prof):
        step()
    barrier()
    launches = (vb.launch_count() - launches0) // n_prof * args.steps
    if prepare is not None:           # + the kernels of the one re-preparation inside the timed region
        launches0 = vb.launch_count()
        prepare()
        launches += vb.launch_count() - launches0
    phases = np.array([vb.phase_ms(i) for i in range(n_prof)], np.float64).mean(axis=0)
    vb.set_profiling(False)
    counters = mesh.counters() if mesh is not None else vb.last_counters()
    t = torch.tensor([elapsed_ms], dtype=torch.float64, device="cuda")
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    ms_per_step = float(t.item()) / args.steps
    value = n_tris / ms_per_step / 1e3          # Mtri/s, whole job

    # ---- end to end through the C ABI with host buffers ----------------------------------------
    # The caller's mesh is indexed (vertices + faces, what trimesh holds in the reference's main()): every step uploads
    # it from pinned host memory (12 B/vertex + 12 B/face), expands it on the GPU, voxelizes, and reads the table back.
    pinned_verts = torch.from_numpy(np.ascontiguousarray(verts)).pin_memory()
    pinned_faces = torch.from_numpy(np.ascontiguousarray(faces)).pin_memory()
    pinned_table = torch.empty(region_bytes // 4, dtype=torch.int32).pin_memory()
    e2e_steps = max(1, min(args.steps, 10))
    e2e_dev_ms, e2e_wall_ms, e2e_h2d_bytes = 0.0, 0.0, 0
    if world == 1:
        for _ in range(2):
            vb.voxelize_host_indexed(grid, pinned_verts, pinned_faces, pinned_table, solid=solid, region=region_arg)
        barrier()
        t0 = time.perf_counter()
        for _ in range(e2e_steps):
            _, ms = vb.voxelize_host_indexed(grid, pinned_verts, pinned_faces, pinned_table, solid=solid, region=region_arg)
            e2e_dev_ms += ms[3]
        torch.cuda.synchronize()
        e2e_wall_ms = (time.perf_counter() - t0) * 1e3
        e2e_h2d_bytes = int(pinned_verts.numel() * 4 + pinned_faces.numel() * 4)
        e2e_phases = dict(zip(("h2d", "prepare_voxelize", "readback"), (round(float(x), 3) for x in ms[:3])))
        e2e_readback = vb.last_readback()
        # the same call with the plain dense copy of the table, for comparison
        vb.set_readback_mode("dense")
        vb.voxelize_host_indexed(grid, pinned_verts, pinned_faces, pinned_table, solid=solid, region=region_arg)
        dense_ms = 0.0
        for _ in range(max(1, e2e_steps // 2)):
            _, ms = vb.voxelize_host_indexed(grid, pinned_verts, pinned_faces, pinned_table, solid=solid, region=region_arg)
            dense_ms += ms[3]
        e2e_dense_ms = dense_ms / max(1, e2e_steps // 2)
        vb.set_readback_mode("auto")
        # ... and for a consumer that takes the table's non-zero words instead of the dense table (voxb200_voxelize_host_nonzero)
        e2e_nonzero = None
        if vb.table_bytes(G) // 4 <= 2 ** 32:
            vb.voxelize_host_nonzero(grid, pinned_verts, pinned_faces, solid=solid, region=region_arg)
            nz_ms, nz_pairs = 0.0, None
            for _ in range(max(1, e2e_steps // 2)):
                nz_pairs, ms = vb.voxelize_host_nonzero(grid, pinned_verts, pinned_faces, solid=solid, region=region_arg)
                nz_ms += ms[3]
            nz_ms /= max(1, e2e_steps // 2)
            e2e_nonzero = {"ms_per_step": round(nz_ms, 3), "value": round(n_tris / nz_ms / 1e3, 2), "unit": "Mtri/s", "d2h_bytes_per_step": int(nz_pairs.nbytes),
                           "nonzero_words": int(len(nz_pairs)), "api": "voxb200_voxelize_host_nonzero: same upload and voxelization, output = ascending {word index, bits} pairs in pinned host memory"}
        vb.voxelize_host_indexed(grid, pinned_verts, pinned_faces, pinned_table, solid=solid, region=region_arg)      # the table the checks below read
    e2e_api = ("voxb200_voxelize_host_indexed (pinned host vertices+faces -> H2D -> tile records / expand -> voxelize -> table in pinned host memory: "
               "dense D2H, or non-zero words + host-thread expansion when the table is sparse)")
    if world > 1:
        # N > 1: the upload is sharded too.  Rank r holds 1/N of the soup in pinned memory, uploads only that, routes it
        # on the GPU to the N slabs, swaps triangles in one all-to-all over NVLink, voxelizes its slab, reads it back.
        from cuda_voxelizer_b200 import sharding
        per = (n_tris + world - 1) // world
        chunk = torch.from_numpy(np.ascontiguousarray(soup[rank * per: min(n_tris, (rank + 1) * per)]).reshape(-1)).pin_memory()
        sv = sharding.ShardedHostVoxelizer(grid, solid=solid)
        for _ in range(2):
            sv(chunk, pinned_table)
        barrier()
        t0 = time.perf_counter()
        e2e_dev_ms = 0.0
        for _ in range(e2e_steps):
            e2e_dev_ms += sv(chunk, pinned_table)
        torch.cuda.synchronize()
        e2e_wall_ms = (time.perf_counter() - t0) * 1e3
        e2e_h2d_bytes = int(chunk.numel() * 4)
        e2e_api = "sharding.ShardedHostVoxelizer (pinned 1/N soup -> H2D -> route x N -> NCCL all-to-all -> voxelize -> D2H table slab)"
    # N > 1, through the C ABI proper: ONE process (rank 0) drives all N devices with voxb200_voxelize_host_multi — one host thread per
    # device, 1/N of the mesh bytes over each PCIe link, peer all-gather, every device's slab straight into one pinned host table.
    # The other ranks wait at the barrier (their GPUs are idle while rank 0's threads use them).
    multi = None
    if world > 1:
        cpu_barrier()
        if rank == 0:
            full_table = torch.empty(vb.table_bytes(G) // 4, dtype=torch.int32).pin_memory()
            for _ in range(2):
                vb.voxelize_host_multi(grid, pinned_verts, pinned_faces, full_table, solid=solid, n_devices=world)
            acc = np.zeros(8)
            t0 = time.perf_counter()
            for _ in range(e2e_steps):
                _, tm = vb.voxelize_host_multi(grid, pinned_verts, pinned_faces, full_table, solid=solid, n_devices=world)
                acc += np.array(tm)
            wall = (time.perf_counter() - t0) * 1e3 / e2e_steps
            acc /= e2e_steps
            multi = {"ms_per_step": round(float(acc[5]), 3), "wall_ms_per_step": round(wall, 3),
                     "phases_ms": dict(zip(("h2d_share", "peer_all_gather", "prepare", "voxelize", "d2h_slab"), (round(float(x), 3) for x in acc[:5]))),
                     "h2d_bytes_per_step": int(pinned_verts.numel() * 4 + pinned_faces.numel() * 4), "d2h_bytes_per_step": int(vb.table_bytes(G)),
                     "host_table_bytes": int(vb.table_bytes(G)), "table": full_table}
            # bytes that crossed the links device -> host: every slab of >= 32 MB with few non-zero words goes as {index, value} pairs
            slab_b = vb.table_bytes(G) // world
            nz = int(np.count_nonzero(full_table.numpy()))
            if slab_b >= (32 << 20) and 8 * nz <= vb.table_bytes(G) // 3:
                multi["d2h_bytes_per_step"] = int(8 * nz + 8 * (vb.table_bytes(G) // 8192 + world))
                multi["readback"] = {"mode": "sparse (per slab)", "nonzero_words": nz}
            torch.cuda.set_device(local_rank)
            vb.init(local_rank)
        cpu_barrier()
    # device-event total per step (H2D start -> D2H end), max over ranks; wall kept alongside
    te = torch.tensor([e2e_dev_ms / e2e_steps, e2e_wall_ms / e2e_steps], dtype=torch.float64, device="cuda")
    if world > 1:
        dist.all_reduce(te, op=dist.ReduceOp.MAX)
    e2e_ms, e2e_wall = float(te[0].item()), float(te[1].item())
    e2e_table_check = None
    if world > 1:
        e2e_table_check = bool(torch.equal(pinned_table, table.cpu()))
    clocks = sampler.stop() if rank == 0 else None

    # ---- N > 1: slab gather over NVLink (NCCL all-gather), timed apart; the gathered table is parity-checked ----
    gather_ms, gathered = None, None
    if world > 1:
        from cuda_voxelizer_b200 import sharding
        step()
        barrier()
        g0, g1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        gathered = sharding.gather_table(table)
        barrier()
        g0.record(stream)
        gathered = sharding.gather_table(table)
        g1.record(stream)
        barrier()
        tg = torch.tensor([g0.elapsed_time(g1)], dtype=torch.float64, device="cuda")
        dist.all_reduce(tg, op=dist.ReduceOp.MAX)
        gather_ms = float(tg.item())
        routed_all = torch.tensor([routed], dtype=torch.int64, device="cuda")
        dist.all_reduce(routed_all, op=dist.ReduceOp.SUM)
        routed_total = int(routed_all.item())

    # ---- parity spot check of the timed output (popcount vs the golden generated from the reference) ----
    check = None
    if rank == 0:
        import oracle
        gold_path = os.path.join(ROOT, "tests", "golden", "golden.json")
        key = {"config4": "icosphere:708:1024|2048|surface|linear", "config3": "icosphere:224:512|1024|solid|linear",
               "config2": "bunny|1024|surface|linear"}.get(wname, "")
        gold = json.load(open(gold_path)).get(key)
        host = (gathered if gathered is not None else table).cpu().numpy().view(np.uint32)
        if gold:
            check = {"popcount": oracle.popcount(host), "golden_popcount": gold["popcount"],
                     "fnv1a64_matches_reference_golden": ("%016x" % oracle.fnv1a64(host)) == gold["fnv1a64"]}

    if rank != 0:
        if world > 1:
            dist.destroy_process_group()
        return

    # ---- roofline of the dominant kernel ----------------------------------------------------------
    peak, peak_src = measured_peak_gbs()
    slab_bytes = region_bytes
    tri_bytes = 36 * (routed if world > 1 else n_tris)
    if mesh is not None and mesh_info["tile_schedule"]:
        # phases of voxb200_mesh_voxelize: [-, tile kernel (clears the slab AND voxelizes the small triangles), side path, -]
        names = ["-", "surface_tile_kernel", "side path: surface_tri_kernel + surface_coop_kernel (big triangles)", "-"]
        alg_bytes = [0, tri_bytes + slab_bytes, 0, 0]
    else:
        names = ["zero_kernel", "surface_tri_kernel" if not solid else "solid_tri_kernel",
                 "surface_coop_kernel" if not solid else "solid_coop_kernel", "-" if not solid else "solid_scan_kernel"]
        alg_bytes = [slab_bytes, tri_bytes, 0, 2 * slab_bytes if solid else 0]
        if solid and counters.get("solid_row_lists"):
            # row-list schedule: no zero-fill (phase 0 is the 64-byte counter reset), the fill writes every table byte once
            names[0], names[3] = "counter_reset", "solid_fill_kernel"
            alg_bytes[0], alg_bytes[3] = 0, slab_bytes
    dom = int(np.argmax(phases))
    dom_ms = float(phases[dom])
    achieved = alg_bytes[dom] / (dom_ms * 1e-3) / 1e9 if dom_ms > 0 else 0.0
    step_alg = tri_bytes + slab_bytes
    traffic, traffic_src = None, None
    tpath = os.path.join(ROOT, "profiles", "traffic.json")       # dram__bytes_read+write per launch from the committed ncu capture
    if os.path.exists(tpath) and world == 1:
        try:
            tj = json.load(open(tpath))
            traffic = tj.get(wname, {}).get(names[dom].split(":")[0].strip())
            traffic_src = tj.get("_source") if traffic is not None else None
        except Exception:
            traffic = None
    roofline = {
        "bound": "hbm", "kernel": names[dom], "achieved": round(achieved, 1), "peak": peak, "unit": "GB/s",
        "frac": round(achieved / peak, 4), "traffic": traffic, "traffic_source": traffic_src, "peak_source": peak_src,
        "kernel_ms": round(dom_ms, 4), "algorithmic_bytes": int(alg_bytes[dom]),
        "phases_ms": {n: round(float(p), 4) for n, p in zip(names, phases) if n != "-"},
        "step": {"algorithmic_bytes": int(step_alg), "achieved_gbs": round(step_alg / (ms_per_step * 1e-3) / 1e9, 1),
                 "frac": round(step_alg / (ms_per_step * 1e-3) / 1e9 / peak, 4),
                 "resident_frac": round(step_alg / (resident_ms * 1e-3) / 1e9 / peak, 4)},
    }

    # ---- CPU baseline on this box's host cores (bounded; N=1 only) ---------------------------------
    cpu_baseline = None
    if world == 1 and not args.no_cpu_baseline:
        threads, max_thr, cal, kind, n_cal = pick_cpu_threads(verts, faces, G, solid)
        per_tri_ms = cal[str(threads)] / n_cal
        n_sample = int(min(len(faces), max(1000, 12000.0 / max(per_tri_ms, 1e-9))))
        ms, kind = cpu_reference_run(verts, faces, G, solid, n_sample, threads)
        cpu_baseline = {"value": round(n_sample / ms / 1e3, 4), "unit": "Mtri/s", "cores": threads, "kind": kind,
                        "sample": "first %d of %d triangles, whole %d^3 grid, one run of %.0f ms; threads by calibration %s ms on %d tris (host max %d)"
                                  % (n_sample, len(faces), G, ms, cal, n_cal, max_thr)}

    # ---- second baseline: the reference's OWN GPU kernels (unmodified voxelize.cu / voxelize_solid.cu, sm_100a) on this GPU ----
    ref_gpu = None
    if world == 1 and not args.no_ref_gpu:
        import oracle
        if oracle.have_ref_gpu():
            try:
                rms, rtab = oracle.ref_gpu_run(list(grid.bbox_min), list(grid.bbox_max), G, soup, solid=solid, warmup=2, reps=5, want_table=True)
                step_ms = rms["mean_ms"] + rms["memset_ms"]
                ref_gpu = {"value": round(n_tris / step_ms / 1e3, 2), "unit": "Mtri/s", "ms_per_step": round(step_ms, 4),
                           "kernel_ms": round(rms["mean_ms"], 4), "memset_ms": round(rms["memset_ms"], 4), "popcount": oracle.popcount(rtab),
                           "kind": "the reference's unmodified voxelize.cu / voxelize_solid.cu built for sm_100a (oracle/_ref/libvoxref_gpu.so), triangles and "
                                   "table device-resident, CUDA events around its voxelize() call + the table memset it leaves to the caller; "
                                   "not a parity target (its float path differs from the reference CPU path: see popcount)"}
            except Exception as exc:
                ref_gpu = {"unavailable": str(exc)[:200]}
        else:
            ref_gpu = {"unavailable": "oracle/_ref/libvoxref_gpu.so not built"}

    e2e_obj = {"value": round(n_tris / e2e_ms / 1e3, 2), "unit": "Mtri/s", "h2d_bytes_per_step": e2e_h2d_bytes, "d2h_bytes_per_step": int(slab_bytes),
               "ms_per_step": round(e2e_ms, 3), "wall_ms_per_step": round(e2e_wall, 3), "steps": e2e_steps,
               "api": e2e_api + ", per rank", "slab_matches_device_path": e2e_table_check}
    if world == 1:
        # bytes that crossed the link device -> host: the dense table, or the {index, value} pairs + the per-block prefix
        if e2e_readback["sparse"]:
            e2e_obj["d2h_bytes_per_step"] = int(8 * e2e_readback["nonzero_words"] + 8 * (slab_bytes // 8192 + 1))
        e2e_obj.update({"host_table_bytes": int(slab_bytes), "phases_ms": e2e_phases,
                        "readback": dict(e2e_readback, mode="sparse" if e2e_readback["sparse"] else "dense", host_threads=os.environ.get("VOXB200_HOST_THREADS", "default (half the hardware threads, <= 16)")),
                        "dense_readback": {"ms_per_step": round(e2e_dense_ms, 3), "value": round(n_tris / e2e_dense_ms / 1e3, 2), "d2h_bytes_per_step": int(slab_bytes)},
                        "nonzero_words_output": e2e_nonzero,
                        "host_table_matches_device_table": bool(torch.equal(pinned_table, table.cpu()))})
    if multi is not None:
        # the headline e2e at N > 1 is the C-ABI call a C++ caller makes; the torch.distributed path is kept beside it
        import oracle
        mt = multi.pop("table").numpy().view(np.uint32)
        multi_ok = check is not None and ("%016x" % oracle.fnv1a64(mt)) == json.load(open(os.path.join(ROOT, "tests", "golden", "golden.json"))).get(
            {"config4": "icosphere:708:1024|2048|surface|linear", "config3": "icosphere:224:512|1024|solid|linear", "config2": "bunny|1024|surface|linear"}.get(wname, ""), {}).get("fnv1a64")
        e2e_obj = {"value": round(n_tris / multi["ms_per_step"] / 1e3, 2), "unit": "Mtri/s", "h2d_bytes_per_step": multi["h2d_bytes_per_step"],
                   "d2h_bytes_per_step": multi["d2h_bytes_per_step"], "host_table_bytes": multi["host_table_bytes"], "readback": multi.get("readback", {"mode": "dense"}),
                   "ms_per_step": multi["ms_per_step"], "wall_ms_per_step": multi["wall_ms_per_step"],
                   "steps": e2e_steps, "phases_ms": multi["phases_ms"], "table_matches_reference_golden": multi_ok,
                   "api": "voxb200_voxelize_host_multi: one process (rank 0), one host thread per device; pinned host vertices+faces -> 1/N of the bytes per PCIe link -> "
                          "peer all-gather -> tile records -> voxelize slab -> D2H into one pinned host table (bytes are whole-job totals)",
                   "per_rank_nccl_path": e2e_obj}
    timed_region = ("voxb200_mesh_update (caller's triangle order -> tile records, buffers reused: %.3f ms) + %d x voxb200_mesh_voxelize" % (prepare_ms, args.steps)
                    if mesh is not None else "%d x voxb200_%s on the device soup" % (args.steps, "solid" if solid else "surface"))
    line = {
        "metric": METRIC, "value": round(value, 2), "unit": "Mtri/s", "n_gpus": world, "steps": args.steps, "warmup": max(args.warmup, 3),
        "ms_per_step": round(ms_per_step, 4), "higher_is_better": True, "scaling": "strong", "vs_baseline": None,
        "dtype": "f32", "data": "synthetic",
        "config": bench_config(w, n_tris),
        "details": {"sharding": "z-slab x%d, triangles routed to the slabs their bbox overlaps, no data-path collective" % world,
                    "launch": "one captured CUDA graph of the step, replayed K times" if graph is not None else "direct launches",
                    "timed_region": timed_region,
                    "l2": "inputs larger than L2 (%.0f MB triangles + %.0f MB table slab per GPU vs 126 MB L2)" % (tri_bytes / 1e6, slab_bytes / 1e6),
                    "mesh_plan": mesh_info},
        "resident": {"ms_per_step": round(resident_ms, 4), "value": round(n_tris / resident_ms / 1e3, 2), "unit": "Mtri/s",
                     "note": "the K steps alone, preparation outside"},
        "prepare_ms": None if prepare_ms is None else round(prepare_ms, 4),
        "one_shot": one_shot,
        "e2e": e2e_obj,
        "gpu_launches": int(launches), "clocks": clocks, "roofline": roofline, "cpu_baseline": cpu_baseline, "ref_gpu_baseline": ref_gpu,
        "counters": counters, "parity": check,
    }
    if world > 1:
        line["gather"] = {"ms": round(gather_ms, 4), "bytes_per_rank_received": int(region_bytes * (world - 1)), "how": "NCCL all_gather_into_tensor of the slabs, outside the timed step",
                          "triangles_routed_total": routed_total, "duplication": round(routed_total / n_tris, 4),
                          "route_ms_outside_step": round(route_ms, 3),
                          "note": "the resident input of a rank is the soup routed to its slab (done once at upload, wall-clock above, incl. its cudaMalloc); the e2e number pays for routing every step"}
    print(json.dumps(line), flush=True)
    if mesh is not None:
        mesh.close()
    if world > 1:
        dist.destroy_process_group()


def main():
    # Exactly ONE line goes to stdout (the JSON): anything libraries print (NCCL's version banner, torchrun notices)
    # is sent to stderr by pointing fd 1 there and keeping the real stdout aside for the final line.
    real_stdout = os.fdopen(os.dup(1), "w")
    os.dup2(2, 1)
    sys.stdout = real_stdout
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--workload", default="config4", choices=sorted(WORKLOADS))
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-ref-gpu", action="store_true", help="skip the reference-GPU-kernel baseline leg")
    ap.add_argument("--reference-sample", type=int, default=0, help="--impl reference: voxelize only the first N triangles per step (default: the whole mesh)")
    ap.add_argument("--one-shot", action="store_true", help="time voxb200_surface on the device soup instead of the prepared-mesh path")
    ap.add_argument("--no-graph", action="store_true", help="time direct launches instead of replaying a captured CUDA graph of the step")
    args = ap.parse_args()
    if args.impl == "reference":
        run_reference(args)
    else:
        run_ours(args)


if __name__ == "__main__":
    main()
