#!/usr/bin/env python
"""Benchmark of the dense-reconstruction hot path (BASELINE.json metric: frames/s TSDF-integrate + ICP at
640x480, 5 mm voxels; HBM GB/s vs roofline).

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference]

A step = one frame of the stream through the hot path.  See DESIGN.md "Measurement" for the definitions of
every number printed here.  One JSON line on stdout (rank 0)."""
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
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

METRIC = "frames/sec TSDF-integrate+ICP @640x480, 5mm voxel"
UNIT = "frames/s"
VOXEL = 0.005


def parse_args():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=200)
    ap.add_argument("--warmup", type=int, default=10)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-odometry", action="store_true", help="skip the secondary DenseTracking measurement (config 3)")
    ap.add_argument("--no-partitioned", action="store_true", help="N > 1: skip the secondary one-stream-over-all-ranks measurement")
    return ap.parse_args()


def measured_traffic(kernel: str):
    """dram__bytes_read.sum + dram__bytes_write.sum of one launch, from the committed `ncu --set full` summary of this round"""
    import glob
    import re
    for path in sorted(glob.glob(os.path.join(ROOT, "profiles", f"r*_{kernel}_ncu_full.md")), reverse=True):
        tot = 0.0
        for m in re.finditer(r"dram__bytes_(?:read|write)\.sum \| ([0-9.]+) \| (\w+)", open(path).read()):
            tot += float(m.group(1)) * {"byte": 1.0, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}.get(m.group(2), 1.0)
        if tot:
            return int(tot), os.path.relpath(path, ROOT)
    return None, None


def peaks():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        return json.load(open(p)), "measured"
    return {"hbm_gbs": 6650.0, "sm_max_mhz": 1965.0}, "fallback"


class ClockSampler:
    """SM clock and throttle reasons sampled DURING the timed region: NVML polled from a thread, every 2 ms at first (the driver times
    as few as 20 steps, 16 ms -- too short for an nvidia-smi process to answer once); nvidia-smi -lms as the fallback."""
    Q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
         "clocks_event_reasons.sw_power_cap")
    REASONS = ((0x8, "hw_slowdown"), (0x40, "hw_thermal_slowdown"), (0x20, "sw_thermal_slowdown"), (0x4, "sw_power_cap"))

    def __init__(self, index: int):
        self.index = index
        self.proc = None
        self.lines = []
        self.nvml = None          # (module, handle) when NVML answers
        self.samples = []         # (sm MHz, reasons bitmask)
        self.max_mhz = None
        self.halt = threading.Event()
        self.thread = None

    def _nvml_handle(self):
        import pynvml
        import torch
        pynvml.nvmlInit()
        try:
            uuid = "GPU-" + str(torch.cuda.get_device_properties(self.index).uuid)
            try:
                return pynvml, pynvml.nvmlDeviceGetHandleByUUID(uuid)
            except TypeError:
                return pynvml, pynvml.nvmlDeviceGetHandleByUUID(uuid.encode())
        except Exception:
            vis = os.environ.get("CUDA_VISIBLE_DEVICES", "")
            ids = [v.strip() for v in vis.split(",") if v.strip()]
            phys = int(ids[self.index]) if self.index < len(ids) and ids[self.index].isdigit() else self.index
            return pynvml, pynvml.nvmlDeviceGetHandleByIndex(phys)

    def _poll(self):
        nv, h = self.nvml
        while not self.halt.is_set():
            try:
                mhz = nv.nvmlDeviceGetClockInfo(h, nv.NVML_CLOCK_SM)
                try:
                    why = nv.nvmlDeviceGetCurrentClocksEventReasons(h)
                except Exception:
                    why = nv.nvmlDeviceGetCurrentClocksThrottleReasons(h)
                self.samples.append((float(mhz), int(why)))
            except Exception:
                break
            self.halt.wait(0.002 if len(self.samples) < 25 else 0.02)  # dense over a short region, sparse over a long one

    def start(self):
        try:
            self.nvml = self._nvml_handle()
            self.max_mhz = float(self.nvml[0].nvmlDeviceGetMaxClockInfo(self.nvml[1], self.nvml[0].NVML_CLOCK_SM))
            self.thread = threading.Thread(target=self._poll, daemon=True)
            self.thread.start()
            return
        except Exception:
            self.nvml = None
        try:
            self.proc = subprocess.Popen(["nvidia-smi", f"--id={self.index}", f"--query-gpu={self.Q}",
                                          "--format=csv,noheader,nounits", "-lms", "100"],
                                         stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.thread = threading.Thread(target=self._read, daemon=True)
            self.thread.start()
        except OSError:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.lines.append(line.strip())

    def stop(self):
        if self.nvml:
            self.halt.set()
            self.thread.join(timeout=2)
            sm = [m for m, _ in self.samples]
            bits = 0
            for _, w in self.samples:
                bits |= w
            return {"sm_mhz": float(np.median(sm)) if sm else None, "sm_max_mhz": self.max_mhz, "samples": len(sm),
                    "reasons": sorted(n for b, n in self.REASONS if bits & b), "source": "NVML polled inside the timed region (every 2 ms for the first 25 samples, then every 20 ms)"}
        if not self.proc:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["neither NVML nor nvidia-smi available"]}
        self.proc.terminate()
        try:
            self.proc.wait(timeout=2)
        except subprocess.TimeoutExpired:
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
            for n, v in zip(names, f[5:9]):
                if v.lower().startswith("active"):
                    reasons.add(n)
        return {"sm_mhz": float(np.median(sm)) if sm else None, "sm_max_mhz": max(mx) if mx else None,
                "samples": len(sm), "reasons": sorted(reasons), "source": "nvidia-smi -lms 100"}


N_TRAJ = 9  # frames 0..8 of the S2 trajectory; step s registers frame k+1 against frame k, k = s mod 8
ICP_ITERS = 30
ICP_THRESHOLD = 0.05
WORKLOAD = ("config2: S2 room stream 640x480 (u16 depth, 307,200 points/frame), per frame point-to-plane ICP "
            "(frame k+1 -> frame k, analytic target normals, 30 iterations, threshold 0.05 m) then TSDF integration "
            "of the frame at 5 mm voxels with the ICP pose")
WORKLOAD4 = ("config4: ONE S1 stream 1280x960 (f32 depth) fused by all ranks into a 2 mm volume partitioned by cube ownership "
             "(slabs of 4 cubes along x, round-robin over the ranks; identity poses, truncation 0.1 m); every rank sees every "
             "frame and updates the cubes it owns, no data-path collective per frame; boundary-cube exchange + Marching Cubes "
             "at the end")


def config_of(world: int) -> dict:
    """The `config` object, identical in both arms (only static facts of the workload; measured details go to `details`)."""
    if world > 1:
        return {"workload": WORKLOAD4, "voxel_m": 0.002, "storage": "f32 20 B/voxel", "image": "1280x960", "stream_frames": 4,
                "l2": "voxel working set of a frame (cubes x 10 KB) exceeds the 126 MB L2; no explicit flush",
                "sharding": f"one stream, volume partitioned over {world} ranks by cube ownership (strong scaling)",
                "single_gpu_workload": WORKLOAD}
    return {"workload": WORKLOAD, "voxel_m": VOXEL, "storage": "f32 20 B/voxel", "image": "640x480", "icp_points": 307200,
            "icp_iterations": ICP_ITERS, "icp_threshold_m": ICP_THRESHOLD, "stream_frames": N_TRAJ,
            "l2": "voxel working set of a frame (cubes x 10 KB) exceeds the 126 MB L2; no explicit flush",
            "sharding": "single GPU"}


def make_stream(cam, rank: int):
    """Per-rank synthetic stream: every rank looks at its own copy of the room (its own sub-volume)."""
    from onepiece_b200 import scenes
    frames = []
    for k in range(N_TRAJ):
        d, c, T, n = scenes.room(cam, k + 3 * rank, with_normals=True)
        cloud = scenes.backproject(d, cam)
        frames.append(dict(depth=d, bgr=c, pose=T.astype(np.float64), cloud=cloud,
                           normals=np.ascontiguousarray(n.reshape(-1, 3)[(d > 0).reshape(-1)])))
    return frames


def make_stream4(cam4, n: int = 4):
    from onepiece_b200 import scenes
    return [scenes.wavy_wall(cam4, k) for k in range(n)]


def pose_delta(A, B):
    A, B = np.asarray(A, np.float64), np.asarray(B, np.float64)
    R = A[:3, :3].T @ B[:3, :3]
    ang = np.linalg.norm([R[2, 1] - R[1, 2], R[0, 2] - R[2, 0], R[1, 0] - R[0, 1]]) / 2
    return float(np.linalg.norm(A[:3, 3] - B[:3, 3])), float(ang)


# ----------------------------------------------------------------------------------------------------------
# reference arm / CPU baseline: the reference's own CPU implementation (oracle/_ref when it was compiled in the
# build container, else the plain-C port), on the host cores of this box
# ----------------------------------------------------------------------------------------------------------
def use_all_host_threads() -> int:
    """torchrun exports OMP_NUM_THREADS=1 to its workers; the reference's OpenMP loops (ICP.cpp:64,184) are to run on all host
    cores like a stand-alone run.  Returns the thread count OpenMP will use."""
    cores = os.cpu_count() or 1
    os.environ["OMP_NUM_THREADS"] = str(cores)
    try:
        gomp = C.CDLL("libgomp.so.1")
        gomp.omp_set_num_threads(cores)
        return int(gomp.omp_get_max_threads())
    except OSError:
        return cores


def cpu_reference_fps(steps: int, warmup: int, budget_s: float | None = None, world: int = 1):
    """K steps of the workload on the reference's own code.  world > 1: the config-4 workload (integration only)."""
    from onepiece_b200 import scenes
    from oracle import oracleapi, refapi
    threads = use_all_host_threads()
    use_ref = refapi.available("f32")
    kind = "reference" if use_ref else "port"
    poses = {}
    if world > 1:
        cam = scenes.Camera().scaled(2)
        frames4 = make_stream4(cam)
        vol = refapi.RefVolume(cam, 0.002) if use_ref else oracleapi.OracleVolume(cam, 0.002)
        I = np.eye(4, dtype=np.float32)

        def step(s):
            d, c = frames4[s % len(frames4)]
            vol.integrate(d, c, I)
        what = "CubeHandler::IntegrateImage (single-threaded in the reference) of 1280x960 frames into one 2 mm volume"
    else:
        cam = scenes.Camera()
        frames = make_stream(cam, 0)
        vol = refapi.RefVolume(cam, VOXEL) if use_ref else oracleapi.OracleVolume(cam, VOXEL)

        def step(s):
            k = s % (N_TRAJ - 1)
            a, b = frames[k], frames[k + 1]
            if use_ref:
                r = refapi.icp(b["cloud"], a["cloud"], a["normals"], np.eye(4), ICP_ITERS, ICP_THRESHOLD, "f32")
            else:
                r = oracleapi.icp(b["cloud"], a["cloud"], a["normals"], np.eye(4), ICP_ITERS, ICP_THRESHOLD)
            poses[k] = np.asarray(r["T"], np.float64)
            pose = (a["pose"] @ r["T"]).astype(np.float32)
            vol.integrate(b["depth"], b["bgr"], pose)
        what = (f"registration::PointToPlane (OpenMP nearest-neighbour search on {threads} threads, the rest single-threaded) + "
                f"CubeHandler::IntegrateImage (single-threaded) on the bench workload")

    for s in range(warmup):
        step(s)
    n = 0
    t0 = time.perf_counter()
    t_used = 0.0
    while n < steps and (budget_s is None or t_used < budget_s):
        step(warmup + n)
        n += 1
        t_used = time.perf_counter() - t0
    fps = n / t_used
    return {"value": fps, "unit": UNIT, "cores": threads, "kind": kind,
            "sample": f"{n} frames of {what}, after {warmup} warm-up frames; {t_used:.1f} s"}, n, t_used, poses


def run_reference(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    world = int(os.environ.get("WORLD_SIZE", str(args.gpus)))
    cb, n, t, _ = cpu_reference_fps(args.steps, args.warmup, None, world)
    out = {"metric": METRIC, "value": cb["value"], "unit": UNIT, "n_gpus": args.gpus, "steps": n, "warmup": args.warmup,
           "ms_per_step": 1e3 * t / n, "higher_is_better": True, "scaling": "strong" if world > 1 else "weak", "vs_baseline": None,
           "dtype": "f32", "data": "synthetic", "impl": "reference", "config": config_of(world), "cpu_baseline": cb,
           "e2e": {"value": cb["value"], "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
           "gpu_launches": 0}
    print(json.dumps(out), flush=True)


# ----------------------------------------------------------------------------------------------------------
# secondary measurement (BASELINE.json config 3): Odometry::DenseTracking chained over the stream as DenseSlam::UpdateFrame
# does (example/DenseFusion/DenseSlam.cpp:22-31) -- every step uploads ONE new 640x480 RGB-D frame from pinned host memory,
# pre-processes it on the device and tracks it against the previous frame (3 levels, {16, 8, 4} iterations, hybrid term);
# pose, rmse, flag and counters come back every step.  Reported beside the headline metric, not part of it.
# ----------------------------------------------------------------------------------------------------------
def bench_dense_odometry(frames, cam, device, steps, with_cpu):
    import torch

    from onepiece_b200.odometry import Odometry
    odo = Odometry(cam, device=device)
    odo.set_profiling(True)
    H = [(torch.from_numpy(f["bgr"]).pin_memory(), torch.from_numpy(f["depth"]).pin_memory()) for f in frames]
    I = np.eye(4)

    def run(n, pairs):
        prev = odo.Frame(H[0][0].numpy(), H[0][1].numpy())
        dev_ms = []
        t0 = time.perf_counter()
        for s in range(n):
            k = (s + 1) % N_TRAJ
            cur = odo.Frame(H[k][0].numpy(), H[k][1].numpy(), k)
            odo.DenseTracking(cur, prev, I, 0, want_correspondences=pairs)
            dev_ms.append(odo.last_tracking_ms())
            prev.Release()
            prev = cur
        torch.cuda.synchronize()
        return (time.perf_counter() - t0) / n, float(np.median(dev_ms)), odo.last_solve_tail_us

    run(5, False)
    wall, dev_ms, tail_us = run(steps, False)
    wall_pairs, _, _ = run(max(steps // 4, 5), True)
    npx = cam.width * cam.height
    out = {"metric": "frames/s Odometry::DenseTracking @640x480, 3 levels {16,8,4}, hybrid term", "unit": "frames/s",
           "value": 1e3 / dev_ms, "device_ms_per_frame": dev_ms,
           "value_note": "CUDA events around pre-processing of the new frame + NormalizeIntensity + 28 solver iterations + result assembly",
           "e2e": {"value": 1.0 / wall_pairs, "unit": "frames/s", "h2d_bytes_per_step": npx * 5,
                   "d2h_bytes_per_step": "4800 + 40 bytes per correspondence (pixel pairs + 3-D point pairs, ~12 MB): everything "
                                         "odometry::DenseTrackingResult holds",
                   "clock": "host wall clock over synchronous RGBDFrame upload + DenseTracking calls, pinned host images"},
           "e2e_pose_only": {"value": 1.0 / wall, "unit": "frames/s", "d2h_bytes_per_step": 4800,
                             "what": "pose, rmse, flag and counters only (DenseSlam::UpdateFrame uses nothing else of the result)"},
           "solve_tail_us_per_iteration": tail_us, "steps": steps}
    if with_cpu:
        from oracle import oracleapi
        t0 = time.perf_counter()
        n = 0
        prev = oracleapi.OracleFrame(frames[0]["bgr"], frames[0]["depth"])
        while n < 12 and time.perf_counter() - t0 < 12.0:
            k = (n + 1) % N_TRAJ
            cur = oracleapi.OracleFrame(frames[k]["bgr"], frames[k]["depth"])
            oracleapi.dense_tracking_frames(cur, prev, cam, I, 0)
            prev = cur
            n += 1
        dt = time.perf_counter() - t0
        out["cpu_baseline"] = {"value": n / dt, "unit": "frames/s", "cores": 1, "kind": "port",
                               "sample": f"{n} chained frame pairs through the plain-C restatement of DenseTracking (the reference's "
                                         f"Odometry.cpp needs OpenCV and cannot be compiled here; it is single-threaded too); {dt:.1f} s"}
    odo.close()
    return out


# ----------------------------------------------------------------------------------------------------------
# BASELINE.json config 4 (the north star's multi-GPU case): ONE 1280x960 stream fused into a 2 mm volume that is partitioned by
# cube ownership over all ranks.  Every rank sees every frame, selects and updates only the cubes of its slabs;
# no data-path collective per frame.  At the end: boundary-cube exchange + Marching Cubes per rank; the vertex counts must add
# up to the unpartitioned volume's.  world == 1 runs the same stream unpartitioned (the strong-scaling reference).
# ----------------------------------------------------------------------------------------------------------
VOXEL4 = 0.002
SLAB4 = 4   # cubes per slab: 8xB200 measured 6,015 / 6,648 / 6,600 frames/s with slabs of 8 / 4 / 2 (load balance against halo size)


def bench_config4(local, rank, world, steps, warmup, reference_on_rank0: bool):
    import torch
    import torch.distributed as dist

    from onepiece_b200 import capi, fusion, scenes
    from onepiece_b200.volume import CubeHandler
    cam = scenes.Camera().scaled(2)
    frames = make_stream4(cam)
    npx = cam.width * cam.height
    I16 = np.ascontiguousarray(np.eye(4, dtype=np.float32)).reshape(16)
    I4 = np.eye(4, dtype=np.float32)
    stream = torch.cuda.Stream()
    D = [(torch.from_numpy(d).cuda(), torch.from_numpy(c).cuda()) for d, c in frames]
    H = [(torch.from_numpy(d).pin_memory(), torch.from_numpy(c).pin_memory()) for d, c in frames]

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def run(shard, n_world, collective: bool):
        """-> dict of measurements of one volume (partitioned over n_world ranks, or whole)"""
        vol = CubeHandler(cam, VOXEL4, max_cubes=max((1 << 18) // n_world * 2, 1 << 14), device=local, shard=shard, stream=stream.cuda_stream)
        # first pass over the stream through the synchronous call: the pool sizes itself (it grows when full)
        for d, c in frames:
            vol.IntegrateImage(d, c, I4)

        banded = collective and n_world > 1
        ring_maps = fusion.attach_frame_ring(vol, rank, n_world, local) if banded else None
        lo, hi = fusion.shard_range(cam.height, rank, n_world) if banded else (0, cam.height)

        def timed(host: bool, bands: bool = False):
            B = H if host else D

            def one(s):
                d, c = B[s % len(B)]
                if bands:   # this rank uploads rows [lo, hi) only; the bands meet in every rank's frame ring over NVLink
                    vol.IntegrateRowsAsync(d.data_ptr() + lo * cam.width * 4, capi.OPB_DEPTH_F32, c.data_ptr() + lo * cam.width * 3, lo, hi - lo, I16)
                else:
                    (vol.IntegrateImageAsync if host else vol.IntegrateImageDevice)(d.data_ptr(), capi.OPB_DEPTH_F32, c.data_ptr(), I16)
            for s in range(warmup):
                one(s)
            vol.Synchronize()
            if collective:
                barrier()
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record(stream)
            for s in range(steps):
                one(warmup + s)
            e1.record(stream)
            vol.Synchronize()
            if collective:
                barrier()
            ms = e0.elapsed_time(e1)
            if collective and world > 1:
                t = torch.tensor([ms], device="cuda", dtype=torch.float64)
                dist.all_reduce(t, op=dist.ReduceOp.MAX)
                ms = float(t[0].item())
            return ms

        ms_dev = timed(False)
        ms_whole = timed(True)
        ms_e2e = timed(True, True) if banded else timed(True)   # (the same number of frames either way: the meshes are compared)
        if banded:
            vol.FrameRingStatus()
        # per-kernel times of this rank's share
        vol.SetProfiling(True)
        vol.ProfileRead(reset=True)
        upd = 0
        n_prof = min(steps, 20)
        for s in range(n_prof):
            d, c = D[s % len(D)]
            vol.IntegrateImageDevice(d.data_ptr(), capi.OPB_DEPTH_F32, c.data_ptr(), I16)
            st = vol.FrameStats()
            upd += st.updated_voxels
        sel_ms, int_ms, nprof = vol.ProfileRead(reset=True)
        vol.SetProfiling(False)
        st = vol.FrameStats()
        out = {"ms_per_frame": ms_dev / steps, "e2e_ms_per_frame": ms_e2e / steps, "e2e_whole_frame_per_rank_ms": ms_whole / steps,
               "cubes": vol.NumCubes(), "frame_cubes": st.frame_cubes,
               "updated_voxels_per_frame": upd / max(n_prof, 1), "select_ms": sel_ms / max(nprof, 1), "integrate_ms": int_ms / max(nprof, 1),
               "pool_grew": st.overflow}
        # mesh: boundary cubes from the owner of the next slab, then Marching Cubes (count only).  The exchange is two kernel
        # launches per rank over peer memory (export into the neighbour's box, import from one's own); timed with CUDA events
        # on the volume's stream between barriers, the second of two exchanges (the first maps the peer pages)
        n_ghost, halo_ms = 0, 0.0
        if collective and world > 1:
            maps = fusion.attach_halo_peers(vol, rank, world, max(2 * vol.NumCubes(), 1024), local)
            for rep in range(2):
                vol.HaloClear()
                barrier()
                e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                e0.record(stream)
                vol.HaloExchangeBegin()
                e1.record(stream)
                n_sent, n_ghost = vol.HaloExchangeEnd()
                halo_ms = e0.elapsed_time(e1)
            barrier()
        nv, nt = vol.CountMesh()
        out.update(halo_exchange_ms=halo_ms, boundary_cubes_imported=int(n_ghost), mesh_vertices=int(nv))
        if collective and world > 1:
            fusion.detach_halo_peers(vol, maps, local)
        if ring_maps is not None:
            fusion.detach_frame_ring(vol, ring_maps, local)
        vol.close()
        return out

    mine = run((rank, world, 0, SLAB4) if world > 1 else None, world, True)
    res = {"partitioned": mine}
    if world > 1:
        tot = torch.tensor([mine["cubes"], mine["mesh_vertices"], mine["boundary_cubes_imported"], int(mine["updated_voxels_per_frame"])],
                           device="cuda", dtype=torch.int64)
        dist.all_reduce(tot)
        mx = torch.tensor([mine["integrate_ms"], mine["select_ms"], mine["halo_exchange_ms"]], device="cuda", dtype=torch.float64)
        dist.all_reduce(mx, op=dist.ReduceOp.MAX)
        res.update(cubes_total=int(tot[0]), mesh_vertices_total=int(tot[1]), boundary_cubes_exchanged=int(tot[2]),
                   updated_voxels_per_frame_total=int(tot[3]), integrate_ms_max_over_ranks=float(mx[0]), select_ms_max_over_ranks=float(mx[1]),
                   halo_exchange_ms_max_over_ranks=float(mx[2]))
        if reference_on_rank0:
            # the same stream, unpartitioned, on rank 0's GPU alone (the other ranks wait): the strong-scaling reference
            if rank == 0:
                res["single_gpu"] = run(None, 1, False)
            barrier()
    return res


# ----------------------------------------------------------------------------------------------------------
# BASELINE.json config 5 at the bench's image size: the DenseFusion loop over ALL ranks -- per frame a point-to-plane ICP whose
# source points are split across the ranks (the 6x6 packet all-reduced over peer memory inside the persistent solver kernel),
# then the frame integrated into the volume partitioned by cube ownership; at the end boundary-cube exchange + Marching Cubes.
# Host buffers, wall clock between barriers.  The solvers are latency-bound at 640x480 (SURVEY.md 8e expected little or
# negative gain from splitting them): this number says what the collectives cost, it is not the scaling headline.
# ----------------------------------------------------------------------------------------------------------
def bench_config5(cam, local, rank, world, steps):
    import torch
    import torch.distributed as dist

    from onepiece_b200 import fusion, registration as reg
    frames = make_stream(cam, 0)   # every rank looks at the SAME stream here
    sp = fusion.SplitICP(local)
    sh = fusion.ShardedCubeHandler(cam, VOXEL, max_cubes=1 << 16, axis=0, slab=8, device_index=local)
    par = reg.ICPParameter(ICP_ITERS, ICP_THRESHOLD, 1.0)

    def pinned(a):
        return torch.from_numpy(np.ascontiguousarray(a)).pin_memory().numpy()

    clouds = []
    for f in frames:
        pc = reg.PointCloud(f["cloud"], f["normals"])
        pc.points, pc.normals = pinned(pc.points), pinned(pc.normals)
        f["depth"], f["bgr"] = pinned(f["depth"]), pinned(f["bgr"])
        clouds.append(pc)

    def step(s):
        k = s % (N_TRAJ - 1)
        r = sp.PointToPlane(clouds[k + 1], clouds[k], np.eye(4), par, gather_pairs=False)
        pose = (frames[k]["pose"] @ r.T.astype(np.float64)).astype(np.float32)
        sh.IntegrateImage(frames[k + 1]["depth"], frames[k + 1]["bgr"], pose)
        return r

    for s in range(3):
        step(s)
    dist.barrier(); torch.cuda.synchronize()
    t0 = time.perf_counter()
    for s in range(steps):
        r = step(3 + s)
    torch.cuda.synchronize(); dist.barrier()
    dt = time.perf_counter() - t0
    t1 = time.perf_counter()
    n_ghost = fusion.exchange_halo(sh.volume, rank, world, sh.device)
    torch.cuda.synchronize(); dist.barrier()
    t2 = time.perf_counter()
    nv, nt = sh.volume.CountMesh()
    tot = torch.tensor([sh.volume.NumCubes(), n_ghost, nv], device="cuda", dtype=torch.int64)
    dist.all_reduce(tot)
    Ts = [None] * world
    dist.all_gather_object(Ts, r.T.tobytes())
    sp.close()
    sh.close()
    return {"what": "config5 at 640x480: one stream fused by all ranks, split point-to-plane ICP (peer-memory packet exchange inside "
                    "the persistent solver kernel) + partitioned integration per frame, host buffers; then halo exchange + Marching Cubes "
                    "count.  Dense odometry is not split across ranks (its correspondence pass is order-dependent; at 1 ms per frame pair "
                    "a split cannot pay for the exchange), so a DenseTracking-based pipeline scales only in its integration leg",
            "frames_per_s": steps / dt, "ms_per_frame": 1e3 * dt / steps, "frames": steps, "cubes_total": int(tot[0]),
            "boundary_cubes_exchanged": int(tot[1]), "halo_exchange_ms": 1e3 * (t2 - t1), "mesh_vertices_total": int(tot[2]),
            "pose_identical_on_all_ranks": bool(all(t == Ts[0] for t in Ts))}


def bench_packed16(cam, frames, local, steps, peak_gbs):
    """The voxel update on OPB_STORAGE_PACKED16 volumes (8-byte voxels: half sdf, half weight, rgb8 -- north_star's "fp16 voxel
    writes") over the same frames with their true poses, next to the float kernel's roofline and with its deviation from the
    float volume.  A secondary: the packed mode is outside the parity contract, the headline stays on 20-byte float voxels."""
    import torch

    from onepiece_b200 import capi
    from onepiece_b200.volume import CubeHandler
    vols = {name: CubeHandler(cam, VOXEL, max_cubes=1 << 17, device=local, storage=st)
            for name, st in (("f32", capi.OPB_STORAGE_F32), ("packed16", capi.OPB_STORAGE_PACKED16))}
    D = [(torch.from_numpy(f["depth"]).cuda(), torch.from_numpy(f["bgr"]).cuda(),
          np.ascontiguousarray(f["pose"].astype(np.float32).T).reshape(16)) for f in frames]
    out = {}
    for name, vol in vols.items():
        for d, c, pose in D:   # first pass: allocation
            vol.IntegrateImageDevice(d.data_ptr(), capi.OPB_DEPTH_U16, c.data_ptr(), pose)
        vol.Synchronize()
        vol.SetProfiling(True)
        vol.ProfileRead(reset=True)
        upd = 0
        for s in range(steps):
            d, c, pose = D[s % len(D)]
            vol.IntegrateImageDevice(d.data_ptr(), capi.OPB_DEPTH_U16, c.data_ptr(), pose)
            upd += vol.FrameStats().updated_voxels
        sel_ms, int_ms, n = vol.ProfileRead(reset=True)
        vol.SetProfiling(False)
        bytes_per_voxel = 20 if name == "f32" else 8
        alg = upd / steps * 2 * bytes_per_voxel + cam.width * cam.height * 5
        k_ms = int_ms / max(n, 1)
        out[name] = {"bytes_per_voxel": bytes_per_voxel, "kernel_ms": k_ms, "updated_voxels_per_frame": int(upd / steps),
                     "algorithmic_bytes_per_launch": int(alg), "achieved_gbs": alg / (k_ms * 1e-3) / 1e9,
                     "frac_of_hbm_peak": alg / (k_ms * 1e-3) / 1e9 / peak_gbs}
    fi, fv = vols["f32"].GetCubeMap()
    pi, pv = vols["packed16"].GetCubeMap()
    same = bool(np.array_equal(fi, pi))
    err = {"cube_set_identical": same}
    if same:
        seen = fv[..., 1] > 0
        err.update(max_abs_sdf_error_m=float(np.abs(fv[..., 0] - pv[..., 0])[seen].max()),
                   max_abs_colour_error=float(np.abs(fv[..., 2:] - pv[..., 2:])[seen].max()),
                   weights_identical=bool(np.array_equal(fv[..., 1], pv[..., 1])))
    nv_f, nv_p = vols["f32"].CountMesh()[0], vols["packed16"].CountMesh()[0]
    err.update(mesh_vertices_f32=int(nv_f), mesh_vertices_packed16=int(nv_p), frames_integrated=len(D) + steps)
    for v in vols.values():
        v.close()
    return {"what": "voxel update alone, frames and poses of the bench stream, 5 mm: 20-byte float voxels (the parity path, kernel "
                    "integrate_pipelined_kernel) against 8-byte packed voxels (integrate_packed_kernel); gate of the packed mode in "
                    "tests/test_packed_gpu.py",
            "kernel_speedup": out["f32"]["kernel_ms"] / out["packed16"]["kernel_ms"], **out, "deviation_from_f32": err}


def config4_line(c4, world, steps, peak_gbs):
    """the numbers of bench_config4 as reported in the JSON line"""
    p = c4["partitioned"]
    npx = 1280 * 960
    upd = c4.get("updated_voxels_per_frame_total", int(p["updated_voxels_per_frame"]))
    out = {"what": WORKLOAD4, "n_gpus": world, "frames_per_s": 1e3 / p["ms_per_frame"], "ms_per_frame": p["ms_per_frame"],
           "e2e_frames_per_s": 1e3 / p["e2e_ms_per_frame"],
           "e2e_note": ("every rank uploads only its band of rows of the frame from pinned host memory (8.6 MB per frame in total over N PCIe links); a "
                        "scatter kernel stores the band into every peer's frame ring over NVLink, the frame is integrated when all bands have "
                        "landed (opb_volume_integrate_rows_async)") if world > 1 else "the frame is uploaded from pinned host memory, double-buffered against the kernels",
           "e2e_whole_frame_per_rank_frames_per_s": 1e3 / p["e2e_whole_frame_per_rank_ms"],
           "cubes_total": c4.get("cubes_total", p["cubes"]), "mesh_vertices_total": c4.get("mesh_vertices_total", p["mesh_vertices"]),
           "boundary_cubes_exchanged": c4.get("boundary_cubes_exchanged", 0),
           "halo_exchange_ms": c4.get("halo_exchange_ms_max_over_ranks", p["halo_exchange_ms"]),
           "rank0": {k: p[k] for k in ("cubes", "frame_cubes", "select_ms", "integrate_ms", "pool_grew")},
           "integrate_ms_max_over_ranks": c4.get("integrate_ms_max_over_ranks", p["integrate_ms"]),
           "select_ms_max_over_ranks": c4.get("select_ms_max_over_ranks", p["select_ms"])}
    # roofline of the voxel update on the slowest rank's share (algorithmic bytes of the whole frame / N ranks on average)
    alg = upd * 40 + world * npx * (4 + 3)
    out["algorithmic_bytes_per_frame_all_ranks"] = int(alg)
    k_ms = out["integrate_ms_max_over_ranks"]
    if k_ms > 0:
        ach = alg / world / (k_ms * 1e-3) / 1e9
        out["roofline_per_gpu"] = {"bound": "hbm", "achieved": ach, "peak": peak_gbs, "unit": "GB/s", "frac": ach / peak_gbs,
                                   "note": "mean algorithmic bytes per rank / slowest rank's integrate kernel time"}
    if "single_gpu" in c4:
        g = c4["single_gpu"]
        out["single_gpu_same_workload"] = {"frames_per_s": 1e3 / g["ms_per_frame"], "ms_per_frame": g["ms_per_frame"],
                                           "e2e_frames_per_s": 1e3 / g["e2e_ms_per_frame"], "cubes": g["cubes"], "mesh_vertices": g["mesh_vertices"],
                                           "select_ms": g["select_ms"], "integrate_ms": g["integrate_ms"],
                                           "how": "same stream, unpartitioned volume, rank 0's GPU alone, measured in this run"}
        out["partition_parity"] = bool(g["mesh_vertices"] == out["mesh_vertices_total"] and g["cubes"] == out["cubes_total"])
        out["speedup_vs_single_gpu"] = g["ms_per_frame"] / p["ms_per_frame"]
    return out


# ----------------------------------------------------------------------------------------------------------
# parity of the timed workload itself: the poses / volume the bench just computed against the reference's
# ----------------------------------------------------------------------------------------------------------
def parity_check(cam, frames, gpu_T, cpu_T, local):
    """gpu_T / cpu_T: {k: 4x4 result.T of PointToPlane(frame k+1 -> frame k)} from the timed GPU steps and from the CPU leg
    (the reference's float32 build).  Adds a float64-reference registration of the first pair and a two-frame integration with
    the GPU poses against the reference's CubeHandler."""
    from onepiece_b200.volume import CubeHandler
    from oracle import oracleapi, refapi
    out = {"frames_compared_with_cpu_leg": 0}
    dts, drs = [], []
    for k, Tc in cpu_T.items():
        if k in gpu_T:
            dt, dr = pose_delta(gpu_T[k], Tc)
            dts.append(dt); drs.append(dr)
    if dts:
        out.update(frames_compared_with_cpu_leg=len(dts), max_dt_m_vs_float32_reference=max(dts), max_drot_rad_vs_float32_reference=max(drs),
                   note="the float32 reference accumulates its 6x6 system sequentially in float; its own deviation from its float64 build is "
                        "reported next to the CUDA path's")
    ok = True
    if refapi.available("f64") and 0 in gpu_T:
        a, b = frames[0], frames[1]
        r64 = refapi.icp(b["cloud"], a["cloud"], a["normals"], np.eye(4), ICP_ITERS, ICP_THRESHOLD, "f64")
        dt, dr = pose_delta(gpu_T[0], r64["T"])
        out.update(dt_m_vs_float64_reference=dt, drot_rad_vs_float64_reference=dr, tolerance="1e-5 m / 1e-4 rad (north_star)")
        if 0 in cpu_T:
            ft, fr = pose_delta(cpu_T[0], r64["T"])
            out.update(float32_reference_dt_m_vs_float64=ft, float32_reference_drot_rad_vs_float64=fr)
        ok = ok and dt < 1e-5 and dr < 1e-4
    # integration: frames 1 and 2 with the GPU poses, CUDA volume vs the reference's CubeHandler (oracle port if it did not travel)
    use_ref = refapi.available("f32")
    rv = refapi.RefVolume(cam, VOXEL) if use_ref else oracleapi.OracleVolume(cam, VOXEL)
    gv = CubeHandler(cam, VOXEL, max_cubes=1 << 16, device=local)
    n_int = 0
    for k in (0, 1):
        if k not in gpu_T:
            continue
        pose = (frames[k]["pose"] @ np.asarray(gpu_T[k], np.float64)).astype(np.float32)
        gv.IntegrateImage(frames[k + 1]["depth"], frames[k + 1]["bgr"], pose)
        rv.integrate(frames[k + 1]["depth"], frames[k + 1]["bgr"], pose)
        n_int += 1
    if n_int:
        gi, gvx = gv.GetCubeMap()
        ri, rvx = rv.download()
        og = np.lexsort((gi[:, 2], gi[:, 1], gi[:, 0]))
        same_ids = bool(np.array_equal(gi[og], ri))
        same_vox = bool(same_ids and np.array_equal(gvx[og].view(np.uint32), np.ascontiguousarray(rvx, np.float32).view(np.uint32)))
        out.update(integrated_frames=n_int, cubes=int(len(gi)), cube_set_identical=same_ids, voxels_bit_identical=same_vox,
                   volume_checked_against="compiled reference CubeHandler" if use_ref else "oracle port")
        ok = ok and same_ids and same_vox
    gv.close()
    out["ok"] = bool(ok)
    return out


# ----------------------------------------------------------------------------------------------------------
# our arm
# ----------------------------------------------------------------------------------------------------------
def run_ours(args):
    import torch
    import torch.distributed as dist

    from onepiece_b200 import capi, scenes
    from onepiece_b200.volume import CubeHandler

    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if not torch.cuda.is_available():
        raise SystemExit("bench.py needs a CUDA device: onepiece_b200 has no CPU path")
    torch.cuda.set_device(local)
    if world > 1:
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    pk, pk_kind = peaks()
    W = max(args.warmup, 3)
    K = args.steps

    # ------------------------------------------------------------------------------------------------------
    # N > 1: the headline is the partitioned fusion of ONE stream (config 4); the config-2 replicas are the secondary
    # ------------------------------------------------------------------------------------------------------
    c4 = None
    clocks4 = None
    if world > 1 and not args.no_partitioned:
        sampler4 = ClockSampler(local)
        if rank == 0:
            sampler4.start()
        c4 = bench_config4(local, rank, world, K, W, True)
        clocks4 = sampler4.stop() if rank == 0 else None

    cam = scenes.Camera()
    frames = make_stream(cam, rank)
    stream = torch.cuda.Stream()
    vol = CubeHandler(cam, VOXEL, max_cubes=1 << 17, device=local, stream=stream.cuda_stream)
    icp = C.c_void_p()
    capi.check(capi.lib.opb_icp_create(local, C.c_void_p(stream.cuda_stream), C.byref(icp)))
    par = capi.IcpParams(ICP_ITERS, ICP_THRESHOLD, 1.0)
    res = capi.IcpResult()
    I16 = np.ascontiguousarray(np.eye(4, dtype=np.float32)).reshape(16)
    npx = cam.width * cam.height
    n_pts = len(frames[0]["cloud"])

    def dev(a):
        return torch.from_numpy(np.ascontiguousarray(a)).cuda()

    def pin(a):
        return torch.from_numpy(np.ascontiguousarray(a)).pin_memory()

    D = [{k: dev(f[k]) for k in ("depth", "bgr", "normals")} for f in frames]            # resident in HBM: `value`
    H = [{k: pin(f[k]) for k in ("depth", "bgr", "cloud", "normals")} for f in frames]   # pinned host: `e2e`
    pairs_host = torch.zeros((n_pts, 2), dtype=torch.int32).pin_memory()
    stats = capi.FrameStats()

    def new_cloud():
        h = C.c_void_p()
        capi.check(capi.lib.opb_cloud_create(local, None, C.byref(h)))
        return h

    def load_cloud(h, B):
        """PointCloud::LoadFromDepth of the frame on the device + its (analytic) normals; only enqueues"""
        capi.check(capi.lib.opb_cloud_load_from_depth(h, C.c_void_p(B["depth"].data_ptr()), capi.OPB_DEPTH_U16, C.c_void_p(B["bgr"].data_ptr()),
                                                      cam.fx, cam.fy, cam.cx, cam.cy, cam.width, cam.height, cam.depth_scale))
        capi.check(capi.lib.opb_cloud_set_normals(h, C.c_void_p(B["normals"].data_ptr()), n_pts))

    # `value`: every frame resident in HBM as a device cloud (images, points, normals)
    resident = [new_cloud() for _ in range(N_TRAJ)]
    nsz = C.c_size_t(0)
    for k in range(N_TRAJ):
        load_cloud(resident[k], D[k])
        capi.check(capi.lib.opb_cloud_size(resident[k], C.byref(nsz)))
        assert nsz.value == n_pts
    gpu_T = {}

    def value_step(s):
        k = s % (N_TRAJ - 1)
        # registration::PointToPlane(source = frame k+1, target = frame k) -> T with p_k = T p_{k+1}
        capi.check(capi.lib.opb_icp_point_to_plane_clouds(icp, resident[k + 1], resident[k], I16.ctypes.data_as(C.c_void_p), C.byref(par),
                                                          C.byref(res), None, 0))
        T = np.array(res.T[:], np.float64).reshape(4, 4).T
        gpu_T[k] = T
        pose = np.ascontiguousarray((frames[k]["pose"] @ T).astype(np.float32).T).reshape(16)
        vol.IntegrateImageDevice(D[k + 1]["depth"].data_ptr(), capi.OPB_DEPTH_U16, D[k + 1]["bgr"].data_ptr(), pose)

    # `e2e`: a streaming caller.  Three frame handles in rotation; each step uploads ONE new frame from pinned host memory
    # (depth, colour, normals: it is the source of the next registration and the target of the one after), registers the newest
    # loaded frame against the previous one, gets pose + inlier pairs back, integrates it, reads the frame counters.  The
    # sequence runs up and down the trajectory (0,1,..,8,7,..,0,1,..) so that consecutive frames are always neighbours.
    ring = [new_cloud() for _ in range(3)]
    period = 2 * (N_TRAJ - 1)

    def tri(s):
        m = s % period
        return m if m < N_TRAJ else period - m

    n_pairs_last = [0]

    def e2e_step(s, want_pairs=True):
        load_cloud(ring[(s + 2) % 3], H[tri(s + 2)])      # frame s+2 travels while frame s+1 is registered
        src, tgt = ring[(s + 1) % 3], ring[s % 3]
        capi.check(capi.lib.opb_icp_point_to_plane_clouds(icp, src, tgt, I16.ctypes.data_as(C.c_void_p), C.byref(par), C.byref(res),
                                                          C.c_void_p(pairs_host.data_ptr()) if want_pairs else None, n_pts if want_pairs else 0))
        n_pairs_last[0] = res.n_local_pairs
        T = np.array(res.T[:], np.float64).reshape(4, 4).T
        pose = np.ascontiguousarray((frames[tri(s)]["pose"] @ T).astype(np.float32).T).reshape(16)
        capi.check(capi.lib.opb_volume_integrate_cloud(vol._h, src, pose.ctypes.data_as(C.c_void_p)))
        capi.check(capi.lib.opb_volume_frame_stats(vol._h, C.byref(stats)))
        capi.check(capi.lib.opb_icp_wait_pairs(icp))      # the pair list travelled under the integration (opb_icp_set_async_pairs)

    def e2e_prime():
        load_cloud(ring[0], H[tri(0)])
        load_cloud(ring[1], H[tri(1)])

    # the literal reference signature: PointToPlane(source cloud, target cloud with normals) from HOST arrays, pairs back,
    # then IntegrateImage(depth, rgb, pose) from host images (what the C++ drop-in calls)
    def hostsig_step(s):
        k = s % (N_TRAJ - 1)
        a, b = H[k], H[k + 1]
        capi.check(capi.lib.opb_icp_point_to_plane(icp, C.c_void_p(b["cloud"].data_ptr()), n_pts, C.c_void_p(a["cloud"].data_ptr()),
                                                   C.c_void_p(a["normals"].data_ptr()), n_pts, I16.ctypes.data_as(C.c_void_p),
                                                   C.byref(par), C.byref(res), C.c_void_p(pairs_host.data_ptr()), n_pts))
        T = np.array(res.T[:], np.float64).reshape(4, 4).T
        pose = np.ascontiguousarray((frames[k]["pose"] @ T).astype(np.float32).T).reshape(16)
        capi.check(capi.lib.opb_volume_integrate(vol._h, C.c_void_p(b["depth"].data_ptr()), capi.OPB_DEPTH_U16,
                                                 C.c_void_p(b["bgr"].data_ptr()), pose.ctypes.data_as(C.c_void_p)))
        capi.check(capi.lib.opb_volume_frame_stats(vol._h, C.byref(stats)))

    def timed(step, steps, warmup, prime=None):
        if prime:
            prime()
        for s in range(warmup):
            step(s)
        vol.Synchronize()
        barrier()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        t0 = time.perf_counter()
        e0.record(stream)
        for s in range(steps):
            step(warmup + s)
        e1.record(stream)
        vol.Synchronize()
        wall = time.perf_counter() - t0
        barrier()
        ms = e0.elapsed_time(e1)
        if world > 1:
            t = torch.tensor([ms, wall], device="cuda", dtype=torch.float64)
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
            ms, wall = float(t[0].item()), float(t[1].item())
        return ms, wall

    sampler = ClockSampler(local)
    if rank == 0:
        sampler.start()
    ms_dev, _ = timed(value_step, K, W)
    clocks = sampler.stop() if rank == 0 else None

    # the same region again with per-kernel CUDA events (on the stream the kernels run on) for the rooflines
    vol.SetProfiling(True)
    vol.ProfileRead(reset=True)
    capi.lib.opb_icp_set_profiling(icp, 1)
    icp_loop_ms, icp_grid_ms, upd_sum, cubes = 0.0, 0.0, 0, 0
    nprof_steps = min(K, 100)
    for s in range(nprof_steps):
        value_step(W + s)
        a, b = C.c_float(0), C.c_float(0)
        capi.lib.opb_icp_last_timing(icp, C.byref(a), C.byref(b))
        icp_grid_ms += a.value
        icp_loop_ms += b.value
        st = vol.FrameStats()
        upd_sum += st.updated_voxels
        cubes = st.frame_cubes
    sel_ms, int_ms, nprof = vol.ProfileRead(reset=True)
    vol.SetProfiling(False)
    capi.lib.opb_icp_set_profiling(icp, 0)
    icp_launches = C.c_int(0)
    capi.lib.opb_icp_last_launch_count(icp, C.byref(icp_launches))
    searched = C.c_uint64(0)
    capi.lib.opb_icp_last_search_count(icp, C.byref(searched))

    # end to end, host buffers in / results out inside the timed region (CUDA events on the work stream bracket it; every call
    # is synchronous, so the wall clock agrees)
    capi.check(capi.lib.opb_icp_set_async_pairs(icp, 1))
    e2e_ms, e2e_wall = timed(e2e_step, K, W, e2e_prime)
    n_pairs = int(n_pairs_last[0])
    e2e_nopairs_ms, _ = timed(lambda s: e2e_step(s, False), K, W, e2e_prime)
    capi.check(capi.lib.opb_icp_set_async_pairs(icp, 0))
    e2e_sync_ms, _ = timed(e2e_step, K, W, e2e_prime)
    hostsig_ms, _ = timed(hostsig_step, K, W)

    c5 = None
    if world > 1 and not args.no_partitioned:
        vol.close()  # make room: the partitioned volume is a second pool on the same GPU
        try:
            c5 = bench_config5(cam, local, rank, world, min(25 * K, 500))
        except Exception as exc:  # noqa: BLE001  -- a secondary measurement must never cost the headline line
            c5 = {"error": f"{type(exc).__name__}: {exc}"}
    if rank != 0:
        if world > 1:
            dist.destroy_process_group()
        return
    upd = upd_sum / max(nprof_steps, 1)
    alg_bytes = upd * 2 * 20 + npx * (2 + 3)  # updated voxels read+written at 20 B, one pass over u16 depth + colour
    k2_ms = int_ms / max(nprof, 1)
    achieved = alg_bytes / (k2_ms * 1e-3) / 1e9 if k2_ms > 0 else 0.0
    icp_iter_ms = icp_loop_ms / nprof_steps / (ICP_ITERS + 1)
    icp_bytes = n_pts * 12 + n_pts * 36  # N_s*12 (source) + N_inl*(12+12+12) (nn point, normal, source)
    icp_ach = icp_bytes / (icp_iter_ms * 1e-3) / 1e9 if icp_iter_ms > 0 else 0.0
    # per step: ICP = grid build + pass loop + Kabsch sums (+ pair compaction in e2e); volume = pack, select, integrate
    launches_per_step = icp_launches.value + 3
    replicas = {"metric": METRIC, "value": world * K / (ms_dev * 1e-3), "unit": UNIT, "ms_per_step": ms_dev / K, "scaling": "weak",
                "what": "one independent config-2 stream (ICP + integrate at 640x480, 5 mm) per GPU, no data-path collective",
                "e2e": world * K / (e2e_ms * 1e-3)}
    details = {"cubes_per_frame": cubes, "updated_voxels_per_frame": int(upd), "icp_points": n_pts,
               "icp_exact_searches_per_frame": int(searched.value), "icp_queries_per_frame": n_pts * (ICP_ITERS + 1),
               "step_breakdown_ms": {"icp_grid_build": icp_grid_ms / nprof_steps, "icp_pass_loop_and_finaliser": icp_loop_ms / nprof_steps,
                                     "cube_selection": sel_ms / max(nprof, 1), "voxel_update": k2_ms}}
    roofline = {"bound": "hbm", "kernel": "integrate_pipelined_kernel (the voxel update, the kernel north_star sets the >=60% target for)",
                "achieved": achieved, "peak": pk["hbm_gbs"], "peak_source": pk_kind, "unit": "GB/s",
                "frac": achieved / pk["hbm_gbs"], "algorithmic_bytes_per_launch": int(alg_bytes), "kernel_ms": k2_ms,
                "traffic": measured_traffic("integrate_pipelined_kernel")[0],
                "traffic_source": measured_traffic("integrate_pipelined_kernel")[1]}
    roofline_icp = {"bound": "hbm", "kernel": "icp_loop2_kernel, one pass of the persistent ICP loop = certify/search + 8x8 sums + solve "
                                              "(time-dominant; working set L2-resident, latency-bound)",
                    "achieved": icp_ach, "peak": pk["hbm_gbs"], "unit": "GB/s", "frac": icp_ach / pk["hbm_gbs"],
                    "algorithmic_bytes_per_launch": int(icp_bytes), "kernel_ms": icp_iter_ms,
                    "traffic": (measured_traffic("icp_loop2_kernel")[0] or 0) // (ICP_ITERS + 1) or None,
                    "traffic_source": measured_traffic("icp_loop2_kernel")[1],
                    "traffic_note": "DRAM bytes of the whole 31-pass launch / 31: the passes run out of L2"}
    e2e = {"value": K / (e2e_ms * 1e-3), "unit": UNIT,
           "h2d_bytes_per_step": npx * (2 + 3) + n_pts * 12, "d2h_bytes_per_step": n_pairs * 8 + 256 + 64,
           "clock": "CUDA events on the work stream around K steps of: upload one new frame (u16 depth + colour + normals, pinned host) "
                    "into a device cloud, PointToPlane on the device clouds with the inlier pairs written back to pinned host memory "
                    "(the call returns with the pose, the list arrives under the integration and is collected with opb_icp_wait_pairs "
                    "at the end of the step), IntegrateImage of the registered frame, FrameStats",
           "pairs_inside_the_call": {"value": K / (e2e_sync_ms * 1e-3), "what": "the same with the pair list complete when PointToPlane returns"},
           "wall_clock_value": K / e2e_wall,
           "pose_only": {"value": K / (e2e_nopairs_ms * 1e-3), "d2h_bytes_per_step": 320, "what": "the same without the inlier pairs"},
           "reference_signature": {"value": K / (hostsig_ms * 1e-3), "h2d_bytes_per_step": 3 * n_pts * 12 + npx * 5,
                                   "d2h_bytes_per_step": n_pairs * 8 + 320,
                                   "what": "PointToPlane(source, target) from HOST point arrays (both clouds + normals uploaded per call, "
                                           "as the reference's signature implies) + IntegrateImage from host images"}}
    if world == 1:
        out = {"metric": METRIC, "value": K / (ms_dev * 1e-3), "unit": UNIT, "n_gpus": 1, "steps": K, "warmup": W,
               "ms_per_step": ms_dev / K, "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f32",
               "data": "synthetic", "config": config_of(1), "details": details, "roofline": roofline, "roofline_icp": roofline_icp,
               "e2e": e2e, "gpu_launches": launches_per_step * K, "clocks": clocks}
    else:
        line4 = config4_line(c4, world, K, pk["hbm_gbs"]) if c4 else None
    if world > 1 and line4 is None:   # --no-partitioned: only the replicas were measured
        out = {"metric": METRIC, "value": replicas["value"], "unit": UNIT, "n_gpus": world, "steps": K, "warmup": W,
               "ms_per_step": ms_dev / K, "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f32",
               "data": "synthetic", "config": dict(config_of(1), sharding=replicas["what"]), "details": details, "roofline": roofline,
               "roofline_icp": roofline_icp, "e2e": dict(e2e, value=replicas["e2e"]), "gpu_launches": launches_per_step * K * world,
               "clocks": clocks}
    elif world > 1:
        out = {"metric": METRIC, "value": line4["frames_per_s"], "unit": UNIT, "n_gpus": world, "steps": K, "warmup": W,
               "ms_per_step": line4["ms_per_frame"], "higher_is_better": True, "scaling": "strong", "vs_baseline": None, "dtype": "f32",
               "data": "synthetic", "config": config_of(world), "partitioned_fusion": line4,
               "roofline": dict(line4.get("roofline_per_gpu", {}), kernel="integrate_pipelined_kernel on the slowest rank's share of the frame",
                                traffic=None),
               "e2e": {"value": line4["e2e_frames_per_s"], "unit": UNIT, "h2d_bytes_per_step": 1280 * 960 * 7,
                       "d2h_bytes_per_step": 0, "clock": line4["e2e_note"]},
               "gpu_launches": 3 * K * world, "clocks": clocks4,
               "replicas": dict(replicas, details=details, roofline=roofline, e2e_detail=e2e, clocks=clocks)}
        if c5 is not None:
            out["dense_fusion_pipeline"] = dict(c5, single_gpu_frames_per_s=replicas["e2e"] / world)
    if not args.no_cpu_baseline and world == 1:
        cb, _, _, cpu_T = cpu_reference_fps(30, 1, budget_s=20.0)
        out["cpu_baseline"] = cb
        try:
            out["parity_check"] = parity_check(cam, frames, gpu_T, cpu_T, local)
        except Exception as exc:  # noqa: BLE001
            out["parity_check"] = {"ok": False, "error": f"{type(exc).__name__}: {exc}"}
    if world == 1 and not args.no_partitioned:
        # the same config-4 stream on this one GPU, so that a per-N series of the partitioned workload starts at N = 1
        try:
            vol.close()
            out["partitioned_fusion"] = config4_line(bench_config4(local, 0, 1, min(K, 50), W, False), 1, min(K, 50), pk["hbm_gbs"])
        except Exception as exc:  # noqa: BLE001  -- a secondary measurement must never cost the headline line
            out["partitioned_fusion"] = {"error": f"{type(exc).__name__}: {exc}"}
    if world == 1 and not args.no_partitioned:
        try:
            out["packed16_voxels"] = bench_packed16(cam, frames, local, min(K, 50), pk["hbm_gbs"])
        except Exception as exc:  # noqa: BLE001
            out["packed16_voxels"] = {"error": f"{type(exc).__name__}: {exc}"}
    if not args.no_odometry and world == 1:
        out["dense_odometry"] = bench_dense_odometry(frames, cam, local, min(K, 100), not args.no_cpu_baseline)
    print(json.dumps(out), flush=True)
    capi.lib.opb_icp_destroy(icp)
    if world > 1:
        dist.destroy_process_group()


def main():
    args = parse_args()
    if args.impl == "reference":
        run_reference(args)
    else:
        run_ours(args)


if __name__ == "__main__":
    main()
