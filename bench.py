#!/usr/bin/env python
"""bench.py - evaluate_ybar cube voxels/s on B200 (the BASELINE.json metric).

Workload: the density path `Component.evaluate_ybar`
(reference: popkinmocks/components/base.py:789-882) on a synthetic 201x201-pixel cube over the
full MILES_BASTI grid (nt=53, nz=12), nv=101, lambda 4700-6500 A -> n_fft=4907; voxel = one
(omega, x1, x2) sample.  One "step" = one whole-cube evaluate_ybar.

    python bench.py [--gpus N] [--steps K] [--warmup W]          # this repo's CUDA path
    python bench.py --impl reference [...]                        # the reference's CPU algorithm

N > 1 (torchrun, one rank per GPU): strong scaling of the named configuration - the flattened
pixel axis of ONE 201x201-pixel cube is sharded over the ranks in contiguous ranges (pixels are
independent).  Every rank's density shard is a slice of the same seeded
cube (pkm_synth_log_density is indexed by the global element index).  Each rank evaluates its
range tile by tile; a finished tile is transposed locally and moved into rank 0's cube by a
strided copy-engine transfer over NVLink peer memory (CUDA IPC), followed by one small NCCL
all-reduce as the barrier; all inside the timed region (`--gather nccl`: NCCL send/recv).
`--scaling weak` evaluates a (201*N)x201 mosaic instead (201x201 pixels per rank).
After the timed region rank 0 checks the assembled cube: 16 spot pixels against the oracle and
a bit checksum that must be the same for every N (`verify`); a mismatch exits non-zero.

The JSON line printed by rank 0 follows the driver contract: `value` = device-resident
throughput, `e2e` = the same call through the C ABI with pinned HOST buffers (H2D of the
density and D2H of the cube inside the timed region), `roofline` for the dominant kernel
(timed live with CUDA events on its launch stream), `cpu_baseline` = the oracle port on the
box's host cores (bounded sample), `clocks` sampled with nvidia-smi during the timed region.
"""
import argparse
import json
import os
import statistics
import subprocess
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

METRIC = "evaluate_ybar cube voxels/s"
UNIT = "voxel/s"
SEED = 20241019
VERIFY_TOL = 1e-10
B_VOX_NOTE = "SURVEY.md 8d canonical bytes/voxel with P-hat staged through HBM once"


def parse_args():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=10)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--config", default="C4", choices=["C2", "C3", "C4", "C5"],
                    help="BASELINE.json configuration (SURVEY.md 8d): C4 = the headline 201x201 synthetic cube "
                         "(default, the driver's bench line); C2 GrowingDisk 51x51 (path A), C3 Mixture 101x101 + "
                         "marginals + moments, C5 evaluate_ybar + ShotNoise 101x101 (bench_workloads.py)")
    ap.add_argument("--nx", type=int, default=201, help="pixels per side of one cube")
    ap.add_argument("--nv", type=int, default=101)
    ap.add_argument("--lmd", type=float, nargs=2, default=(4700.0, 6500.0))
    ap.add_argument("--scaling", default="strong", choices=["weak", "strong"])
    ap.add_argument("--gather", default="peer", choices=["peer", "nccl"],
                    help="N>1: assemble the cube on rank 0 by peer-memory stores from the transpose kernel, or by NCCL send/recv")
    ap.add_argument("--gather-tiles", type=int, default=1, help="sub-blocks per rank for the NCCL gather")
    ap.add_argument("--e2e-steps", type=int, default=3)
    ap.add_argument("--no-e2e", action="store_true")
    ap.add_argument("--workspace-gb", type=float, default=0.0, help="device workspace cap of the plan (0 = library default)")
    ap.add_argument("--no-cpu", action="store_true", help="skip the cpu_baseline leg")
    ap.add_argument("--no-verify", action="store_true", help="skip the oracle spot check of the evaluated cube")
    ap.add_argument("--no-other", action="store_true", help="skip the brief timings of the other hot-path rows")
    ap.add_argument("--cpu-procs", type=int, default=0, help="host processes of the CPU legs (0 = all cores, max 64)")
    ap.add_argument("--cpu-px", type=int, default=2, help="pixels per CPU task")
    return ap.parse_args()


# ----------------------------------------------------------------------------------------------
# workload description shared by both arms
# ----------------------------------------------------------------------------------------------
def build_cube(args, nx1=None, nx2=None):
    import popkinmocks_b200 as pkm
    ssps = pkm.milesSSPs(lmd_min=args.lmd[0], lmd_max=args.lmd[1])
    # the spatial extents scale with the pixel counts so dx, dy (and the result per pixel) are unchanged
    nx1 = nx1 or args.nx
    nx2 = nx2 or args.nx
    return pkm.ifu_cube.IFUCube(ssps=ssps, nx1=nx1, nx2=nx2, nv=args.nv,
                                x1rng=(-1.0 * nx1 / args.nx, 1.0 * nx1 / args.nx),
                                x2rng=(-1.0 * nx2 / args.nx, 1.0 * nx2 / args.nx), vrng=(-1000.0, 1000.0))


def workload_config(args, cube):
    nz, nt = cube.ssps.par_dims
    return {
        "workload": f"C4 synthetic density cube {args.nx}x{args.nx} px, full MILES_BASTI grid "
                    f"nt={nt} nz={nz}, nv={args.nv}, lmd {args.lmd[0]:.0f}-{args.lmd[1]:.0f} A, "
                    f"n_fft={cube.ssps.n_fft} (path B: Component.evaluate_ybar on log_p_tvxz)",
        "nx1": args.nx, "nx2": args.nx, "nv": args.nv, "nt": int(nt), "nz": int(nz),
        "n_fft": int(cube.ssps.n_fft),
        "l2": "inputs (20.8 GB density per cube) are far larger than the 126 MB L2; no flush needed",
    }


# ----------------------------------------------------------------------------------------------
# CPU legs: the oracle port of the reference algorithm on host cores
# ----------------------------------------------------------------------------------------------
_CPU = {}


def _cpu_init(consts):
    _CPU.update(consts)


def _cpu_task(seed):
    """One tile of `px` pixels through the reference's literal algorithm (oracle/restatement.py)."""
    from oracle import restatement as R
    c = _CPU
    rng = np.random.default_rng(seed)
    p = rng.random((c["nt"], c["nv"], 1, c["px"], c["nz"]))
    t0 = time.perf_counter()
    y = R.ybar_density_literal(p, c["v_edg"], c["FXw"], c["par_dims"], c["n_fft"], c["dv"],
                               c["delta_t"], c["delta_z"], c["dx"], c["dy"])
    return time.perf_counter() - t0, float(y.sum())


def _cpu_task_gaussian(seed):
    """One tile of `px` pixels through the oracle's path A (parametric.py:355-373)."""
    from oracle import restatement as R
    c = _CPU
    rng = np.random.default_rng(seed)
    P = rng.random((c["nt"], 1, c["px"], c["nz"]))
    P /= P.sum()
    mu = 400.0 * (rng.random((c["nt"], 1, c["px"])) - 0.5)
    sg = 20.0 + 200.0 * rng.random((c["nt"], 1, c["px"]))
    t0 = time.perf_counter()
    y = R.ybar_gaussian(P, mu, sg, c["FXw"], c["par_dims"], c["n_fft"], c["dv"])
    return time.perf_counter() - t0, float(y.sum())


def cpu_consts(cube, px):
    s = cube.ssps
    nz, nt = s.par_dims
    return dict(nt=int(nt), nz=int(nz), nv=cube.nv, px=px, v_edg=cube.v_edg, FXw=s.FXw,
                par_dims=tuple(s.par_dims), n_fft=s.n_fft, dv=s.dv, delta_t=s.delta_t,
                delta_z=s.delta_z, dx=cube.dx, dy=cube.dy)


def cpu_procs(args):
    n = os.cpu_count() or 1
    try:
        n = len(os.sched_getaffinity(0))
    except AttributeError:
        pass
    return max(1, min(args.cpu_procs or n, 64))


class CpuArm:
    """A fork pool of host workers, created BEFORE CUDA is initialised in this process."""

    def __init__(self, args, cube, task=None):
        import multiprocessing as mp
        self.task = task or _cpu_task
        self.nproc = cpu_procs(args)
        self.px = args.cpu_px
        self.n_fft = cube.ssps.n_fft
        self.pool = mp.get_context("fork").Pool(self.nproc, initializer=_cpu_init,
                                                initargs=(cpu_consts(cube, self.px),))
        self.seed = 1000

    def step(self):
        """One bounded sample: every worker evaluates one tile.  Returns (seconds, voxels)."""
        seeds = list(range(self.seed, self.seed + self.nproc))
        self.seed += self.nproc
        t0 = time.perf_counter()
        self.pool.map(self.task, seeds, chunksize=1)
        dt = time.perf_counter() - t0
        return dt, self.nproc * self.px * self.n_fft

    def close(self):
        self.pool.close()
        self.pool.join()


def run_reference(args):
    """--impl reference: the reference's CPU algorithm on all host cores, bounded sample per step."""
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    cube = build_cube(args)
    arm = CpuArm(args, cube)
    for _ in range(args.warmup):
        arm.step()
    tot_s, tot_v = 0.0, 0
    for _ in range(args.steps):
        s, v = arm.step()
        tot_s += s
        tot_v += v
    arm.close()
    value = tot_v / tot_s
    sample = (f"{arm.nproc} workers x {arm.px} px per step (oracle port of base.py:836-859, "
              f"irfft per (t,z)); reference is single-threaded, one tile per core")
    line = {
        "impl": "reference", "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": args.gpus,
        "steps": args.steps, "warmup": args.warmup, "ms_per_step": 1e3 * tot_s / max(args.steps, 1),
        "higher_is_better": True, "scaling": args.scaling, "vs_baseline": None, "dtype": "f64",
        "data": "synthetic", "config": workload_config(args, cube),
        "cpu_baseline": {"value": value, "unit": UNIT, "cores": arm.nproc, "kind": "port", "sample": sample},
        "e2e": {"value": value, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    print(json.dumps(line))


# ----------------------------------------------------------------------------------------------
# clocks
# ----------------------------------------------------------------------------------------------
class ClockSampler:
    FIELDS = ("clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,"
              "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
              "clocks_event_reasons.sw_power_cap")

    def __init__(self, uuid):
        self.proc = None
        self.path = f"/tmp/pkm_clocks_{os.getpid()}.csv"
        try:
            self.fh = open(self.path, "w")
            self.proc = subprocess.Popen(
                ["nvidia-smi", "-i", uuid, f"--query-gpu={self.FIELDS}", "--format=csv,noheader,nounits",
                 "-lms", "50"], stdout=self.fh, stderr=subprocess.DEVNULL)
        except Exception:  # noqa: BLE001 - nvidia-smi missing
            self.proc = None
        self.marks = []

    def mark(self):
        """Current number of lines written (to cut out the timed window later)."""
        try:
            self.fh.flush()
            with open(self.path) as f:
                return sum(1 for _ in f)
        except Exception:  # noqa: BLE001
            return 0

    def stop(self, windows):
        if self.proc is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        time.sleep(0.12)
        self.proc.terminate()
        try:
            self.proc.wait(timeout=5)
        except Exception:  # noqa: BLE001
            self.proc.kill()
        self.fh.close()
        rows = []
        with open(self.path) as f:
            lines = [ln.strip() for ln in f if ln.strip()]
        os.unlink(self.path)
        keep = []
        for a, b in windows:
            keep += lines[a:max(b, a + 1)]
        if not keep:
            keep = lines
        reasons = set()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        power = []
        for ln in keep:
            parts = [p.strip() for p in ln.split(",")]
            if len(parts) < 7:
                continue
            try:
                rows.append((float(parts[0]), float(parts[1])))
                power.append(float(parts[2]))
            except ValueError:
                continue
            for nm, val in zip(names, parts[3:7]):
                if val.lower().startswith("active"):
                    reasons.add(nm)
        if not rows:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["no samples"]}
        return {"sm_mhz": statistics.median(r[0] for r in rows), "sm_max_mhz": max(r[1] for r in rows),
                "reasons": sorted(reasons), "samples": len(rows),
                "power_w_max": max(power) if power else None}


# ----------------------------------------------------------------------------------------------
# the CUDA arm
# ----------------------------------------------------------------------------------------------
def run_b200(args):
    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if world != args.gpus and world > 1:
        raise SystemExit(f"--gpus {args.gpus} but WORLD_SIZE={world}")

    # ---- host-side setup and the CPU baseline leg (fork pool: before any CUDA initialisation)
    cube = build_cube(args)
    cpu = None
    if rank == 0 and world == 1 and not args.no_cpu:
        arm = CpuArm(args, cube)
        arm.step()                      # warm the workers (imports, first-touch)
        tot_s, tot_v = 0.0, 0
        while tot_s < 8.0:
            s, v = arm.step()
            tot_s += s
            tot_v += v
        arm.close()
        cpu = {"value": tot_v / tot_s, "unit": UNIT, "cores": arm.nproc, "kind": "port",
               "sample": f"{tot_v // cube.ssps.n_fft} pixels in {tot_s:.1f} s wall: {arm.nproc} workers x "
                         f"{arm.px} px tiles of the same synthetic workload through oracle/restatement.py "
                         f"ybar_density_literal (reference order: irfft per (t,z), base.py:836-859)"}

    import torch
    import torch.distributed as dist
    from popkinmocks_b200 import _lib, engine

    # NUMA-local cores for this rank: the page-locked e2e buffers are then allocated next to the GPU
    numa_cpus = None
    if world > 1 and not os.environ.get("PKM_BENCH_NO_AFFINITY"):
        from popkinmocks_b200 import sharding as _sh
        numa_cpus = _sh.bind_to_gpu_numa_node(local)
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)

    nz, nt = cube.ssps.par_dims
    nv, nx, n_fft = args.nv, args.nx, int(cube.ssps.n_fft)
    plan = engine.Plan(cube, device=local)
    if args.workspace_gb > 0:
        plan.set_workspace_limit(int(args.workspace_gb * (1 << 30)))

    # ---- this rank's share of the pixels
    from popkinmocks_b200 import sharding
    # Pixels are independent, so the cube is sharded at PIXEL granularity (contiguous ranges of the
    # flattened (x1, x2) index, sizes differing by at most one pixel): locally the range is handled
    # as an (npix x 1) image, globally it is a contiguous slab of the cube's pixel axis.
    if world == 1:
        total_pix, r0, r1 = nx * nx, 0, nx * nx
    elif args.scaling == "strong":            # ONE nx x nx cube sharded over the ranks
        total_pix = nx * nx
        r0, r1 = sharding.my_rows(total_pix, world, rank)
    else:                                     # weak: a (world*nx) x nx mosaic, nx*nx pixels per rank
        total_pix = nx * nx * world
        r0, r1 = rank * nx * nx, (rank + 1) * nx * nx
    if world == 1:
        rows, nxl = nx, nx                    # keep the natural (nx, nx) image on one GPU
    else:
        rows, nxl = r1 - r0, 1
    npix = rows * nxl
    nx1_total = total_pix if world > 1 else nx
    voxels_per_step = total_pix * n_fft

    # ---- synthetic density (log p, as Component(cube, log_p_tvxz) holds it), generated on device by a
    # counter-based generator indexed by the GLOBAL element index (pkm_synth_log_density): every
    # rank's shard is a slice of ONE seeded cube, so N = 1, 2, 4, 8 evaluate identical inputs
    gen = torch.Generator(device=dev).manual_seed(20241019 + rank)   # only for the other_paths operands
    dV = float(np.mean(cube.ssps.delta_t) * np.mean(cube.ssps.delta_z) * cube.dx * cube.dy * cube.ssps.dv)
    synth = dict(nt=int(nt), nv=nv, total_pix=int(total_pix), nz=int(nz), seed=SEED,
                 scale=1.0 / (0.5 * nt * nv * total_pix * nz * dV))
    logp = engine.synth_log_density(pix0=r0, npix=npix, device=local, **synth).view(nt, nv, rows, nxl, nz)

    gather_note = None
    if world == 1:
        out = torch.empty((n_fft, rows, nxl), dtype=torch.float64, device=dev)
        tg = None
    else:
        if args.gather == "peer":
            # rank 0 owns the cube; every rank moves its finished tiles into it over NVLink (CUDA IPC).
            # If peer mapping is not available on this box, all ranks fall back to the NCCL gather.
            ok = torch.ones(1, device=dev)
            try:
                pc = sharding.PeerCube(n_fft, nx1_total, nxl, dev, dist)
            except Exception as exc:  # noqa: BLE001
                ok.zero_()
                gather_note = f"peer-memory cube unavailable ({type(exc).__name__}); used the NCCL gather"
            dist.all_reduce(ok, op=dist.ReduceOp.MIN)
            if float(ok.item()) < 0.5:
                args.gather = "nccl"
        if args.gather == "peer":
            out = pc.cube
            tg = None
            flag = torch.zeros(1, device=dev)
    if world > 1 and args.gather == "nccl":
        tg = sharding.TileGather(nx1_total, nxl, n_fft, dev, dist=dist, ntile=args.gather_tiles)
        out = torch.empty((n_fft, nx1_total, nxl), dtype=torch.float64, device=dev) if rank == 0 else None

    def step():
        if world == 1:
            plan.ybar_density(logp, is_log=True, out=out)
            return
        if args.gather == "peer":
            pc.evaluate_rows(plan, logp, 0, rows, cube_row0=r0)
            if not os.environ.get("PKM_BENCH_NO_BARRIER"):
                dist.all_reduce(flag)  # stream-ordered barrier: all peer stores precede rank 0's next use
            return
        # sub-blocks of this rank's rows -> spaxel-major tiles, each shipped to rank 0 over NCCL
        # while the next one is computed; rank 0 transposes once to [n_fft, x1, x2]
        for t, (a, b, view) in enumerate(tg.my_tiles()):
            if b > a:
                plan.ybar_density(logp, is_log=True, rows=(a, b), out=view, layout=_lib.LAYOUT_PIXEL_MAJOR)
            tg.post(t)
        full = tg.finish()
        if rank == 0:
            _lib.check(_lib.load().pkm_transpose_to_cube(_lib.ptr(full), total_pix, n_fft,
                                                         _lib.ptr(out), local, None))

    def barrier():
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()
            torch.cuda.synchronize()

    uuid = "GPU-" + str(torch.cuda.get_device_properties(dev).uuid).replace("GPU-", "")
    sampler = ClockSampler(uuid) if rank == 0 else None
    fp64_peak = engine.fp64_fma_peak(local)

    for _ in range(max(args.warmup, 3)):
        step()
    barrier()
    plan.set_profiling(True)
    l0 = plan.launches
    w0 = sampler.mark() if sampler else 0
    e0 = torch.cuda.Event(enable_timing=True)
    e1 = torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(args.steps):
        step()
    e1.record()
    barrier()
    w1 = sampler.mark() if sampler else 0
    ms_total = e0.elapsed_time(e1)
    ktimes = plan.kernel_times()
    plan.set_profiling(False)
    l1 = plan.launches
    kernel_timing = "CUDA events inside the timed region"
    if world > 1 and args.gather == "peer" and not os.environ.get("PKM_TILE_LANES"):
        # a shard's pixel tiles run on two concurrent lanes (each fills the other's wave tails), so the
        # per-kernel event intervals of the timed region overlap: the kernel table and the roofline come
        # from three extra, untimed steps with the lanes serialised (`value` is the timed region's)
        os.environ["PKM_TILE_LANES"] = "1"
        step()
        barrier()
        plan.set_profiling(True)
        for _ in range(3):
            step()
        barrier()
        ktimes = {k: (m * args.steps / 3.0, c * args.steps / 3.0) for k, (m, c) in plan.kernel_times().items()}
        plan.set_profiling(False)
        os.environ.pop("PKM_TILE_LANES")
        kernel_timing = "CUDA events in 3 extra steps with the tile lanes serialised (PKM_TILE_LANES=1)"
    launches = l1 - l0 + (args.steps if world > 1 and rank == 0 and args.gather == "nccl" else 0)
    per_rank = None
    if world > 1:
        # diagnostics: every rank's own step time and kernel sum (the headline uses the MAX)
        mine = torch.tensor([ms_total / args.steps, sum(m for m, _ in ktimes.values()) / args.steps],
                            dtype=torch.float64, device=dev)
        allr = [torch.zeros_like(mine) for _ in range(world)]
        dist.all_gather(allr, mine)
        per_rank = {"step_ms": [round(float(a[0]), 3) for a in allr],
                    "kernel_sum_ms": [round(float(a[1]), 3) for a in allr]}
        t = torch.tensor([ms_total], dtype=torch.float64, device=dev)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        ms_total = float(t.item())
    ms_step = ms_total / args.steps
    value = voxels_per_step / (ms_step * 1e-3)

    # ---- verification of the assembled cube (after the timed region, rank 0): spot pixels against
    # the oracle's literal restatement of base.py:836-859 on the same seeded inputs, plus an
    # order-independent bit checksum that must be identical for every N
    verify = None
    if rank == 0 and not args.no_verify:
        verify = verify_cube(out.reshape(n_fft, total_pix), cube, synth, total_pix)
    spot_px, spot_ref = [], None
    if rank == 0 and verify is not None:         # kept for the class-API leg below (`out` is reused)
        spot_px = [p for p in verify["pixel_index"] if r0 <= p < r1][:4]
        spot_ref = out.reshape(n_fft, total_pix)[:, spot_px].cpu().numpy()
    if world > 1:
        bad = torch.tensor([1.0 if (verify and not verify["ok"]) else 0.0], device=dev)
        dist.all_reduce(bad, op=dist.ReduceOp.MAX)
        verify_failed = bool(bad.item() > 0)
    else:
        verify_failed = bool(verify and not verify["ok"])

    # ---- the other rows of the hot-path table, timed briefly on the same cube (N=1 only; reported
    # under "other_paths", not part of `value`)
    other = None
    if world == 1 and not args.no_other:
        def timed(fn, reps=3):
            fn()
            torch.cuda.synchronize()
            a0 = torch.cuda.Event(enable_timing=True)
            a1 = torch.cuda.Event(enable_timing=True)
            a0.record()
            for _ in range(reps):
                fn()
            a1.record()
            torch.cuda.synchronize()
            return a0.elapsed_time(a1) / reps
        other = {}
        P = torch.rand((nt, rows, nxl, nz), dtype=torch.float64, device=dev, generator=gen)
        P /= P.sum()
        mu_v = 400.0 * (torch.rand((nt, rows, nxl), dtype=torch.float64, device=dev, generator=gen) - 0.5)
        sig_v = 20.0 + 200.0 * torch.rand((nt, rows, nxl), dtype=torch.float64, device=dev, generator=gen)
        hbm_peak_o = 6650.0
        try:
            hbm_peak_o = float(json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))["hbm_gbs"])
        except Exception:  # noqa: BLE001
            pass

        def hbm_roof(nbytes, ms):
            return {"bound": "hbm", "achieved": nbytes / (ms * 1e-3) / 1e9, "peak": hbm_peak_o, "unit": "GB/s",
                    "frac": nbytes / (ms * 1e-3) / 1e9 / hbm_peak_o}

        ms = timed(lambda: plan.ybar_gaussian(P, mu_v, sig_v, out=out))
        # executed FP64 work of path A per (k, t, x): 2*nz FMAs of the metallicity contraction (complex x
        # real), a 6-operation recurrence for exp(-i mu w - sig^2 w^2 / 2) and a complex multiply-add
        flops_a = (n_fft // 2 + 1) * nt * (4.0 * nz + 14.0) * npix
        other["gaussian_path_A"] = {"voxel_per_s": voxels_per_step / (ms * 1e-3), "ms": ms,
                                    "ref": "ParametricComponent.evaluate_ybar parametric.py:335-407",
                                    "roofline": {"bound": "fp64", "achieved": flops_a / (ms * 1e-3) / 1e12,
                                                 "peak": fp64_peak, "unit": "TFLOP/s",
                                                 "frac": flops_a / (ms * 1e-3) / 1e12 / fp64_peak,
                                                 "note": "executed flops (the kernel replaces SURVEY 8d's 68 nt "
                                                         "transcendental flops per frequency by a recurrence); "
                                                         "includes the inverse FFT and the transpose"}}
        # the same path from the model PARAMETERS: construction + device maps + evaluate_ybar, wall clock
        import popkinmocks_b200 as pkm_
        import bench_workloads as W_

        def model_wall(make):
            make(cube).evaluate_ybar(device=local)            # warm (tables cached on the cube)
            torch.cuda.synchronize()
            walls = []
            for _ in range(3):                                # median of three (page-locked result buffers recycle)
                t0_ = time.perf_counter()
                c_ = make(cube)
                c_.evaluate_ybar(device=local)
                walls.append(1e3 * (time.perf_counter() - t0_))
                del c_
            return statistics.median(walls)
        other["growing_disk_setup_plus_evaluate_ybar"] = {
            "ms": model_wall(W_.make_disk), "note": "GrowingDisk(cube) + set_p_t/set_p_x_t/set_t_dep/set_mu_v/"
            "set_sig_v + evaluate_ybar() wall clock: the O(pixels) maps are generated in HBM (pkm_disk_maps, "
            "pkm_assemble_txz); includes the D2H of the 1.6 GB cube into a pageable NumPy array",
            "ref": "growing_disk.py:30-239, parametric.py:335-407"}
        other["stream_setup_plus_evaluate_ybar"] = {
            "ms": model_wall(W_.make_stream), "ref": "stream.py:28-138, parametric.py:335-407"}
        # analytic pixelation of the same Gaussian component into a second 5-D device array
        log_w = torch.log(P)
        pix = torch.empty_like(logp)
        ms = timed(lambda: engine.pixelate_gaussian(log_w, mu_v, sig_v, cube.v_edg, device=local, out=pix))
        other["pixelate_gaussian"] = {"ms": ms, "GB_per_s": pix.numel() * 8 / (ms * 1e-3) / 1e9,
                                      "roofline": hbm_roof(pix.numel() * 8, ms),
                                      "note": "pkm_pixelate_gaussian: writes log p(t,v,x,z) once (HBM-write bound)",
                                      "ref": "ParametricComponent._get_log_p_tvxz parametric.py:567-596"}
        del P, mu_v, sig_v, log_w, pix
        plan32 = engine.Plan(cube, device=local, precision="float32")
        P32 = torch.rand((nt, rows, nxl, nz), dtype=torch.float64, device=dev, generator=gen)
        P32 /= P32.sum()
        mu32 = 400.0 * (torch.rand((nt, rows, nxl), dtype=torch.float64, device=dev, generator=gen) - 0.5)
        sg32 = 20.0 + 200.0 * torch.rand((nt, rows, nxl), dtype=torch.float64, device=dev, generator=gen)
        ms = timed(lambda: plan32.ybar_gaussian(P32, mu32, sg32, out=out))
        y32 = out.clone()
        plan.ybar_gaussian(P32, mu32, sg32, out=out)
        other["gaussian_path_A_float32"] = {
            "voxel_per_s": voxels_per_step / (ms * 1e-3), "ms": ms,
            "max_rel_diff_vs_float64": float((y32 - out).abs().max() / out.abs().max()),
            "note": "optional float32 arithmetic of the Gaussian path (k_gaussian_f32), tolerance 1e-5; the inverse FFT stays float64"}
        del P32, mu32, sg32, y32
        ms = timed(lambda: plan32.ybar_density(logp, is_log=True, out=out))
        other["density_path_B_float32"] = {"voxel_per_s": voxels_per_step / (ms * 1e-3), "ms": ms,
                                           "note": "optional float32 contraction, tolerance 1e-5"}
        plan32.close()
        del plan32
        log_dv = np.log(np.diff(cube.v_edg))
        for keep, ref in (("vx", "Component._get_log_p_vx base.py:1062-1079"),
                          ("tz", "Component._get_log_p_tz base.py:941-960")):
            ms = timed(lambda: plan.reduce_logp(logp, log_dv, keep, False))
            other["log_marginal_" + keep] = {
                "ms": ms, "GB_per_s": logp.numel() * 8 / (ms * 1e-3) / 1e9,
                "roofline": hbm_roof(logp.numel() * 8, ms),
                "note": "pkm_reduce_logp: one streaming pass (online log-sum-exp) over the 5-D array; "
                        "one table exponential per element: issue/FP64 bound at about half the HBM rate",
                "ref": ref}
        ms = timed(lambda: engine.noise(out, 100.0, 0, seed=1, want_sigma=False))
        other["shot_noise"] = {"voxel_per_s": voxels_per_step / (ms * 1e-3), "ms": ms,
                               "roofline": hbm_roof(3 * out.numel() * 8, ms),   # max pass + read + write
                               "ref": "ShotNoise.get_noisy_data noise.py:16-71"}
        # the same cube over the full MILES wavelength range (3540-7410 A): n_fft = 11179 = 7*1597,
        # 2.3x more voxels per pixel; the inverse FFT no longer fits shared memory (M = 32768)
        import popkinmocks_b200 as pkm
        ssps_full = pkm.milesSSPs(lmd_min=3540.5, lmd_max=7409.6)
        cube_full = pkm.ifu_cube.IFUCube(ssps=ssps_full, nx1=nx, nx2=nx, nv=nv, x1rng=(-1.0, 1.0), x2rng=(-1.0, 1.0),
                                         vrng=(-1000.0, 1000.0))
        plan_full = engine.Plan(cube_full, device=local)
        out_full = torch.empty((plan_full.n_fft, rows, nxl), dtype=torch.float64, device=dev)
        ms = timed(lambda: plan_full.ybar_density(logp, is_log=True, out=out_full), reps=2)
        other["density_path_B_full_wavelength_range"] = {
            "voxel_per_s": total_pix * plan_full.n_fft / (ms * 1e-3), "ms": ms, "n_fft": plan_full.n_fft,
            "note": "milesSSPs(lmd_min=3540.5, lmd_max=7409.6): the 'full-resolution MILES' variant of SURVEY.md 8d C4"}
        plan_full.close()
        del plan_full, out_full
        torch.cuda.empty_cache()

    # ---- e2e: same call through the C ABI with pinned host buffers
    e2e = None
    windows = [(w0, w1)]
    host_bytes = world * (logp.numel() * 8 + n_fft * rows * nxl * 8)
    try:
        import psutil
        ram = psutil.virtual_memory().total
    except Exception:  # noqa: BLE001
        ram = 1 << 40
    if not args.no_e2e and host_bytes > 0.6 * ram:
        e2e = {"value": None, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0,
               "note": f"skipped: pinned host buffers of all ranks ({host_bytes / 1e9:.0f} GB) exceed 60% of host RAM"}
    elif not args.no_e2e:
        # the call a user of the reference makes: Component(cube, log_p_tvxz).evaluate_ybar() on a
        # HOST array (page-locked in place by the class on first use), result = a host NumPy cube
        import popkinmocks_b200 as pkm
        # Multi-GPU: this leg is bound by host-to-device copies, and the GPUs of one box do not all
        # see the same host bandwidth (tools/h2d_probe.py: PCIe switches shared by 4 GPUs).  The
        # pixel ranges of the e2e leg are therefore sized in proportion to each rank's MEASURED
        # concurrent H2D rate (256 MB probe), so that all ranks finish their copies together.
        e0_, e1_ = r0, r1
        h2d_rates = None
        if world > 1 and args.scaling == "strong" and not os.environ.get("PKM_BENCH_EQUAL_E2E"):
            pb = torch.empty(1 << 25, dtype=torch.float64, pin_memory=True)
            db = torch.empty(1 << 25, dtype=torch.float64, device=dev)
            db.copy_(pb, non_blocking=True)
            barrier()
            tp0 = time.perf_counter()
            for _ in range(4):
                db.copy_(pb, non_blocking=True)
            torch.cuda.synchronize()
            rate = torch.tensor([4 * pb.numel() * 8 / (time.perf_counter() - tp0) / 1e9], dtype=torch.float64, device=dev)
            allr = [torch.zeros_like(rate) for _ in range(world)]
            dist.all_gather(allr, rate)
            h2d_rates = [float(a) for a in allr]
            cum = np.concatenate([[0.0], np.cumsum(h2d_rates)]) / sum(h2d_rates)
            bounds = [int(round(total_pix * c)) for c in cum]
            e0_, e1_ = bounds[rank], bounds[rank + 1]
            del pb, db
        del logp
        torch.cuda.empty_cache()
        rows_e = e1_ - e0_ if world > 1 else rows
        shard = engine.synth_log_density(pix0=e0_, npix=(e1_ - e0_) if world > 1 else npix, device=local, **synth)
        shard = shard.view(nt, nv, rows_e, nxl, nz)
        h_in = torch.empty(shard.shape, dtype=torch.float64, pin_memory=True)
        h_in.copy_(shard)
        np_in = h_in.numpy()
        del shard
        torch.cuda.empty_cache()
        cube_e2e = cube if world == 1 else build_cube(args, nx1=rows_e, nx2=nxl)
        cube_e2e.dx, cube_e2e.dy = cube.dx, cube.dy     # the shard's pixels ARE the global cube's pixels
        cmp = pkm.components.Component(cube=cube_e2e, log_p_tvxz=np_in)

        def step_e2e():
            cmp.evaluate_ybar(device=local)

        step_e2e()
        barrier()
        e2e_err = None
        if spot_ref is not None and spot_px:
            # the class-API cube must be the cube verified above (rank 0's pixel range of it)
            keep = [i for i, p in enumerate(spot_px) if e0_ <= p < e1_] if world > 1 else list(range(len(spot_px)))
            if keep:
                got = cmp.ybar.reshape(n_fft, -1)[:, [spot_px[i] - (e0_ if world > 1 else r0) for i in keep]]
                ref_ = spot_ref[:, keep]
                e2e_err = float(np.max(np.abs(got - ref_)) / np.max(np.abs(ref_)))
        w2 = sampler.mark() if sampler else 0
        t0 = time.perf_counter()
        for _ in range(args.e2e_steps):
            step_e2e()
        torch.cuda.synchronize()
        dt = time.perf_counter() - t0
        w3 = sampler.mark() if sampler else 0
        windows.append((w2, w3))
        if world > 1:
            t = torch.tensor([dt], dtype=torch.float64, device=dev)
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
            dt = float(t.item())
        e2e = {"value": voxels_per_step * args.e2e_steps / dt, "unit": UNIT,
               "h2d_bytes_per_step": int(voxels_per_step // n_fft * nt * nv * nz * 8) if world > 1 else int(h_in.numel() * 8),
               "d2h_bytes_per_step": int(voxels_per_step * 8) if world > 1 else int(cmp.ybar.size * 8),
               "h2d_GBps_per_rank_probe": h2d_rates,
               "pixels_per_rank": ("proportional to the probed H2D rates" if h2d_rates else "equal"),
               "steps": args.e2e_steps, "ms_per_step": 1e3 * dt / args.e2e_steps,
               "api": "popkinmocks_b200.components.Component(cube, log_p_tvxz).evaluate_ybar()",
               "max_rel_diff_vs_device_cube": e2e_err,
               "note": "per rank: host log_p_tvxz (page-locked in place) -> Component.evaluate_ybar -> host "
                       "NumPy cube; H2D of tile i+1, kernels of tile i and D2H of tile i-1 overlap inside "
                       "pkm_ybar_density; multi-GPU ranks leave their pixel ranges on their own host buffers; "
                       "byte counts are whole-job totals"}

    clocks = sampler.stop(windows) if sampler else None

    if rank == 0:
        # ---- roofline of the dominant kernel (live CUDA-event times of this run)
        nfo, ntz, nfi = n_fft // 2 + 1, nt * nz, nv // 2 + 1
        # executed FP64 work per pixel: contraction = per (k, t, z) 6 FMA for the 3-weight spline
        # form + 4 FMA complex MAC (20 flops) + 6 DADD per (t,z) shared by the 25 frequencies of a warp;
        # coefficient GEMM = nfi x nv real MACs per (t, z) column
        flops_px = {"contract": (20.0 + 6.0 / 25.0) * nfo * ntz, "coeff": 2.0 * nfi * nv * ntz}
        tot_k = sum(m for m, _ in ktimes.values()) or 1.0
        kern = {k: {"ms_per_step": m / args.steps, "launches_per_step": c / args.steps, "share": m / tot_k}
                for k, (m, c) in ktimes.items() if c}
        dom = max(kern, key=lambda k: kern[k]["ms_per_step"])
        roof = None
        if dom in flops_px:
            ach = flops_px[dom] * npix / (kern[dom]["ms_per_step"] * 1e-3) / 1e12
            roof = {"bound": "fp64", "kernel": "k_" + dom, "achieved": ach, "peak": fp64_peak,
                    "unit": "TFLOP/s", "frac": ach / fp64_peak, "traffic": ncu_traffic(dom, npix),
                    "peak_source": "pkm_fp64_fma_peak measured in this run (MEASURED_PEAKS.json has no FP64 figure)",
                    "flops_per_pixel": flops_px[dom],
                    "note": "P-hat is never materialised (SURVEY.md 8f rank 1): the contraction is "
                            "FP64-pipe bound, not HBM bound; see roofline_hbm for the canonical byte figure. "
                            "peak is the 2-register-operand DFMA rate; DFMAs with 3 fresh register "
                            "operands (this kernel's mix) measure 25.5 TFLOP/s (tools/dfma_rf.cu)"}
        peaks = {}
        try:
            peaks = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))
        except Exception:  # noqa: BLE001
            pass
        hbm_peak = float(peaks.get("hbm_gbs", 6650.0))
        b_vox = (8.0 * nt * nv * nz + 2 * 16.0 * nfo * ntz + 8.0 * n_fft) / n_fft
        b_min = (8.0 * nt * nv * nz + 8.0 * n_fft) / n_fft
        per_gpu = value / world
        line = {
            "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps,
            "warmup": max(args.warmup, 3), "ms_per_step": ms_step, "higher_is_better": True,
            "scaling": args.scaling, "vs_baseline": None, "dtype": "f64",
            "data": "synthetic", "config": workload_config(args, cube),
            "e2e": e2e, "gpu_launches": int(launches),
            "roofline": roof,
            "roofline_hbm": {"bound": "hbm", "achieved": b_vox * per_gpu / 1e9, "peak": hbm_peak,
                             "unit": "GB/s", "frac": b_vox * per_gpu / 1e9 / hbm_peak,
                             "bytes_per_voxel": b_vox, "note": B_VOX_NOTE,
                             "compulsory": {"bytes_per_voxel": b_min, "achieved": b_min * per_gpu / 1e9,
                                            "frac": b_min * per_gpu / 1e9 / hbm_peak},
                             "peak_source": "MEASURED_PEAKS.json" if peaks else "fallback 6650"},
            "gather": (args.gather if world > 1 else None), "gather_note": gather_note,
            "numa_affinity": (f"{len(numa_cpus)} local cores" if numa_cpus else None),
            "verify": verify,
            "kernels": kern, "kernel_timing": kernel_timing, "per_rank": per_rank, "other_paths": other, "cpu_baseline": cpu, "clocks": clocks,
        }
        print(json.dumps(line))
    if world > 1:
        dist.barrier()
        if args.gather == "peer":
            torch.cuda.synchronize()
            out = None
            pc.close()
        dist.barrier()
        dist.destroy_process_group()
    if verify_failed:
        raise SystemExit(f"bench.py: the evaluated cube differs from the oracle by more than {VERIFY_TOL:g}")


def spot_pixels(total_pix):
    """First and last pixel of every block of an 8-way partition of the flattened pixel axis (the
    same list for every N, so the spot checks of N = 1, 2, 4, 8 look at the same pixels)."""
    from popkinmocks_b200 import sharding
    b = sharding.row_partition(total_pix, 8)
    return sorted({p for r in range(8) if b[r + 1] > b[r] for p in (b[r], b[r + 1] - 1)})


def cube_checksum(cube2d):
    """(bit checksum, float sum) of a float64 CUDA tensor.  The checksum is the sum of the IEEE-754
    bit patterns modulo 2^64 - exact, independent of summation order and of how the cube was
    sharded; it changes if any voxel changes in its last bit."""
    import torch
    bits = cube2d.contiguous().view(torch.int64).reshape(-1)
    lo = int((bits & 0xFFFFFFFF).sum().item())
    hi = int(((bits >> 32) & 0xFFFFFFFF).sum().item())
    return f"{(lo + (hi << 32)) & ((1 << 64) - 1):016x}", float(cube2d.sum().item())


def verify_cube(cube2d, cube, synth, total_pix):
    """cube2d: [n_fft, total_pix] on the device.  Returns the `verify` object of the JSON line."""
    from oracle import restatement as R
    from popkinmocks_b200 import engine
    s = cube.ssps
    px = spot_pixels(total_pix)
    logp = engine.synth_log_density_host(synth["nt"], synth["nv"], total_pix, px, synth["nz"],
                                         synth["seed"], synth["scale"])
    p = np.exp(logp)[:, :, None, :, :]                                   # [nt, nv, 1, npx, nz]
    y_ref = R.ybar_density_literal(p, cube.v_edg, s.FXw, s.par_dims, s.n_fft, s.dv, s.delta_t, s.delta_z,
                                   cube.dx, cube.dy)[:, 0, :]
    y_gpu = cube2d[:, px].cpu().numpy()
    err = float(np.max(np.abs(y_gpu - y_ref)) / np.max(np.abs(y_ref)))
    chk, total = cube_checksum(cube2d)
    return {"max_rel_err": err, "tol": VERIFY_TOL, "ok": bool(err < VERIFY_TOL), "pixels": len(px),
            "pixel_index": px, "checksum": chk, "sum": total,
            "oracle": "oracle/restatement.py ybar_density_literal (base.py:836-859) on the same seeded inputs",
            "note": "checksum = sum of the float64 bit patterns of the whole cube mod 2^64 (identical "
                    "for N = 1, 2, 4, 8 iff the assembled cubes are bit-identical)"}


def ncu_traffic(kernel, npix):
    """DRAM bytes per launch from the committed ncu capture of the same launch geometry, or None."""
    try:
        tab = json.load(open(os.path.join(ROOT, "profiles", "ncu_traffic.json")))
        ent = tab.get(kernel)
        if ent and ent.get("npix") == npix:
            return ent["dram_bytes_per_launch"]
    except Exception:  # noqa: BLE001
        pass
    return None


def run_config(args):
    """--config C2|C3|C5: the other BASELINE.json configurations through bench_workloads.py, reported
    with the same JSON contract.  N > 1 runs one independent replica per GPU (these configurations
    fit one GPU; the sharded multi-GPU path is the C4 line)."""
    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    import bench_workloads as W
    nx = {"C2": 51, "C3": 101, "C5": 101}[args.config] if args.nx == 201 else args.nx
    cpu = None
    if rank == 0 and not args.no_cpu:
        cube = W.make_cube(nx, args.nv)
        if args.config == "C2":
            arm = CpuArm(args, cube, task=_cpu_task_gaussian)
            kind = ("oracle/restatement.py ybar_gaussian (parametric.py:355-373) on random maps of the same shape")
        else:
            arm = CpuArm(args, cube)
            kind = ("oracle/restatement.py ybar_density_literal (base.py:836-859) on random densities of the same "
                    "shape - the cube evaluation, which dominates the reference's time for this configuration")
        arm.step()
        tot_s, tot_v = 0.0, 0
        while tot_s < 6.0:
            sec, vox = arm.step()
            tot_s += sec
            tot_v += vox
        arm.close()
        cpu = {"value": tot_v / tot_s, "unit": UNIT, "cores": arm.nproc, "kind": "port",
               "sample": f"{tot_v // cube.ssps.n_fft} pixels in {tot_s:.1f} s wall, {arm.nproc} workers: {kind}"}
    import torch
    import torch.distributed as dist
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)
    uuid = "GPU-" + str(torch.cuda.get_device_properties(dev).uuid).replace("GPU-", "")
    sampler = ClockSampler(uuid) if rank == 0 else None
    w0 = sampler.mark() if sampler else 0
    if world > 1:
        dist.barrier()
    r = W.RUNNERS[args.config](nx=nx, nv=args.nv, steps=args.steps, warmup=max(args.warmup, 3), device=local,
                               e2e_steps=args.e2e_steps)
    w1 = sampler.mark() if sampler else 0
    ms, e2e_ms, ok = r["ms_per_step"], r["e2e"]["ms_per_step"], 1.0 if r["verify"]["ok"] else 0.0
    if world > 1:
        t = torch.tensor([ms, e2e_ms, -ok], dtype=torch.float64, device=dev)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        ms, e2e_ms, ok = float(t[0]), float(t[1]), -float(t[2])
    clocks = sampler.stop([(w0, w1)]) if sampler else None
    if rank == 0:
        cfg = dict(r["config"])
        cfg["workload"] = r["workload"]
        e2e = dict(r["e2e"])
        e2e.update({"value": world * r["voxels"] / (e2e_ms * 1e-3), "unit": UNIT, "ms_per_step": e2e_ms})
        line = {"metric": METRIC, "value": world * r["voxels"] / (ms * 1e-3), "unit": UNIT, "n_gpus": world,
                "steps": args.steps, "warmup": max(args.warmup, 3), "ms_per_step": ms, "higher_is_better": True,
                "scaling": "weak" if world > 1 else "strong", "vs_baseline": None, "dtype": "f64", "data": "synthetic",
                "config": cfg, "e2e": e2e, "gpu_launches": int(r["launches"]), "roofline": r["roofline"],
                "verify": r["verify"], "kernels": r["kernels"], "other_paths": r["other_paths"],
                "cpu_baseline": cpu, "clocks": clocks,
                "replicas": ("one independent replica per GPU" if world > 1 else None)}
        print(json.dumps(line))
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()
    if ok < 0.5:
        raise SystemExit(f"bench.py --config {args.config}: verification against the oracle failed")


if __name__ == "__main__":
    a = parse_args()
    if a.impl == "reference":
        run_reference(a)
    elif a.config != "C4":
        run_config(a)
    else:
        run_b200(a)
