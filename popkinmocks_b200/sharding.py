"""Spatial sharding of a datacube across the GPUs of one box (one process per GPU).

Every spatial pixel of `evaluate_ybar` is independent, so the cube is
partitioned into contiguous blocks of x1 rows, one per rank, with no exchange during compute.
The only collective is the final gather of the per-rank cube tiles onto rank 0 (NCCL over
NVLink on the GPU box; gloo in the CPU tests): ranks produce spaxel-major tiles
[rows*nx2, n_fft] so the gathered buffer is one contiguous [nx1*nx2, n_fft] array that a
single transpose turns into the reference layout [n_fft, nx1, nx2].
"""
import numpy as np


def row_partition(nx1, world_size):
    """Contiguous x1-row blocks: rank r owns rows [b[r], b[r+1]); sizes differ by at most 1."""
    if world_size < 1 or nx1 < 0:
        raise ValueError("need world_size >= 1 and nx1 >= 0")
    return [(r * nx1) // world_size for r in range(world_size + 1)]


def my_rows(nx1, world_size, rank):
    b = row_partition(nx1, world_size)
    return b[rank], b[rank + 1]


def bind_to_gpu_numa_node(device_index):
    """Restrict this process to the CPU cores local to GPU `device_index` (NVML's ideal CPU
    affinity), so that page-locked host buffers allocated afterwards are first-touched on the
    GPU's own NUMA node and host<->device copies do not cross the socket interconnect.
    Returns the list of CPUs, or None if NVML / the affinity call is unavailable."""
    import os
    try:
        import pynvml
        pynvml.nvmlInit()
        visible = os.environ.get("CUDA_VISIBLE_DEVICES")
        index = int(visible.split(",")[device_index]) if visible and visible.replace(",", "").isdigit() else device_index
        handle = pynvml.nvmlDeviceGetHandleByIndex(index)
        words = (os.cpu_count() + 63) // 64
        mask = pynvml.nvmlDeviceGetCpuAffinity(handle, words)
        cpus = [64 * w + b for w, m in enumerate(mask) for b in range(64) if (m >> b) & 1]
        allowed = sorted(set(cpus) & set(os.sched_getaffinity(0)))
        if not allowed:
            return None
        os.sched_setaffinity(0, allowed)
        return allowed
    except Exception:  # noqa: BLE001 - NVML missing, containers without the syscall, ...
        return None


def gather_pixel_major(tile, nx1, nx2, dist=None, dst=0):
    """Gather per-rank spaxel-major tiles [rows_r*nx2, n] on `dst`.

    Returns the [nx1*nx2, n] array on `dst` (None elsewhere).  `tile` is a torch tensor on the
    backend's device (CUDA for nccl, CPU for gloo).  With `dist=None` (single process) the
    tile is returned unchanged."""
    if dist is None or not dist.is_initialized() or dist.get_world_size() == 1:
        return tile
    import torch
    world, rank = dist.get_world_size(), dist.get_rank()
    bounds = row_partition(nx1, world)
    n = tile.shape[1]
    if tile.shape[0] != (bounds[rank + 1] - bounds[rank]) * nx2:
        raise ValueError("tile does not match this rank's row block")
    if rank == dst:
        full = torch.empty((nx1 * nx2, n), dtype=tile.dtype, device=tile.device)
        parts = [full[bounds[r] * nx2:bounds[r + 1] * nx2] for r in range(world)]
        # ranks own different row counts: point-to-point instead of a padded gather
        reqs = [dist.irecv(parts[r], src=r) for r in range(world) if r != dst and parts[r].numel()]
        parts[dst].copy_(tile)
        for q in reqs:
            q.wait()
        return full
    if tile.numel():
        dist.send(tile.contiguous(), dst=dst)
    return None


def pixel_major_to_cube(full, nx1, nx2):
    """[nx1*nx2, n] -> [n, nx1, nx2] (host/NumPy or torch, any device) - the reference layout."""
    if hasattr(full, "numpy") or type(full).__module__.startswith("torch"):
        return full.t().contiguous().reshape(full.shape[1], nx1, nx2)
    return np.ascontiguousarray(full.T).reshape(full.shape[1], nx1, nx2)


class TileGather:
    """Gather of spaxel-major cube tiles onto rank 0, overlapped with the computation.

    Each rank splits its x1-row block into `ntile` sub-blocks.  As soon as a sub-block has been
    evaluated (`post`), it is sent to rank 0 on a side stream while the next one is computed;
    rank 0 evaluates its own sub-blocks straight into the gathered buffer and receives the
    others in the background.  Sends and receives of one round are issued as one NCCL group.
    Works with gloo/CPU tensors too (no streams), which is how the CPU tests drive it.
    """

    def __init__(self, nx1, nx2, n, device, dist=None, ntile=4, dtype=None):
        import torch
        self.dist = dist if (dist is not None and dist.is_initialized() and dist.get_world_size() > 1) else None
        self.world = self.dist.get_world_size() if self.dist else 1
        self.rank = self.dist.get_rank() if self.dist else 0
        self.nx1, self.nx2, self.n = nx1, nx2, n
        self.device = torch.device(device)
        dtype = dtype or torch.float64
        self.bounds = row_partition(nx1, self.world)
        self.ntile = max(1, int(ntile))
        # global row boundaries of every rank's sub-blocks
        self.tiles = []
        for r in range(self.world):
            r0, r1 = self.bounds[r], self.bounds[r + 1]
            sub = row_partition(r1 - r0, self.ntile)
            self.tiles.append([(r0 + sub[t], r0 + sub[t + 1]) for t in range(self.ntile)])
        r0, r1 = self.bounds[self.rank], self.bounds[self.rank + 1]
        if self.rank == 0:
            self.full = torch.empty((nx1 * nx2, n), dtype=dtype, device=self.device)
            self.local = self.full[r0 * nx2:r1 * nx2]
        else:
            self.full = None
            self.local = torch.empty(((r1 - r0) * nx2, n), dtype=dtype, device=self.device)
        self.cuda = self.device.type == "cuda"
        self.comm = torch.cuda.Stream(device=self.device) if (self.cuda and self.dist) else None
        self.pending = []

    def my_tiles(self):
        """[(local_row_begin, local_row_end, out_view)] for this rank, in evaluation order."""
        r0 = self.bounds[self.rank]
        out = []
        for a, b in self.tiles[self.rank]:
            out.append((a - r0, b - r0, self.local[(a - r0) * self.nx2:(b - r0) * self.nx2]))
        return out

    def post(self, t):
        """Sub-block `t` of this rank has been enqueued on the current stream: ship it."""
        if self.dist is None:
            return
        import torch
        ops = []
        if self.rank == 0:
            for r in range(1, self.world):
                a, b = self.tiles[r][t]
                if b > a:
                    ops.append(self.dist.P2POp(self.dist.irecv, self.full[a * self.nx2:b * self.nx2], r))
        else:
            a, b = self.tiles[self.rank][t]
            if b > a:
                r0 = self.bounds[self.rank]
                ops.append(self.dist.P2POp(self.dist.isend, self.local[(a - r0) * self.nx2:(b - r0) * self.nx2], 0))
        if not ops:
            return
        if self.cuda:
            ev = torch.cuda.Event()
            ev.record()
            with torch.cuda.stream(self.comm):
                self.comm.wait_event(ev)
                self.pending += self.dist.batch_isend_irecv(ops)
        else:
            self.pending += self.dist.batch_isend_irecv(ops)

    def finish(self):
        """Wait for the transfers; returns the gathered [nx1*nx2, n] buffer on rank 0, else None."""
        import torch
        for req in self.pending:
            req.wait()
        self.pending = []
        if self.cuda and self.comm is not None:
            torch.cuda.current_stream(self.device).wait_stream(self.comm)
        return self.full


class PeerCube:
    """The full cube [n_fft, nx1, nx2] on rank 0, written by every rank through peer memory.

    Rank 0 allocates the cube and shares it over CUDA IPC; every other rank maps it and moves
    each finished pixel tile (transposed locally) into rank 0's memory with one strided 2-D
    copy-engine transfer over NVLink, overlapped with the next tile's kernels
    (pkm_ybar_density_into_cube).  There is no gather pass on rank 0; one barrier per evaluation
    orders the peer writes before rank 0 reads.  CUDA only (NCCL for the handle broadcast / barrier)."""

    def __init__(self, n_fft, nx1, nx2, device, dist):
        import ctypes as C
        import torch
        from . import _lib
        self.dist, self.nx1, self.nx2, self.n_fft = dist, nx1, nx2, n_fft
        self.rank, self.world = dist.get_rank(), dist.get_world_size()
        self.device = torch.device(device)
        self._peer_base = None
        self._own_base = None
        payload = [None, None]
        error = None
        if self.rank == 0:
            # a plain cudaMalloc allocation (torch's caching allocator sub-allocates and cannot be
            # exported portably), viewed as a torch tensor through __cuda_array_interface__
            try:
                base = C.c_void_p()
                nbytes = n_fft * nx1 * nx2 * 8
                _lib.check(_lib.load().pkm_device_alloc(self.device.index or 0, nbytes, C.byref(base)))
                self._own_base = base
                handle = (C.c_uint8 * 64)()
                _lib.check(_lib.load().pkm_ipc_export(base, handle))

                class _View:
                    __cuda_array_interface__ = {"shape": (n_fft, nx1, nx2), "typestr": "<f8",
                                                "data": (base.value, False), "version": 3, "strides": None}
                self.cube = torch.as_tensor(_View(), device=self.device)
                payload = [bytes(handle), 0]
                self.ptr = base.value
            except Exception as exc:  # noqa: BLE001 - still take part in the broadcast below
                error = exc
        else:
            self.cube = None
        dist.broadcast_object_list(payload, src=0)
        if error is not None:
            raise error
        if payload[0] is None:
            raise RuntimeError("rank 0 could not allocate / export the cube")
        if self.rank != 0:
            handle = (C.c_uint8 * 64).from_buffer_copy(payload[0])
            base = C.c_void_p()
            _lib.check(_lib.load().pkm_ipc_open(handle, self.device.index or 0, C.byref(base)))
            self._peer_base = base
            self.ptr = base.value + payload[1]

    def evaluate_rows(self, plan, dens, row_begin, row_end, cube_row0, is_log=True):
        """Rows [row_begin,row_end) of this rank's density block -> rank 0's cube at `cube_row0`."""
        import ctypes as C
        from . import _lib
        nx1_local = int(dens.shape[2])
        _lib.check(_lib.load().pkm_ybar_density_into_cube(
            plan._h, _lib.ptr(dens), int(bool(is_log)), nx1_local, self.nx2, int(row_begin), int(row_end),
            C.c_void_p(self.ptr), self.nx1, int(cube_row0), None))

    def close(self):
        from . import _lib
        if self._peer_base is not None:
            _lib.load().pkm_ipc_close(self._peer_base)
            self._peer_base = None
        if self._own_base is not None:
            self.cube = None
            _lib.load().pkm_device_free(self._own_base)
            self._own_base = None
