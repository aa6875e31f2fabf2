"""Sort-first split of one frame over the GPUs of a box, and the composite to the presenting GPU.

The reference exposes exactly one physical device (vulkan-driver/src/api_impl.rs:5092-5119), so the split is internal:
rank r renders the rows [y0, y1) of its band (KznFramebuffer.band_y0/band_y1 clip setup, binning and raster to the band;
geometry is replicated), then ONE exchange step moves every band into the presenting rank's image.  Bands are disjoint
sets of pixels, so seams need no blending and depth is never exchanged.

Plumbing only: torch.distributed point-to-point (NCCL over NVLink on the box, gloo in the CPU tests).
"""
from __future__ import annotations

from dataclasses import dataclass

TILE_ROWS = 8  # raster tile height (kazan_b200/csrc/raster.cu kTileH): bands start on tile rows so no tile is shared


def band_rows(height: int, world: int, rank: int, align: int = TILE_ROWS) -> tuple[int, int]:
    """Rows [y0, y1) of `rank`'s band: `world` horizontal bands of near-equal height, starts aligned to `align` rows."""
    if not (0 <= rank < world):
        raise ValueError(f"rank {rank} outside world of {world}")
    units = -(-height // align)                      # tile rows in the frame
    lo = (units * rank) // world * align
    hi = (units * (rank + 1)) // world * align
    return min(lo, height), min(hi, height)


def shard_invocations(total: int, world: int, rank: int, granule: int = 256) -> tuple[int, int]:
    """(first_invocation, count) of `rank`'s slice of a compute dispatch, in whole 256-invocation CTAs (kzn_cmd_dispatch)."""
    units = -(-total // granule)
    lo = min((units * rank) // world * granule, total)
    hi = min((units * (rank + 1)) // world * granule, total)
    return lo, hi - lo


@dataclass
class SplitFrame:
    width: int
    height: int
    world: int
    rank: int
    presenter: int = 0
    texel_bytes: int = 4

    @property
    def band(self) -> tuple[int, int]:
        return band_rows(self.height, self.world, self.rank)

    def band_of(self, rank: int) -> tuple[int, int]:
        return band_rows(self.height, self.world, rank)

    def byte_range(self, rank: int) -> tuple[int, int]:
        y0, y1 = self.band_of(rank)
        pitch = self.width * self.texel_bytes
        return y0 * pitch, y1 * pitch

    def composite_bytes(self) -> int:
        """Bytes that cross the links in one composite (everything but the presenter's own band)."""
        return sum(b - a for r in range(self.world) if r != self.presenter for a, b in [self.byte_range(r)])

    def composite(self, frame, dist) -> None:
        """`frame`: flat uint8 tensor of the whole LINEAR image on every rank (only the own band holds valid pixels).
        After the call the presenter's tensor holds every band.  One grouped send/recv step; returns when complete."""
        if self.world == 1:
            return
        ops = []
        if self.rank == self.presenter:
            for r in range(self.world):
                if r == self.presenter:
                    continue
                a, b = self.byte_range(r)
                if b > a:
                    ops.append(dist.P2POp(dist.irecv, frame[a:b], r))
        else:
            a, b = self.byte_range(self.rank)
            if b > a:
                ops.append(dist.P2POp(dist.isend, frame[a:b], self.presenter))
        if ops:
            for w in dist.batch_isend_irecv(ops):
                w.wait()


class PeerSplitFrame:
    """Split frame whose render targets live on the presenting GPU and are written by every rank over NVLink peer memory.

    The presenter allocates the colour targets and one flag block and exports them (kzn_memory_export); the other ranks
    import them (kzn_memory_import) and bind the imported image as their colour attachment, so the raster kernels' tile
    epilogue stores the band directly into the presenter's HBM.  Per frame:
        band rank : [wait credit] -> draw band -> signal `arrived` (+1, after a system-scope fence)
        presenter : draw own band -> wait `arrived` >= (world-1) * frames -> frame complete -> signal each rank's `credit`
    `credit` bounds how far a band rank may run ahead of the presenter (the number of rotating targets), so a target is
    never overwritten before the presenter has seen the frame that used it.  No collective is on the data path;
    torch.distributed only carries the 64-byte handles once at start-up.
    """

    FLAG_BYTES = 256

    def __init__(self, device, queue, width: int, height: int, world: int, rank: int, dist, n_targets: int = 2, presenter: int = 0):
        self.split = SplitFrame(width, height, world, rank, presenter)
        self.device, self.queue, self.dist = device, queue, dist
        self.world, self.rank, self.presenter, self.n_targets = world, rank, presenter, n_targets
        self.frame_bytes = width * height * 4
        self.frames = 0
        self._owned, self._imported = [], []
        is_presenter = rank == presenter
        # my own flag block: presenter -> {arrived @0, status @64}; band rank -> {credit @0, status @64}
        self.flags = device.allocate(self.FLAG_BYTES)
        self._owned.append(self.flags)
        queue.cmd_fill_buffer(self.flags.address, self.FLAG_BYTES, 0, self.FLAG_BYTES, 0)
        if is_presenter:
            self.targets = [device.allocate(self.frame_bytes) for _ in range(n_targets)]
            self._owned += self.targets
            for t in self.targets:
                queue.cmd_fill_buffer(t.address, self.frame_bytes, 0, self.frame_bytes, 0)
        queue.wait_idle()
        mine = {"flags": self.flags.export(), "targets": [t.export() for t in self.targets] if is_presenter else []}
        if world > 1:
            everyone = [None] * world
            dist.all_gather_object(everyone, mine)
        else:
            everyone = [mine]
        if is_presenter:
            self.target_addresses = [t.address for t in self.targets]
            self.arrived_address = self.flags.address
            self.credit_addresses = {}
            for r in range(world):
                if r != presenter:
                    m = device.import_memory(everyone[r]["flags"], self.FLAG_BYTES)
                    self._imported.append(m)
                    self.credit_addresses[r] = m.address
        else:
            self.target_addresses = []
            for h in everyone[presenter]["targets"]:
                m = device.import_memory(h, self.frame_bytes)
                self._imported.append(m)
                self.target_addresses.append(m.address)
            m = device.import_memory(everyone[presenter]["flags"], self.FLAG_BYTES)
            self._imported.append(m)
            self.arrived_address = m.address
        self.status_address = self.flags.address + 64
        if world > 1:
            dist.barrier()

    def target(self, frame: int | None = None) -> int:
        """Address, in THIS rank's address space, of the colour target used by frame number `frame` (default: the next one)."""
        return self.target_addresses[(self.frames if frame is None else frame) % self.n_targets]

    def begin_frame(self):
        """Band ranks: do not start frame f before the presenter has completed frame f - n_targets."""
        if self.rank != self.presenter and self.frames >= self.n_targets:
            self.queue.cmd_wait_flag(self.flags.address, self.frames - self.n_targets + 1, self.status_address)

    def end_frame(self):
        """After the band's draws were enqueued: publish the band / collect all bands."""
        self.frames += 1
        if self.world == 1:
            return
        if self.rank != self.presenter:
            self.queue.cmd_signal_flag(self.arrived_address, 1)
        else:
            self.queue.cmd_wait_flag(self.arrived_address, (self.world - 1) * self.frames, self.status_address)
            for addr in self.credit_addresses.values():
                self.queue.cmd_signal_flag(addr, 1)

    def check(self):
        """Raise if any wait on this rank gave up (a peer stopped making progress)."""
        import ctypes as C
        word = (C.c_uint32 * 1)()
        self.queue.wait_idle()
        self.queue.download(C.addressof(word), self.status_address, 4)
        self.queue.wait_idle()
        if word[0]:
            raise RuntimeError(f"rank {self.rank}: a split-frame flag wait timed out")

    def close(self):
        self.queue.wait_idle()
        if self.world > 1:
            self.dist.barrier()  # nobody frees memory a peer may still be writing
        for m in self._imported:
            m.free()
        if self.world > 1:
            self.dist.barrier()
        for m in self._owned:
            m.free()
        self._imported, self._owned = [], []
