"""Sort-first split of one frame over the GPUs of a box, and the composite to the presenting GPU.

The reference exposes exactly one physical device (vulkan-driver/src/api_impl.rs:5092-5119), so the split is internal:
rank r renders the rows [y0, y1) of its band (KznFramebuffer.band_y0/band_y1 clip setup, binning and raster to the band;
geometry is replicated), then ONE exchange step moves every band into the presenting rank's image.  Bands are disjoint
sets of pixels, so seams need no blending and depth is never exchanged.

Plumbing only: torch.distributed point-to-point (NCCL over NVLink on the box, gloo in the CPU tests).
"""
from __future__ import annotations

from dataclasses import dataclass

TILE_ROWS = 16  # height of a raster brick (kazan_b200/csrc/raster.cu kBrick): bands start on brick rows, so no warp's pixel block is shared between ranks


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
    """Split frame whose render targets live on the presenting GPU and are filled by every rank over NVLink peer memory.

    The presenter allocates the colour targets and one flag block and exports them (kzn_memory_export); the other ranks
    import them (kzn_memory_import).  Two ways to move a band (`mode`):
      "push"  (default) the band rank renders into a local image and its COPY QUEUE pushes the band into the presenter's
              image with the copy engines (kzn_queue_copy_dma) while the SMs already render the next frame; semaphores
              order the two queues, two local images rotate.  No SM time is spent on the transfer.
      "store" the band rank binds the imported image as its colour attachment, so the raster kernels' tile epilogue stores
              the band directly into the presenter's HBM (no copy at all).  Best at 2 GPUs; with more writers the remote
              stores back-pressure the raster kernels (measured: profiles/).
    Per frame:
        band rank : [wait credit] -> draw band -> (push) -> signal ITS `arrived` flag (+1, after a system-scope fence)
        presenter : draw own band; on its PRESENT queue: wait for that draw and until EVERY band rank's `arrived` flag >= frames
                    -> frame complete -> signal each rank's `credit`.  Its render queue goes on with the next frame meanwhile
                    (another target) and only waits for "frame complete" of the frame that used the same target, so the
                    presenter's SMs never idle while the other ranks' bands are on the wire.
    (One arrival flag per band rank: with one aggregated counter a rank running a frame ahead would be counted for a rank
    that is a frame behind, and the presenter would declare a frame complete with a band missing.)
    `credit` bounds how far a band rank may run ahead of the presenter (the number of rotating targets), so a target is
    never overwritten before the presenter has seen the frame that used it.  No collective is on the data path;
    torch.distributed only carries the 64-byte handles once at start-up.
    """

    FLAG_BYTES = 256
    LOCAL_IMAGES = 2

    def __init__(self, device, queue, width: int, height: int, world: int, rank: int, dist, n_targets: int = 2, presenter: int = 0,
                 mode: str = "push"):
        from .host import Semaphore
        if mode not in ("push", "store"):
            raise ValueError(mode)
        self.split = SplitFrame(width, height, world, rank, presenter)
        self.device, self.queue, self.dist, self.mode = device, queue, dist, mode
        self.world, self.rank, self.presenter, self.n_targets = world, rank, presenter, n_targets
        self.frame_bytes = width * height * 4
        self.frames = 0
        self._owned, self._imported = [], []
        is_presenter = rank == presenter
        # my own flag block: presenter -> {arrived[band rank] @ 8 * index, status @192}; band rank -> {credit @0, status @192}
        self.flags = device.allocate(self.FLAG_BYTES)
        self._owned.append(self.flags)
        queue.cmd_fill_buffer(self.flags.address, self.FLAG_BYTES, 0, self.FLAG_BYTES, 0)
        self.targets = []
        if is_presenter:
            self.targets = [device.allocate(self.frame_bytes) for _ in range(n_targets)]
            self._owned += self.targets
            for t in self.targets:
                queue.cmd_fill_buffer(t.address, self.frame_bytes, 0, self.frame_bytes, 0)
        self.copy_queue, self.local, self.rendered, self.copied = None, [], [], []
        self.present_queue, self.completed = None, []
        if is_presenter and world > 1:
            self.present_queue = device.create_queue()
            self.rendered = [Semaphore(device) for _ in range(n_targets)]
            self.completed = [Semaphore(device) for _ in range(n_targets)]
        if not is_presenter and mode == "push":
            self.copy_queue = device.create_queue()
            self.local = [device.allocate(self.frame_bytes) for _ in range(self.LOCAL_IMAGES)]
            self._owned += self.local
            self.rendered = [Semaphore(device) for _ in range(self.LOCAL_IMAGES)]
            self.copied = [Semaphore(device) for _ in range(self.LOCAL_IMAGES)]
        queue.wait_idle()
        mine = {"flags": self.flags.export(), "targets": [t.export() for t in self.targets]}
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
            self.arrived_address = m.address + 8 * self.band_index(rank)  # this rank's own arrival flag
        self.status_address = self.flags.address + 192
        if world > 1:
            dist.barrier()

    def band_index(self, rank: int) -> int:
        """position of `rank` among the band ranks (everyone but the presenter): the index of its arrival flag"""
        return rank if rank < self.presenter else rank - 1

    @property
    def pushes(self) -> bool:
        return self.copy_queue is not None

    def render_addresses(self) -> list[int]:
        """Every address begin_frame() can return on this rank (to prepare framebuffer descriptions once)."""
        return [m.address for m in self.local] if self.pushes else list(self.target_addresses)

    def target(self, frame: int | None = None) -> int:
        """Address, in THIS rank's address space, of the presenter's colour target used by frame number `frame`."""
        return self.target_addresses[(self.frames if frame is None else frame) % self.n_targets]

    def begin_frame(self) -> int:
        """Returns the address of the image this rank renders the next frame's band into."""
        f = self.frames
        if self.pushes:
            slot = f % self.LOCAL_IMAGES
            if f >= self.LOCAL_IMAGES:
                self.queue.wait_semaphore(self.copied[slot])  # the push of frame f - 2 has left this image
            return self.local[slot].address
        if self.rank != self.presenter and f >= self.n_targets:
            # do not start frame f before the presenter has completed frame f - n_targets
            self.queue.cmd_wait_flag(self.flags.address, f - self.n_targets + 1, self.status_address)
        if self.present_queue and f >= self.n_targets:
            self.queue.wait_semaphore(self.completed[f % self.n_targets])  # the frame that used this target is complete
        return self.target(f)

    def end_frame(self):
        """After the band's draws were enqueued: publish the band / collect all bands."""
        f = self.frames
        self.frames += 1
        if self.world == 1:
            return
        if self.rank == self.presenter:
            slot, pq = f % self.n_targets, self.present_queue
            self.queue.signal_semaphore(self.rendered[slot])
            pq.wait_semaphore(self.rendered[slot])
            pq.cmd_wait_flags(self.arrived_address, self.world - 1, self.frames, self.status_address)
            for addr in self.credit_addresses.values():
                pq.cmd_signal_flag(addr, 1)
            pq.signal_semaphore(self.completed[slot])
        elif self.pushes:
            slot = f % self.LOCAL_IMAGES
            a, b = self.split.byte_range(self.rank)
            cq = self.copy_queue
            self.queue.signal_semaphore(self.rendered[slot])
            cq.wait_semaphore(self.rendered[slot])
            if f >= self.n_targets:
                cq.cmd_wait_flag(self.flags.address, f - self.n_targets + 1, self.status_address)
            cq.copy_dma(self.target(f) + a, self.local[slot].address + a, b - a)
            cq.cmd_signal_flag(self.arrived_address, 1)
            cq.signal_semaphore(self.copied[slot])
        else:
            self.queue.cmd_signal_flag(self.arrived_address, 1)

    def wait_idle(self):
        self.queue.wait_idle()
        if self.copy_queue:
            self.copy_queue.wait_idle()
        if self.present_queue:
            self.present_queue.wait_idle()

    def check(self):
        """Raise if any wait on this rank gave up (a peer stopped making progress)."""
        import ctypes as C
        word = (C.c_uint32 * 1)()
        self.wait_idle()
        self.queue.download(C.addressof(word), self.status_address, 4)
        self.queue.wait_idle()
        if word[0]:
            raise RuntimeError(f"rank {self.rank}: a split-frame flag wait timed out")

    def close(self):
        self.wait_idle()
        if self.world > 1:
            self.dist.barrier()  # nobody frees memory a peer may still be writing
        for m in self._imported:
            m.free()
        if self.world > 1:
            self.dist.barrier()
        for m in self._owned:
            m.free()
        for sem in self.rendered + self.copied + self.completed:
            sem.destroy()
        if self.copy_queue:
            self.copy_queue.destroy()
        if self.present_queue:
            self.present_queue.destroy()
        self._imported, self._owned, self.rendered, self.copied, self.copy_queue = [], [], [], [], None
        self.completed, self.present_queue = [], None
