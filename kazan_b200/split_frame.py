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
