"""Regenerate tests/golden/oracle_golden.json: SHA-256 of the oracle's output for every case of tests/cases.py.

    python tests/golden/make_golden.py

The reference holds no golden vectors for transfer / raster results ("parity unpinned"), so
these digests pin the ORACLE against accidental change; the GPU parity tests compare full buffers with the
oracle at run time and, as a second check that needs no oracle, with these digests.
"""
from __future__ import annotations

import hashlib
import json
import sys
from pathlib import Path

ROOT = Path(__file__).resolve().parents[2]
sys.path.insert(0, str(ROOT))

from tests import cases, runners  # noqa: E402

OUT = Path(__file__).resolve().parent / "oracle_golden.json"


def digest(b: bytes) -> str:
    return hashlib.sha256(b).hexdigest()


def generate() -> dict:
    g: dict[str, str] = {}
    for name, spec, ranges in cases.CLEAR_CASES:
        g[f"clear/{name}"] = digest(runners.oracle_clear(spec, ranges))
    for name, spec, regions, nbytes in cases.BUFFER_IMAGE_CASES:
        g[f"b2i/{name}"] = digest(runners.oracle_buffer_to_image(spec, regions, nbytes))
        g[f"i2b/{name}"] = digest(runners.oracle_image_to_buffer(spec, regions, nbytes))
    for name, s, d, regions, filt in cases.blit_cases():
        g[f"blit/{name}"] = digest(runners.oracle_blit(s, d, regions, filt))
    for name, s, d, regions in cases.copy_image_cases():
        g[f"copy_image/{name}"] = digest(runners.oracle_copy_image(s, d, regions))
    for name, spec in cases.CONVERT_TILING_CASES:
        g[f"to_tiled/{name}"] = digest(runners.oracle_convert_tiling(spec, True))
        g[f"to_linear/{name}"] = digest(runners.oracle_convert_tiling(spec, False))
    for sc in cases.draw_scenes():
        color, depth, stats = sc.run_oracle()
        g[f"draw/{sc.name}/color"] = digest(color.tobytes())
        g[f"draw/{sc.name}/depth"] = digest(depth.tobytes())
        g[f"draw/{sc.name}/stats"] = f"{stats.triangles_in},{stats.triangles_culled},{stats.fragments_covered},{stats.fragments_depth_passed}"
    return g


if __name__ == "__main__":
    g = generate()
    OUT.write_text(json.dumps(g, indent=1, sort_keys=True) + "\n")
    print(f"wrote {len(g)} digests to {OUT}")
