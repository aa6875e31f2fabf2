"""The C ABI used from plain C (examples/config0.c): BASELINE configs[0] — clear + copy of a 1024x1024 image — recorded,
submitted and fenced with the calls the reference's host would make, no Python or torch on the path.

CPU: the example compiles against include/kazan_b200.h, links against the library, and (without a CUDA device) fails
loudly with VK_ERROR_INITIALIZATION_FAILED instead of falling back to anything.  GPU: every texel matches.
"""
import json
import shutil
import subprocess
from pathlib import Path

import pytest

ROOT = Path(__file__).resolve().parent.parent
LIB_DIR = ROOT / "kazan_b200"


def _compile(tmp_path_factory, name):
    if not shutil.which("gcc"):
        pytest.skip("gcc not available")
    if not (LIB_DIR / "libkazan_b200.so").exists():
        pytest.fail("libkazan_b200.so is missing: run `python -m kazan_b200.build`")
    out = tmp_path_factory.mktemp("c_example") / name
    subprocess.run(["gcc", "-std=c11", "-O2", "-Wall", "-Werror", f"-I{ROOT / 'include'}", str(ROOT / "examples" / f"{name}.c"),
                    f"-L{LIB_DIR}", "-lkazan_b200", f"-Wl,-rpath,{LIB_DIR}", "-o", str(out)], check=True)
    return out


@pytest.fixture(scope="module")
def binary(tmp_path_factory):
    return _compile(tmp_path_factory, "config0")


@pytest.fixture(scope="module")
def split_binary(tmp_path_factory):
    return _compile(tmp_path_factory, "split_frame")


def test_example_builds_and_fails_loudly_without_a_device(binary):
    import torch
    if torch.cuda.is_available():
        pytest.skip("a CUDA device is present: covered by the gpu test")
    r = subprocess.run([str(binary)], capture_output=True, text=True)
    assert r.returncode != 0 and "ok" not in r.stdout
    assert "kzn_device_count" in r.stderr or "no CUDA device" in r.stderr


@pytest.mark.gpu
def test_example_runs_config0_on_the_gpu(binary):
    r = subprocess.run([str(binary)], capture_output=True, text=True, timeout=120)
    assert r.returncode == 0, r.stderr
    lines = r.stdout.strip().splitlines()
    assert lines[-1] == "ok"
    report = json.loads(lines[0])
    assert report["mismatching_texels"] == 0
    assert report["kernel_launches_per_submit"] == 3  # the barriers keep clear, copy in and copy out apart


def test_split_frame_example_builds_and_fails_loudly_without_a_device(split_binary):
    import torch
    if torch.cuda.is_available():
        pytest.skip("a CUDA device is present: covered by the gpu tests")
    r = subprocess.run([str(split_binary)], capture_output=True, text=True)
    assert r.returncode != 0 and "ok" not in r.stdout


@pytest.mark.gpu
def test_split_frame_through_the_c_abi_one_band(split_binary):
    """the split-frame entry points with one device: one band, no exchange — what a 1-GPU box runs"""
    r = subprocess.run([str(split_binary), "1"], capture_output=True, text=True, timeout=300)
    assert r.returncode == 0, r.stderr
    assert r.stdout.strip().splitlines()[-1] == "ok"


@pytest.mark.gpu
def test_split_frame_through_the_c_abi_over_all_gpus_of_the_box(split_binary):
    """SURVEY.md §8e behind the C ABI, no Python on the path: N bands on N GPUs in one process, composited to device 0 over
    NVLink peer memory, identical to the one-device frame (needs >= 2 GPUs: gpurun --gpus 2)"""
    import torch
    n = torch.cuda.device_count()
    if n < 2:
        pytest.skip("one GPU on this box")
    for gpus in sorted({2, min(n, 4), n}):
        r = subprocess.run([str(split_binary), str(gpus), "1920", "1080", "500", "250"], capture_output=True, text=True, timeout=600)
        assert r.returncode == 0, (gpus, r.stdout, r.stderr)
        report = json.loads(r.stdout.strip().splitlines()[0])
        assert report["mismatching_bytes"] == 0 and report["gpus"] == gpus
