"""The C-ABI boundary without a GPU: symbols, struct layouts, loud failure without a device."""
import ctypes as C
import subprocess
import sys
from pathlib import Path

import pytest

from kazan_b200 import abi, host

ROOT = Path(__file__).resolve().parent.parent
VK_INCLUDE = Path("/root/reference/external/Vulkan-Headers/include")

STRUCTS = {
    "KznOffset3D": abi.Offset3D, "KznExtent3D": abi.Extent3D, "KznOffset2D": abi.Offset2D, "KznExtent2D": abi.Extent2D,
    "KznRect2D": abi.Rect2D, "KznViewport": abi.Viewport, "KznClearColorValue": abi.ClearColorValue,
    "KznClearDepthStencilValue": abi.ClearDepthStencilValue, "KznImageSubresourceRange": abi.ImageSubresourceRange,
    "KznImageSubresourceLayers": abi.ImageSubresourceLayers, "KznBufferCopy": abi.BufferCopy,
    "KznBufferImageCopy": abi.BufferImageCopy, "KznImageCopy": abi.ImageCopy, "KznImageBlit": abi.ImageBlit,
    "KznClearRect": abi.ClearRect, "KznImageDesc": abi.ImageDesc, "KznGraphicsPipelineState": abi.GraphicsPipelineState,
    "KznVertexInput": abi.VertexInput, "KznFramebuffer": abi.Framebuffer, "KznDrawStatistics": abi.DrawStatistics,
}
# Kzn struct -> Vulkan struct it must be layout-identical to
VK_TWINS = {
    "KznOffset3D": "VkOffset3D", "KznExtent3D": "VkExtent3D", "KznOffset2D": "VkOffset2D", "KznExtent2D": "VkExtent2D",
    "KznRect2D": "VkRect2D", "KznViewport": "VkViewport", "KznClearColorValue": "VkClearColorValue",
    "KznClearDepthStencilValue": "VkClearDepthStencilValue", "KznImageSubresourceRange": "VkImageSubresourceRange",
    "KznImageSubresourceLayers": "VkImageSubresourceLayers", "KznBufferCopy": "VkBufferCopy",
    "KznBufferImageCopy": "VkBufferImageCopy", "KznImageCopy": "VkImageCopy", "KznImageBlit": "VkImageBlit",
    "KznClearRect": "VkClearRect",
}


def _fields(ct, prefix=""):
    out = []
    for name, _t in ct._fields_:
        out.append(name)
    return out


def _compile_and_run(src: str, tmp_path, extra=()):
    c = tmp_path / "probe.c"
    exe = tmp_path / "probe"
    c.write_text(src)
    subprocess.run(["gcc", "-std=c11", f"-I{ROOT / 'include'}", *extra, str(c), "-o", str(exe)], check=True)
    return subprocess.run([str(exe)], check=True, capture_output=True, text=True).stdout


def test_library_loads_and_exports_every_declared_symbol():
    lib = host.load_library()
    declared = abi.declared_symbols()
    assert len(declared) >= 40
    for name in declared:
        assert hasattr(lib, name), f"{name} is declared in include/kazan_b200.h but not exported"
    assert set(declared) == set(abi.ALL_PROTOTYPES), "abi prototypes and the headers disagree"
    assert set(abi.declared_symbols(abi.BACKEND_HEADER)) == set(abi.BACKEND_PROTOTYPES)
    assert lib.kzn_abi_version() == 1


def test_ctypes_structs_match_the_c_header(tmp_path):
    lines = ['#include <stdio.h>', '#include <stddef.h>', '#include "kazan_b200.h"', "int main(void){"]
    for cname, ct in STRUCTS.items():
        lines.append(f'printf("{cname} %zu\\n", sizeof({cname}));')
        for f in _fields(ct):
            lines.append(f'printf("{cname}.{f} %zu\\n", offsetof({cname}, {f}));')
    lines.append("return 0;}")
    out = dict(l.rsplit(" ", 1) for l in _compile_and_run("\n".join(lines), tmp_path).splitlines())
    for cname, ct in STRUCTS.items():
        assert int(out[cname]) == C.sizeof(ct), cname
        for f in _fields(ct):
            assert int(out[f"{cname}.{f}"]) == getattr(ct, f).offset, f"{cname}.{f}"
    assert C.sizeof(abi.BufferImageCopy) == 56  # SURVEY.md §0.5 probe of the vendored header


@pytest.mark.skipif(not VK_INCLUDE.exists(), reason="reference Vulkan headers not present (GPU box)")
def test_structs_are_layout_identical_to_vulkan_1_1_130(tmp_path):
    lines = ['#include <stdio.h>', '#include <stddef.h>', '#include <vulkan/vulkan_core.h>', '#include "kazan_b200.h"',
             "int main(void){", 'printf("VK_HEADER_VERSION %d\\n", VK_HEADER_VERSION);']
    for k, v in VK_TWINS.items():
        lines.append(f'printf("{k} %d\\n", (int)(sizeof({k}) == sizeof({v})));')
        for f in _fields(STRUCTS[k]):
            lines.append(f'printf("{k}.{f} %d\\n", (int)(offsetof({k},{f}) == offsetof({v},{f})));')
    consts = {
        "KZN_FORMAT_R8G8B8A8_UNORM": "VK_FORMAT_R8G8B8A8_UNORM", "KZN_FORMAT_D32_SFLOAT": "VK_FORMAT_D32_SFLOAT",
        "KZN_SUCCESS": "VK_SUCCESS", "KZN_NOT_READY": "VK_NOT_READY", "KZN_TIMEOUT": "VK_TIMEOUT",
        "KZN_ERROR_OUT_OF_DEVICE_MEMORY": "VK_ERROR_OUT_OF_DEVICE_MEMORY", "KZN_ERROR_DEVICE_LOST": "VK_ERROR_DEVICE_LOST",
        "KZN_ERROR_INITIALIZATION_FAILED": "VK_ERROR_INITIALIZATION_FAILED",
        "KZN_ERROR_FEATURE_NOT_PRESENT": "VK_ERROR_FEATURE_NOT_PRESENT",
        "KZN_ERROR_FORMAT_NOT_SUPPORTED": "VK_ERROR_FORMAT_NOT_SUPPORTED",
        "KZN_ERROR_VALIDATION_FAILED": "VK_ERROR_VALIDATION_FAILED_EXT",
        "KZN_IMAGE_ASPECT_COLOR_BIT": "VK_IMAGE_ASPECT_COLOR_BIT", "KZN_IMAGE_ASPECT_DEPTH_BIT": "VK_IMAGE_ASPECT_DEPTH_BIT",
        "KZN_FILTER_LINEAR": "VK_FILTER_LINEAR", "KZN_INDEX_TYPE_UINT32": "VK_INDEX_TYPE_UINT32",
        "KZN_WHOLE_SIZE": "VK_WHOLE_SIZE", "KZN_REMAINING_ARRAY_LAYERS": "VK_REMAINING_ARRAY_LAYERS",
    }
    for k, v in consts.items():
        lines.append(f'printf("{k} %d\\n", (int)((long long){k} == (long long){v}));')
    lines.append("return 0;}")
    out = dict(l.rsplit(" ", 1) for l in _compile_and_run("\n".join(lines), tmp_path, [f"-I{VK_INCLUDE}"]).splitlines())
    assert out.pop("VK_HEADER_VERSION") == "130"
    bad = [k for k, v in out.items() if v != "1"]
    assert not bad, bad
    # enum values mirrored in Python
    assert (abi.COMPARE_LESS, abi.COMPARE_ALWAYS, abi.BLEND_SRC_ALPHA, abi.BLEND_ONE_MINUS_SRC_ALPHA) == (1, 7, 6, 7)


def test_no_cpu_fallback_without_a_device():
    """Without a CUDA device the product refuses to run instead of falling back."""
    if host.Device.count() > 0:
        pytest.skip("a CUDA device is present")
    with pytest.raises(host.KazanError) as e:
        host.Device(0)
    assert e.value.code == abi.ERROR_INITIALIZATION_FAILED


def test_product_never_imports_the_oracle():
    pkg = ROOT / "kazan_b200"
    for p in list(pkg.rglob("*.py")) + list(pkg.rglob("*.cu")) + list(pkg.rglob("*.cpp")) + list(pkg.rglob("*.h")):
        text = p.read_text()
        assert "kzn_oracle" not in text and "kzo_" not in text and "oracle." not in text, p
    out = subprocess.run(["nm", "-D", str(host.library_path())], capture_output=True, text=True, check=True).stdout
    assert "kzo_" not in out


def test_image_memory_requirements_follow_image_rs():
    # vulkan-driver/src/image.rs:58-88: size = 4*w*h*d*layers rounded up to IMAGE_ALIGNMENT (64)
    for w, h, d, l in [(1024, 1024, 1, 1), (3, 5, 1, 1), (7, 9, 3, 2), (3840, 2160, 1, 1)]:
        size, align = host.Image(0, w, h, d, l).memory_requirements()
        assert align == 64 and size == (4 * w * h * d * l + 63) // 64 * 64
    size, _ = host.Image(0, 50, 37, 1, 2, tiling=abi.TILING_TILED).memory_requirements()
    assert size == 4 * 3 * 16 * 16 * 4 * 2
    with pytest.raises(host.KazanError) as e:
        host.Image(0, 4, 4, format=44).memory_requirements()  # B8G8R8A8: unimplemented! in the reference
    assert e.value.code == abi.ERROR_FORMAT_NOT_SUPPORTED
    with pytest.raises(host.KazanError) as e:
        host.Image(0, 4, 4, mip_levels=2).memory_requirements()
    assert e.value.code == abi.ERROR_FEATURE_NOT_PRESENT


def test_image_memory_requirements_reject_sizes_that_overflow():
    """No GPU needed: kzn_image_memory_requirements is host arithmetic (image.rs:58-88 + checked_mul :74-82)."""
    from kazan_b200 import host
    for w, h, d, layers, tiling in [(2**32 - 1, 2**32 - 1, 1, 1, 0), (2**31, 2**31, 4, 4, 0), (2**32 - 1, 2**32 - 1, 1, 1, 1)]:
        with __import__("pytest").raises(host.KazanError) as e:
            host.Image(0x1000, w, h, d, layers, abi.FORMAT_R8G8B8A8_UNORM, tiling).memory_requirements()
        assert e.value.code == abi.ERROR_VALIDATION_FAILED
    assert host.Image(0x1000, 1024, 1024).memory_requirements() == (1024 * 1024 * 4, 64)
    assert host.Image(0x1000, 50, 37, tiling=abi.TILING_TILED).memory_requirements() == (4 * 3 * 1024, 64)
