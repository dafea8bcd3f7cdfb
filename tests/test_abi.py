"""The C-ABI library loads on a machine without a GPU, exports every symbol include/hvb.h
declares, and refuses loudly to create a context when there is no device (no CPU fallback)."""
import ctypes
import os
import re
import subprocess

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
HEADER = os.path.join(ROOT, "include", "hvb.h")


def declared_symbols():
    src = open(HEADER).read()
    return sorted(set(re.findall(r"HVB_API\s+(?:const\s+char\*|int)\s+(hvb_\w+)\s*\(", src)))


def test_library_built_and_loads():
    from hvb import _ffi
    assert os.path.exists(_ffi.LIB_PATH), "run __graft_entry__.build() first"
    lib = _ffi.lib()
    assert lib.hvb_version() == 100


def test_every_declared_symbol_is_exported_and_bound():
    from hvb import _ffi
    lib = ctypes.CDLL(_ffi.LIB_PATH)
    names = declared_symbols()
    assert len(names) >= 50
    for n in names:
        assert hasattr(lib, n), "symbol %s declared in hvb.h but not exported" % n
        assert n in _ffi.PROTOTYPES, "symbol %s has no ctypes prototype" % n
    assert set(_ffi.PROTOTYPES) == set(names)


def test_no_extra_exports():
    from hvb import _ffi
    out = subprocess.run(["nm", "-D", "--defined-only", _ffi.LIB_PATH], capture_output=True, text=True).stdout
    exported = {l.split()[-1] for l in out.splitlines() if " T " in l}
    assert {e for e in exported if e.startswith("hvb_")} == set(declared_symbols())


def test_struct_layouts_match_header():
    from hvb import _ffi
    assert _ffi.IMG_META.itemsize == 32 and _ffi.IMG_META.fields["out_slot"][1] == 28
    assert _ffi.CROP_DESC.fields["pitch"][1] == 8 and _ffi.CROP_DESC.fields["w"][1] == 16
    assert _ffi.COLOR_RAW.fields["sums"][1] == 176 and _ffi.COLOR_RAW.fields["sumsq"][1] == 224
    assert _ffi.LB_TILE.fields["gain"][1] == 56


def test_device_count_and_loud_failure_without_gpu():
    import torch
    from hvb import _ffi
    lib = _ffi.lib()
    n = ctypes.c_int(-1)
    assert lib.hvb_device_count(ctypes.byref(n)) == 0
    if torch.cuda.is_available():
        pytest.skip("a GPU is present; the no-device path is covered on the CPU box")
    assert n.value == 0
    h = ctypes.c_void_p()
    st = lib.hvb_ctx_create(0, ctypes.byref(h))
    assert st == _ffi.HVB_ERR_NO_DEVICE
    assert b"no CPU fallback" in lib.hvb_last_error()
    from hvb.runtime import Context
    with pytest.raises(_ffi.HvbError):
        Context(0)
    with pytest.raises(_ffi.HvbError):
        from hvb.runtime import get_context
        get_context("cpu")


def test_product_never_imports_oracle():
    pkg = os.path.join(ROOT, "hockey-vision-analytics_b200")
    for dirpath, _, files in os.walk(pkg):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".h")):
                txt = open(os.path.join(dirpath, f)).read()
                assert not re.search(r"^\s*(from|import)\s+oracle\b", txt, re.M), "%s imports the oracle" % f


def test_stage_frames_copies_every_frame_with_any_thread_count():
    """hvb_stage_frames (host-side staging of separately allocated frames into one buffer) needs no GPU."""
    from hvb import _ffi
    rng = np.random.default_rng(0)
    frames = [rng.integers(0, 256, (37, 53, 3), dtype=np.uint8) for _ in range(11)]
    ptrs = (ctypes.c_void_p * len(frames))(*[f.ctypes.data for f in frames])
    for threads in (1, 3, 8, 100):
        dst = np.zeros((len(frames), 37, 53, 3), np.uint8)
        _ffi.check(_ffi.lib().hvb_stage_frames(ctypes.cast(ptrs, ctypes.c_void_p), len(frames), frames[0].nbytes, dst.ctypes.data, threads))
        assert all(np.array_equal(dst[k], f) for k, f in enumerate(frames))
    assert _ffi.lib().hvb_stage_frames(None, 0, 10, None, 4) == 0                   # empty chunk
    assert _ffi.lib().hvb_stage_frames(None, 2, 10, dst.ctypes.data, 4) != 0        # loud on a null table
