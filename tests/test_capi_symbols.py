"""The C-ABI library loads without a GPU and exports every symbol include/mpeghb200.h declares."""
import ctypes
import os
import re

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def declared_symbols():
    text = open(os.path.join(ROOT, "include", "mpeghb200.h")).read()
    text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
    return sorted(set(re.findall(r"\b(mpeghb200_[a-z0-9_]+)\s*\(", text)))


def test_header_declares_entry_points():
    syms = declared_symbols()
    for must in ("mpeghb200_create", "mpeghb200_destroy", "mpeghb200_fd_synth_dev", "mpeghb200_fd_execute"):
        assert must in syms


def test_library_exports_every_declared_symbol(so_lib):
    for s in declared_symbols():
        assert hasattr(so_lib.c, s), f"{s} declared in include/mpeghb200.h but not exported"


def test_no_torch_or_python_types_in_header():
    text = open(os.path.join(ROOT, "include", "mpeghb200.h")).read()
    assert "torch" not in text and "PyObject" not in text and "std::" not in text


def test_product_does_not_reference_oracle():
    """The product sources never include, link or load anything under oracle/."""
    for base in ("libmpegh_b200", "include"):
        for d, _, files in os.walk(os.path.join(ROOT, base)):
            if "build" in d.split(os.sep):
                continue
            for f in files:
                if f.endswith((".cu", ".cpp", ".h", ".py", "Makefile", ".inc")):
                    src = open(os.path.join(d, f), errors="ignore").read()
                    for line in src.splitlines():
                        code = line.split("//")[0]
                        assert not re.search(r'(#include|import|CDLL|dlopen).*oracle', code), (f, line)


def test_no_cpu_fallback_without_gpu(so_lib):
    import torch
    if torch.cuda.is_available():
        pytest.skip("GPU present")
    from libmpegh_b200.lib import MpeghError, Lib
    with pytest.raises(MpeghError) as e:
        Lib().create(0)
    assert (e.value.code & 0xFFFFFFFF) == 0xFFFF8001


def test_null_handles_are_rejected_without_a_device(so_lib):
    """Argument checks come before any device work: a NULL renderer / context is MPEGHB200_ERR_BAD_ARG."""
    c = so_lib.c
    assert (c.mpeghb200_hoa_renderer_set_tensor_mode(None, 1) & 0xFFFFFFFF) == 0xFFFF8003
    assert (c.mpeghb200_hoa_render_dev(None, None, None, None, None, 1, None) & 0xFFFFFFFF) == 0xFFFF8003
    assert c.mpeghb200_hoa_nfc_state_floats(None) == 0
