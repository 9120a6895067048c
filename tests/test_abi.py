"""CPU checks of the drop-in boundary: the C-ABI library loads, exports every symbol include/geotrax_b200.h declares, the
ctypes mirror of gt_config matches the C struct, and the product path refuses to run without a GPU (no fallback)."""
import ctypes as C
import os
import re
import subprocess
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
HEADER = os.path.join(ROOT, "include", "geotrax_b200.h")


@pytest.fixture(scope="module")
def lib():
    sys.path.insert(0, ROOT)
    import __graft_entry__ as ge
    import importlib.util
    spec = importlib.util.spec_from_file_location("_gt_build", os.path.join(ROOT, "geo-trax_b200", "build.py"))
    mod = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mod)
    mod.build()                                # no-op when the .so is newer than its sources
    import geotrax_b200
    return geotrax_b200.load_library()


def _declared_symbols():
    src = open(HEADER).read()
    src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
    return sorted(set(re.findall(r"\b(gt_[a-z0-9_]+)\s*\(", src)))


def test_header_symbols_all_exported(lib):
    from geotrax_b200 import _lib
    declared = _declared_symbols()
    assert len(declared) >= 30
    for name in declared:
        assert hasattr(lib, name), f"{name} is declared in include/geotrax_b200.h but not exported by the .so"
    assert set(declared) == set(_lib.SYMBOLS), f"ctypes binding drifted from the header: {set(declared) ^ set(_lib.SYMBOLS)}"


def test_no_torch_or_cxx_types_in_signatures():
    src = open(HEADER).read()
    assert 'extern "C"' in src
    for bad in ("torch", "at::", "std::", "Tensor", "&"):
        body = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
        assert bad not in body, f"{bad!r} leaks into the C-ABI header"


def test_config_struct_layout_matches_c(lib, tmp_path):
    """sizeof / field offsets of the ctypes gt_config vs the C compiler's."""
    from geotrax_b200._lib import gt_config, gt_conv_desc
    src = tmp_path / "sz.c"
    fields = ["abi_version", "frame_h", "max_batch", "imgsz", "max_nms", "downsample_ratio", "mask_margin_ratio", "ransac_max_iter", "seed",
              "act_dtype", "reserved"]
    prints = "".join(f'printf("{f} %zu\\n", offsetof(gt_config, {f}));' for f in fields)
    src.write_text(f'#include <stdio.h>\n#include <stddef.h>\n#include "{HEADER}"\nint main(){{printf("size %zu\\n", sizeof(gt_config));'
                   f'printf("desc %zu\\n", sizeof(gt_conv_desc));{prints}return 0;}}\n')
    exe = tmp_path / "sz"
    subprocess.check_call(["gcc", "-std=c99", str(src), "-o", str(exe)])      # also proves the header is plain C
    out = dict(l.split() for l in subprocess.check_output([str(exe)], text=True).splitlines())
    assert int(out["size"]) == C.sizeof(gt_config)
    assert int(out["desc"]) == C.sizeof(gt_conv_desc)
    for f in fields:
        assert int(out[f]) == getattr(gt_config, f).offset, f


def test_defaults_equal_reference_preset(lib):
    """gt_default_config == /root/reference/geotrax/cfg/default.yaml:103-137, 235-245."""
    from geotrax_b200._lib import gt_config
    c = gt_config()
    lib.gt_default_config(C.byref(c))
    assert (c.frame_h, c.frame_w, c.imgsz, c.nc, c.max_det, c.max_nms) == (2160, 3840, 1920, 4, 1000, 30000)
    assert (c.max_features, c.ransac_max_iter, c.mask_use, c.query_is_current) == (2000, 5000, 1, 1)
    assert abs(c.downsample_ratio - 0.5) < 1e-9 and abs(c.ref_multiplier - 2.0) < 1e-9 and abs(c.mask_margin_ratio - 0.15) < 1e-7
    assert abs(c.filter_ratio - 0.9) < 1e-7 and abs(c.ransac_threshold - 2.0) < 1e-9
    assert lib.gt_abi_version() == 1


def test_create_fails_loudly_without_gpu(lib):
    import torch
    if torch.cuda.is_available():
        pytest.skip("a GPU is present")
    from geotrax_b200._lib import gt_config
    c = gt_config()
    lib.gt_default_config(C.byref(c))
    h = C.c_void_p()
    rc = lib.gt_create(C.byref(c), 0, C.byref(h))
    assert rc < 0 and not h.value
    assert b"no CPU fallback" in lib.gt_last_error(None) or b"CUDA" in lib.gt_last_error(None)
    import geotrax_b200
    with pytest.raises(geotrax_b200.GtError):
        geotrax_b200.Engine()


def test_missing_library_raises(tmp_path):
    import geotrax_b200
    with pytest.raises(geotrax_b200.GtError):
        geotrax_b200.load_library(str(tmp_path / "nope.so"))


def test_product_never_imports_oracle():
    """The oracle is test infrastructure: nothing under geo-trax_b200/ or tools/ may import or execute it."""
    for sub in ("geo-trax_b200", "tools"):
        for dirpath, _, files in os.walk(os.path.join(ROOT, sub)):
            for f in files:
                if f.endswith((".py", ".cu", ".cuh", ".h")):
                    txt = open(os.path.join(dirpath, f)).read()
                    assert not re.search(r"^\s*(from|import)\s+oracle\b", txt, flags=re.M), f"{sub}/{f} imports the oracle"
