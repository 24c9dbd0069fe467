"""The C-ABI shared library loads without a GPU and exports every symbol include/mskc.h declares;
the product path fails loudly (no CPU fallback) when no CUDA device is present."""
import ctypes as C
import os
import re

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.fixture(scope="module")
def lib():
    from mulskips_b200 import capi
    if not os.path.exists(capi.lib_path()):
        import __graft_entry__ as g
        g.build()
    return capi.load_library()


def header_symbols():
    txt = open(os.path.join(ROOT, "include", "mskc.h")).read()
    txt = re.sub(r"/\*.*?\*/", "", txt, flags=re.S)
    return sorted(set(re.findall(r"\b(mskc_[a-z_0-9]+)\s*\(", txt)))


def test_header_matches_binding(lib):
    from mulskips_b200 import capi
    assert header_symbols() == sorted(capi.SYMBOLS)


def test_fortran_binding_names():
    """fortran/mskc_bind.f90 (the ISO_C_BINDING module a Fortran host uses) binds exactly the header's symbols."""
    txt = open(os.path.join(ROOT, "fortran", "mskc_bind.f90")).read()
    bound = sorted(set(re.findall(r'bind\(C,\s*name="(mskc_[a-z_0-9]+)"\)', txt)))
    assert bound == header_symbols()


def test_exports(lib):
    for s in header_symbols():
        assert hasattr(lib, s), s


def test_struct_sizes(tmp_path):
    """ctypes mirrors have the layout the C compiler gives include/mskc.h (compiled here with gcc)."""
    import subprocess
    from mulskips_b200 import capi
    src = tmp_path / "sz.c"
    src.write_text('#include <stdio.h>\n#include "mskc.h"\nint main(){printf("%zu %zu %zu %zu %zu %zu\\n", sizeof(mskc_parity_args), sizeof(mskc_parity_result), '
                   'sizeof(mskc_parallel_args), sizeof(mskc_parallel_result), sizeof(mskc_counters), sizeof(mskc_census));return 0;}\n')
    exe = tmp_path / "sz"
    subprocess.run(["gcc", "-I", os.path.join(ROOT, "include"), str(src), "-o", str(exe)], check=True)
    sizes = [int(v) for v in subprocess.run([str(exe)], capture_output=True, text=True, check=True).stdout.split()]
    mirrors = [capi._ParityArgs, capi.ParityResult, capi._ParallelArgs, capi.ParallelResult, capi.Counters, capi.Census]
    assert sizes == [C.sizeof(m) for m in mirrors]


def test_no_cpu_fallback(lib):
    import torch
    if torch.cuda.is_available() or os.environ.get("MSKC_EMU") == "1":
        pytest.skip("GPU present (or the SIMT-emulated library was selected explicitly)")
    from mulskips_b200 import Mskc, MskcError
    with pytest.raises(MskcError) as e:
        Mskc(96, 96, 96, [14])
    assert "no CUDA device" in str(e.value) or "CUDA" in str(e.value)


def test_product_does_not_touch_oracle():
    """Nothing under mulskips_b200/ or include/ may reference oracle/ (the oracle is test infrastructure)."""
    for base in ("mulskips_b200", "include"):
        for dp, _, fs in os.walk(os.path.join(ROOT, base)):
            for f in fs:
                if f.endswith((".py", ".cu", ".cuh", ".h", ".cpp", ".hpp", "Makefile")):
                    t = open(os.path.join(dp, f), errors="ignore").read()
                    assert "oracle_binding" not in t and "mskc_oracle" not in t and "libmskc_oracle" not in t, os.path.join(dp, f)


def test_null_handle_is_an_error_not_a_crash(lib):
    """Every entry point that takes a handle returns a negative status for a NULL handle and leaves a message in
    mskc_last_error(NULL) -- errors come back as int + text, never as a crash or an exception across the boundary
    (include/mskc.h conventions).  No compute call is made: nothing here needs a GPU."""
    null = C.c_void_p(None)
    calls = {
        "mskc_set_rates": (null, None, None, None, None, C.c_double(1.0)),
        "mskc_set_la": (null, 1, C.c_double(1688.0), C.c_double(1e17), C.c_double(280.0), C.c_double(1080.0), 1600, 1700),
        "mskc_set_rate_field": (null, 1, None, None, None, None, None, C.c_double(1.0), 900, 1000, 0),
        "mskc_set_tree_variant": (null, 1),
        "mskc_set_la_fork": (null, 1, None, None, None, None, 1600, 1700, None, 0, 0),
        "mskc_upload_state": (null, None, 0, None, None, None, None, 0, None, None, 0, None, None),
        "mskc_run_parity": (null, None, None),
        "mskc_run_parallel": (null, None, None),
        "mskc_download_lattice": (null, None),
        "mskc_get_counters": (null, None),
        "mskc_set_temperature": (null, None),
        "mskc_extract_frame": (null, 0, None, None, None, None),
        "mskc_extract_frame_all": (null, 0, None, None, None, None),
        "mskc_extract_antisites": (null, 0, 0, None, None, None, None),
        "mskc_gather_lattice": (null, 0, None, None),
        "mskc_replay_events": (null, 0, None, None, None, C.c_double(1.0)),
    }
    for name, args in calls.items():
        rc = getattr(lib, name)(*args)
        assert rc < 0, name
        msg = lib.mskc_last_error(None).decode()
        assert msg, name
