"""The C-ABI library loads on a CPU-only box and exports every symbol include/swr_b200.h declares
(no compute calls here: kernels need a GPU).  Also pins the record layout shared by the C executor,
the host lowering and the CPU interpreter."""
import ctypes
import os
import re

import numpy as np

from scenario_wise_rec_b200 import _native as N

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
HEADER = os.path.join(ROOT, "include", "swr_b200.h")


def declared_symbols():
    src = open(HEADER).read()
    return sorted(set(re.findall(r"SWR_API\s+[\w\s\*]+?\b(swr_\w+)\s*\(", src)))


def test_library_exports_every_declared_symbol():
    names = declared_symbols()
    assert len(names) >= 9
    lib = ctypes.CDLL(N.LIB_PATH)
    for n in names:
        assert hasattr(lib, n), f"{n} is declared in include/swr_b200.h but not exported by libswr_b200.so"
    assert set(names) == set(N.EXPORTS)


def test_abi_version_and_error_string():
    L = N.lib()
    assert L.swr_abi_version() == N.ABI_VERSION
    assert N.last_error() == ""
    assert N.launch_count() >= 0


def test_record_layout_matches_header():
    src = open(HEADER).read()
    ints = int(re.search(r"#define SWR_REC_INTS (\d+)", src).group(1))
    floats = int(re.search(r"#define SWR_REC_FLOATS (\d+)", src).group(1))
    slots = int(re.search(r"#define SWR_REC_SLOTS (\d+)", src).group(1))
    assert (ints, floats, slots) == (N.REC_INTS, N.REC_FLOATS, N.REC_SLOTS)
    assert N.REC_DTYPE.itemsize == 8 + 4 * (ints + floats + slots)
    # op kinds: header enum == python constants == interpreter constants
    from oracle import ops_ref
    for name, val in re.findall(r"SWR_(OP_\w+) = (\d+)", src):
        assert getattr(N, name) == int(val), name
        if hasattr(ops_ref, name):
            assert getattr(ops_ref, name) == int(val), name


def test_malformed_program_is_rejected_without_a_gpu():
    recs = np.zeros(1, dtype=N.REC_DTYPE)
    recs[0]["kind"] = 999
    slots = np.zeros(1, dtype=np.uint64)
    st = N.lib().swr_program_run(recs.ctypes.data, 1, slots.ctypes.data, 1, None)
    assert st == -1 and "unknown op kind" in N.last_error()


def test_fc_mode_switch_roundtrip():
    """swr_set_fc_mode / swr_get_fc_mode need no device: the mode is process state read at launch time."""
    prev = N.get_fc_mode()
    assert prev in (N.FC_SIMT, N.FC_TC, N.FC_AUTO)
    try:
        assert N.set_fc_mode(N.FC_SIMT) == prev and N.get_fc_mode() == N.FC_SIMT
        assert N.set_fc_mode(7) == N.FC_SIMT and N.get_fc_mode() == N.FC_SIMT      # out of range: ignored
        assert N.set_fc_mode(N.FC_TC) == N.FC_SIMT and N.get_fc_mode() == N.FC_TC
    finally:
        N.set_fc_mode(prev)


def test_header_is_plain_c_and_links_from_c(tmp_path):
    """The drop-in boundary is a C ABI: include/swr_b200.h compiles as C99 with nothing but <stdint.h>, and a C program
    links against libswr_b200.so and calls it (no device needed for swr_abi_version)."""
    import shutil
    import subprocess
    if shutil.which("gcc") is None:
        import pytest
        pytest.skip("gcc not available")
    src = tmp_path / "abi.c"
    src.write_text('#include "swr_b200.h"\nint main(void) { return swr_abi_version() == SWR_ABI_VERSION ? 0 : 1; }\n')
    libdir = os.path.dirname(N.LIB_PATH) if hasattr(N, "LIB_PATH") else os.path.join(ROOT, "scenario-wise-rec_b200", "scenario_wise_rec_b200")
    exe = tmp_path / "abi"
    subprocess.run(["gcc", "-std=c99", "-Wall", "-Werror", "-I", os.path.join(ROOT, "include"), str(src), "-o", str(exe),
                    "-L", libdir, "-l:libswr_b200.so", f"-Wl,-rpath,{libdir}"], check=True, capture_output=True)
    assert subprocess.run([str(exe)]).returncode == 0
