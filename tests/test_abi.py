"""CPU: the C-ABI library loads and exports every symbol include/bds.h declares."""
import os
import re


def test_library_loads_and_exports_all_symbols():
    from bilateral_driving_b200 import _lib

    assert _lib.lib.bds_abi_version() == _lib.ABI_VERSION == 5
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    header = open(os.path.join(root, "include", "bds.h")).read()
    declared = set(re.findall(r"\b(bds_[a-z0-9_]+)\s*\(", header))
    assert declared == set(_lib.ABI_SYMBOLS), declared ^ set(_lib.ABI_SYMBOLS)
    for name in declared:
        assert hasattr(_lib.lib, name), f"{name} not exported by libbds_b200.so"


def test_ops_fail_loudly_without_cuda():
    import pytest
    import torch

    from bilateral_driving_b200._lib import BdsError
    from bilateral_driving_b200.bilateral import multiscale_bilateral

    if torch.cuda.is_available():
        pytest.skip("CPU-only check")
    with pytest.raises(BdsError):
        multiscale_bilateral(torch.rand(8, 8, 3), [torch.zeros(12, 1, 2, 2)], [(2, 2, 1)], None)


def test_header_is_plain_c_and_links(tmp_path):
    """include/bds.h is a C header (extern "C", plain pointers and sizes): a C99 translation unit that includes it
    compiles with -pedantic, links against libbds_b200.so and sees the same ABI version."""
    import shutil
    import subprocess

    if shutil.which("gcc") is None:
        import pytest
        pytest.skip("no gcc")
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    libdir = os.path.join(root, "bilateral_driving_b200")
    src = tmp_path / "c_abi.c"
    src.write_text('#include "include/bds.h"\n#include <stddef.h>\n'
                   'int main(void) { return (bds_abi_version() == BDS_ABI_VERSION && bds_last_error() != NULL) ? 0 : 1; }\n')
    exe = tmp_path / "c_abi"
    subprocess.run(["gcc", "-std=c99", "-Wall", "-Wextra", "-pedantic", "-Werror", "-I", root, str(src), "-L", libdir,
                    "-lbds_b200", f"-Wl,-rpath,{libdir}", "-o", str(exe)], check=True, capture_output=True)
    assert subprocess.run([str(exe)]).returncode == 0
