"""CPU: the C-ABI library builds, loads and exports every symbol include/regda_b200.h declares.
No compute call is made (there is no GPU here)."""
import ctypes
import os
import subprocess

from regda_b200 import build, capi


def test_library_builds_and_exports_every_declared_symbol():
    path = build.build_library()
    assert os.path.exists(path)
    L = ctypes.CDLL(path)
    names = [n for n, _, _ in capi.declared_symbols()]
    assert len(names) >= 7 and len(set(names)) == len(names)
    for n in names:
        assert hasattr(L, n), f"{n} declared in include/regda_b200.h but not exported"
    L.regda_abi_version.restype = ctypes.c_int
    assert L.regda_abi_version() == 3


def test_sass_is_sm100a_only():
    out = subprocess.run(["cuobjdump", "--list-elf", build.LIB], capture_output=True, text=True).stdout
    archs = {ln.split(".")[-2] for ln in out.splitlines() if ".cubin" in ln}
    assert archs == {"sm_100a"}, archs


def test_product_never_imports_the_oracle():
    root = os.path.dirname(os.path.abspath(capi.__file__))
    for dp, _, files in os.walk(root):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".h", ".cpp")):
                src = open(os.path.join(dp, f), errors="ignore").read()
                assert "import oracle" not in src and "from oracle" not in src and "oracle/" not in src, os.path.join(dp, f)
