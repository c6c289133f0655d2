"""The drop-in boundary from a plain C99 caller: include/hp_b200.h must compile as C, and a C program linked against
libhp_b200.so gets version / workspace size / negative argument-error codes without any torch or C++ in the way
(examples/c_abi_demo.c).  The CPU run makes no CUDA call; `--gpu` (opt-in: HP_RUN_C_DEMO_GPU=1 on a GPU box) also decodes
one map through the library."""
import importlib
import os
import shutil
import subprocess

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
PKG = "domain-adaptative-hand-pose-estimation_b200"


def _cuda_lib_dir():
    for d in ("/usr/local/cuda/lib64", "/usr/local/cuda/targets/x86_64-linux/lib"):
        if os.path.exists(os.path.join(d, "libcudart.so")):
            return d
    return None


@pytest.fixture(scope="module")
def demo(tmp_path_factory):
    if shutil.which("gcc") is None or _cuda_lib_dir() is None:
        pytest.skip("gcc or libcudart not available")
    lib = importlib.import_module(PKG + ".build").build()
    pkg_dir = os.path.dirname(lib)
    exe = str(tmp_path_factory.mktemp("cdemo") / "c_abi_demo")
    cmd = ["gcc", "-std=c99", "-Wall", "-Wextra", "-Werror", "-I", os.path.join(ROOT, "include"),
           os.path.join(ROOT, "examples", "c_abi_demo.c"), "-L", pkg_dir, "-lhp_b200", "-L", _cuda_lib_dir(), "-lcudart",
           f"-Wl,-rpath,{pkg_dir}", f"-Wl,-rpath,{_cuda_lib_dir()}", "-o", exe]
    p = subprocess.run(cmd, capture_output=True, text=True)
    assert p.returncode == 0, p.stderr
    return exe


def test_header_is_valid_c99_and_cxx17():
    if shutil.which("gcc") is None:
        pytest.skip("gcc not available")
    hdr = os.path.join(ROOT, "include", "hp_b200.h")
    for args in (["gcc", "-std=c99", "-pedantic", "-x", "c"], ["g++", "-std=c++17", "-x", "c++"]):
        p = subprocess.run(args + ["-Wall", "-Wextra", "-Werror", "-fsyntax-only", hdr], capture_output=True, text=True)
        assert p.returncode == 0, p.stderr


def test_plain_c_caller_links_and_gets_error_codes(demo):
    p = subprocess.run([demo], capture_output=True, text=True, timeout=120)
    assert p.returncode == 0, p.stdout + p.stderr
    assert "hp_version = " in p.stdout and p.stdout.strip().endswith("ok")


@pytest.mark.gpu
@pytest.mark.skipif(os.environ.get("HP_RUN_C_DEMO_GPU") != "1", reason="opt-in: HP_RUN_C_DEMO_GPU=1")
def test_plain_c_caller_decodes_on_the_gpu(demo):
    p = subprocess.run([demo, "--gpu"], capture_output=True, text=True, timeout=120)
    assert p.returncode == 0, p.stdout + p.stderr
    assert "decoded (42, 17) max 0.75" in p.stdout
