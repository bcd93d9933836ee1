"""A compiled consumer of the drop-in boundary: tests/cdriver/abi_driver.c includes include/*.h as plain C11 and links
libhalo_b200.so / libhalo_host.so the way the reference crate's build.rs would -- no ctypes, no Python
marshalling between the caller and the ABI.  It drives MSMs, PCDL commit / open / check and one ASDL accumulation step at
n = 2^10 and compares every result with the oracle inside the same process."""
import os
import subprocess

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
LIB = os.path.join(ROOT, "halo-accumulation_b200", "lib")
ORC = os.path.join(ROOT, "oracle")


def _build(curve):
    import halo_accumulation_b200 as H

    H.build()
    subprocess.check_call(["make", "-C", ORC, "all"], stdout=subprocess.DEVNULL)
    sfx = "" if curve == "pallas" else "_vesta"
    exe = os.path.join(ROOT, "tests", "cdriver", f"abi_driver{sfx}.bin")
    subprocess.check_call(["/usr/bin/gcc", "-std=c11", "-Wall", "-Wextra", "-Werror", "-O1", "-o", exe,
                           os.path.join(ROOT, "tests", "cdriver", "abi_driver.c"),
                           f"-L{LIB}", f"-lhalo_host{sfx}", f"-lhalo_b200{sfx}", f"-L{ORC}", f"-l:liboracle{sfx}.so",
                           f"-Wl,-rpath,{LIB}", f"-Wl,-rpath,{ORC}"])
    return exe


def _have_cuda():
    import torch

    return torch.cuda.is_available()


@pytest.mark.parametrize("curve", ["pallas", "vesta"])
def test_headers_compile_as_c_and_fail_loudly_without_a_device(curve):
    """The headers are valid C11 (-Wall -Wextra -Werror), the libraries link from C, and without a GPU the program reports
    that there is no CPU fallback (exit status 2) instead of computing anything."""
    exe = _build(curve)
    if _have_cuda():
        pytest.skip("a GPU is present: the run is covered by the gpu-marked test")
    r = subprocess.run([exe], capture_output=True, text=True, timeout=120)
    assert r.returncode == 2, (r.returncode, r.stdout, r.stderr)
    assert "no CPU fallback" in r.stderr and f"({curve})" in r.stderr


@pytest.mark.gpu
@pytest.mark.parametrize("curve", ["pallas", "vesta"])
def test_c_program_matches_the_oracle_on_the_gpu(curve):
    exe = _build(curve)
    r = subprocess.run([exe], capture_output=True, text=True, timeout=600)
    assert r.returncode == 0, (r.returncode, r.stdout[-2000:], r.stderr[-2000:])
    assert f"abi_driver ok ({curve})" in r.stdout
