"""The index views the two-slot kernel walks (symbolic.h: build_owner_ell / build_assembly_ell -- ELL-by-owner views of A and
P, assembly lists of S = P + sigma I + A' diag(rho) A per factorisation round) against dense / CSC algebra on random values.
CPU only (tests/cpp/ell_views.cpp): a wrong table entry would otherwise only show as a parity failure on the device."""
import os
import subprocess

import pytest

from common import ROOT, package


def _exe():
    package()  # builds libgomp_host.so if needed
    out = os.path.join(ROOT, "tests", "cpp", "build")
    os.makedirs(out, exist_ok=True)
    exe = os.path.join(out, "ell_views")
    src = os.path.join(ROOT, "tests", "cpp", "ell_views.cpp")
    dep = os.path.join(ROOT, "osqp-solver_b200", "csrc", "symbolic.h")
    if not os.path.exists(exe) or any(os.path.getmtime(s) > os.path.getmtime(exe) for s in (src, dep)):
        pkg = os.path.join(ROOT, "osqp-solver_b200")
        subprocess.check_call(["g++", "-std=c++17", "-O1", "-Wall", "-I", os.path.join(ROOT, "include"), "-I", os.path.join(pkg, "csrc"), src,
                               "-L", pkg, "-lgomp_host", f"-Wl,-rpath,{pkg}", "-o", exe])
    return exe


@pytest.mark.parametrize("args", ["7 50 1 1 192", "6 30 1 1 192", "7 36 2 2 192", "7 8 1 1 192", "7 4 1 1 192", "3 20 1 0 192", "7 50 1 1 224"])
def test_owner_and_assembly_views(args):
    r = subprocess.run([_exe()] + args.split(), capture_output=True, text=True)
    assert r.returncode == 0 and r.stdout.startswith("ok "), r.stdout + r.stderr
