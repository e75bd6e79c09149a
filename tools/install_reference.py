#!/usr/bin/env python
"""Install the UNMODIFIED reference next to the repo so that it travels to the GPU box with `gpurun`
(`baseline/_ref/` is git-ignored but not gpurun-ignored; nothing of it enters the history).

    python tools/install_reference.py            # idempotent; run by __graft_entry__.build() when /root/reference exists

`pip install --no-index --no-build-isolation --target baseline/_ref /root/reference` does not work in this image: the
reference's setup.py (setup.py:64-197) unconditionally builds Cython extensions against Marmot and MKL, which are not
installed.  So the install is done by hand, exactly as SURVEY.md Appendix B describes:

  1. copy the `edelweissfe/` package and the regression jobs the tests use (`testfiles/<job>/`) as they are,
  2. compile the three dependency-free Cython modules in place with the reference's own directives (setup.py:41-46):
     numerics/csrgenerator.pyx, utils/elementresultcollector.pyx, solvers/nonlinearimplicitstaticparallelmk2.pyx.

Missing optional dependencies (h5py, prettytable, matplotlib, Marmot wrappers, MKL pardiso) are stubbed at import time by
tools/refshim.py — no reference file is edited.
"""
import os
import shutil
import subprocess
import sys
import sysconfig

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
SRC = os.environ.get("EDELWEISS_REFERENCE_SRC", "/root/reference")
DST = os.path.join(ROOT, "baseline", "_ref")

JOBS = ["WallShearHexa8", "WallShearHexa20", "CantileverBeamHexa8", "TensionBarHexa8", "SimpleBeamHexa8", "WallShearHexa8VonMises",
        "LinearElasticIsotropic", "CantileverBeamQuad8NeoHookeWa"]
PYX = [("edelweissfe/numerics/csrgenerator.pyx", False), ("edelweissfe/utils/elementresultcollector.pyx", False),
       ("edelweissfe/solvers/nonlinearimplicitstaticparallelmk2.pyx", True)]


def installed() -> bool:
    suffix = sysconfig.get_config_var("EXT_SUFFIX")
    return all(os.path.exists(os.path.join(DST, os.path.splitext(p)[0] + suffix)) for p, _ in PYX)


def install(force=False) -> bool:
    """Returns True when baseline/_ref is usable afterwards."""
    if installed() and not force:
        return True
    if not os.path.isdir(os.path.join(SRC, "edelweissfe")):
        return False
    import numpy

    os.makedirs(DST, exist_ok=True)
    shutil.copytree(os.path.join(SRC, "edelweissfe"), os.path.join(DST, "edelweissfe"), dirs_exist_ok=True,
                    ignore=shutil.ignore_patterns("__pycache__", "*.pyc"))
    for job in JOBS:
        src = os.path.join(SRC, "testfiles", job)
        if os.path.isdir(src):
            shutil.copytree(src, os.path.join(DST, "testfiles", job), dirs_exist_ok=True)
    suffix = sysconfig.get_config_var("EXT_SUFFIX")
    cxx = "/usr/bin/g++" if os.path.exists("/usr/bin/g++") else "g++"
    for rel, openmp in PYX:
        pyx = os.path.join(DST, rel)
        cpp = os.path.splitext(pyx)[0] + ".cpp"
        so = os.path.splitext(pyx)[0] + suffix
        subprocess.check_call([sys.executable, "-m", "cython", "-3", "--cplus", "-X", "boundscheck=False", "-X", "wraparound=False",
                               "-X", "nonecheck=False", "-X", "initializedcheck=False", "-I", DST, pyx, "-o", cpp])
        cmd = [cxx, "-O3", "-fPIC", "-shared", "-I" + sysconfig.get_paths()["include"], "-I" + numpy.get_include(), cpp, "-o", so]
        if openmp:
            cmd.insert(1, "-fopenmp")
            cmd.insert(2, "-Wno-maybe-uninitialized")
        subprocess.check_call(cmd)
        os.remove(cpp)
    return installed()


if __name__ == "__main__":
    ok = install(force="--force" in sys.argv)
    print("baseline/_ref:", "installed" if ok else "unavailable (no reference tree at %s)" % SRC)
    sys.exit(0 if ok else 1)
