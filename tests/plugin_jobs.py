"""Shared helper of the plugin tests: run one of the reference's own regression jobs (testfiles/<job>/test.inp) through
the reference's driver with `solver=NIST` or `solver=NISTB200` and a chosen backend factory.  Needs the reference tree
(baseline/_ref — travels to the GPU box — or /root/reference); importing this module does not."""
import os
import re
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)
from tools import refshim  # noqa: E402


def available():
    return refshim.reference_available()


def job_text(testdir, inp="test.inp"):
    return open(os.path.join(refshim.REFERENCE_ROOT, "testfiles", testdir, inp)).read()


def uref(testdir):
    return np.loadtxt(os.path.join(refshim.REFERENCE_ROOT, "testfiles", testdir, "U.ref"))


def config0_text():
    """BASELINE config 1, testfiles/LinearElasticIsotropic/test.inp, with the element / material provider switched from the
    un-vendored Marmot library to the reference's own Python classes (same C3D8 formulation, same isotropic Hooke law)."""
    t = job_text("LinearElasticIsotropic")
    t = t.replace("*element, type=C3D8, provider=marmot", "*element, type=C3D8, provider=edelweiss")
    t = t.replace("*material, name=LinearElastic, id=myMaterial", "*material, name=linearelastic, id=myMaterial, provider=edelweiss")
    return t


def run(text, tag, solver="NISTB200", backend=None, solver_options=""):
    """Returns (U, model, fieldOutputController, created backends)."""
    refshim.bootstrap()
    refshim.build_native_helpers()
    from edelweissfe.drivers.inputfiledrivensimulation import finiteElementSimulation
    from edelweissfe.utils.inputfileparser import parseInputFile

    from edelweissfe_b200 import nistb200

    created = []
    if backend is None:
        backend = nistb200.default_backend

    def factory(*a, **k):
        b = backend(*a, **k)
        created.append(b)
        return b

    nistb200.register(factory)
    if solver != "NIST":
        text = re.sub(r"solver=NIST,", f"solver={solver},", text)
        if solver_options:  # options are data lines of the *solver keyword (helpers/inputfilehelpers.py:279-285)
            text = re.sub(r"(\*solver,[^\n]*\n)", r"\1" + solver_options + "\n", text, count=1)
    tmp = f"/tmp/ewb_{tag}_{solver}.inp"
    open(tmp, "w").write(text)
    cwd = os.getcwd()
    os.chdir("/tmp")  # outputs go to a scratch dir
    try:
        inputFile = parseInputFile(tmp)
        inputFile["*output"] = [o for o in inputFile["*output"] if o.get("type") != "ensight"]
        model, foc = finiteElementSimulation(inputFile, verbose=False, suppressPlots=True)
    finally:
        os.chdir(cwd)
    U = np.hstack([f["U"].flatten() for f in model.nodeFields.values()] + [v.value for v in model.scalarVariables.values()])
    return U, model, foc, created


def field_outputs(foc):
    return {name: np.array(fo.getLastResult(), dtype=float) for name, fo in foc.fieldOutputs.items()}
