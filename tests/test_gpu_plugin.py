"""The drop-in, end to end, on the GPU: the reference's own regression jobs run through the reference's driver with
`*solver, solver=NISTB200` and the CUDA backend (libedelweiss_b200.so through the C ABI), against the shipped U.ref golden
vectors and the reference's own serial NIST solver.  The reference tree travels to the GPU box as baseline/_ref
(tools/install_reference.py); without it these tests skip."""
import re

import numpy as np
import plugin_jobs as jobs
import pytest

pytestmark = [pytest.mark.gpu, pytest.mark.skipif(not jobs.available(), reason="reference tree not present (run tools/install_reference.py)")]

JOBS = ["WallShearHexa8", "WallShearHexa20", "CantileverBeamHexa8", "TensionBarHexa8", "SimpleBeamHexa8"]


@pytest.mark.parametrize("testdir", JOBS)
def test_reference_jobs_cuda_backend(testdir):
    U, model, foc, created = jobs.run(jobs.job_text(testdir), testdir)
    assert created and type(created[0]).__name__ == "ElementAssembly" and created[0].launch_count() > 0
    assert np.abs(U - jobs.uref(testdir)).max() < 1e-6  # the reference's acceptance test (_cli/_run_tests_edelweissfe.py:102-105)
    U0, _, foc0, _ = jobs.run(jobs.job_text(testdir), testdir, solver="NIST")
    assert np.abs(U - U0).max() < 1e-8
    f, f0 = jobs.field_outputs(foc), jobs.field_outputs(foc0)
    assert set(f) == set(f0)
    for name in f0:
        assert np.isfinite(f[name]).all(), name
        assert np.abs(f[name] - f0[name]).max() <= 1e-5 * np.abs(f0[name]).max() + 1e-12, name


def test_config0_cuda_backend():
    """BASELINE configs[0] (testfiles/LinearElasticIsotropic/test.inp, provider switched to the reference's Python classes):
    device element loop + device face-pressure load against the shipped U.ref."""
    U, model, foc, created = jobs.run(jobs.config0_text(), "config0")
    Uref = jobs.uref("LinearElasticIsotropic")
    assert created and U.shape == Uref.shape == (24,)
    assert np.abs(U - Uref).max() < 1e-9


def test_von_mises_job_cuda_backend():
    src = jobs.job_text("WallShearHexa8VonMises", "testLong.inp")
    src = re.sub(r"nX\s*=\s*\d+", "nX=4", src)
    src = re.sub(r"nY\s*=\s*\d+", "nY=4", src)
    res = {s: jobs.run(src, "vm_small", solver=s)[0] for s in ("NIST", "NISTB200")}
    assert np.abs(res["NIST"] - res["NISTB200"]).max() < 1e-5 * np.abs(res["NIST"]).max()


@pytest.mark.parametrize("testdir", ["WallShearHexa8", "TensionBarHexa8"])
def test_device_resident_matrix_and_pcg(testdir):
    """b200solver=pcg: assembleStiffnessCSR returns a device handle, applyDirichletK and linearSolve run on the device; only
    dof-sized vectors cross PCIe.  Same U.ref acceptance test."""
    U, model, foc, created = jobs.run(jobs.job_text(testdir), testdir + "_pcg", solver_options="b200solver=pcg\nb200pcgtol=1e-13")
    assert created and not hasattr(created[0], "_pin_K")  # the CSR values were never copied to the host
    assert np.abs(U - jobs.uref(testdir)).max() < 1e-6
