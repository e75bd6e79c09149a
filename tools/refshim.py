"""Bootstrap to import the UNMODIFIED EdelweissFE reference from /root/reference in this
container (it is Python; missing optional deps are stubbed in sys.modules, nothing is
copied or edited).  Used by tests/golden/make_golden.py (fixture generation), by the plugin tests (CPU: oracle-backed
stand-in; GPU box: the CUDA backend) and by bench.py's reference arm.  The tree is looked up at
baseline/_ref (tools/install_reference.py — git-ignored, travels with gpurun) and then at
/root/reference (build container only).  Never imported by the product path or smoke().

Recipe follows SURVEY.md Appendix B.
"""
import os
import sys
import types

_REPO = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
_INSTALLED = os.path.join(_REPO, "baseline", "_ref")  # tools/install_reference.py (travels to the GPU box, git-ignored)


def _default_root():
    if os.path.isdir(os.path.join(_INSTALLED, "edelweissfe")):
        return _INSTALLED
    return "/root/reference"


REFERENCE_ROOT = os.environ.get("EDELWEISS_REFERENCE_ROOT") or _default_root()


def reference_available() -> bool:
    return os.path.isdir(os.path.join(REFERENCE_ROOT, "edelweissfe"))


def _stub(name, **attrs):
    m = types.ModuleType(name)
    for k, v in attrs.items():
        setattr(m, k, v)
    sys.modules[name] = m
    return m


class _PrettyTable:
    def __init__(self, *a, **k):
        self.rows = []
        self.field_names = []
        self.align = {}
        self.min_width = {}
        self.border = self.header = True

    def add_row(self, row):
        self.rows.append(row)

    def __str__(self):
        return "\n".join(" ".join(str(c) for c in r) for r in self.rows)

    def get_string(self, *a, **k):
        return str(self)


class _Anything:
    def __init__(self, *a, **k):
        pass

    def __getattr__(self, name):
        return _Anything()

    def __call__(self, *a, **k):
        return _Anything()


def bootstrap():
    """Make `import edelweissfe` work against /root/reference (pure-Python parts)."""
    if "edelweissfe" in sys.modules and getattr(sys.modules["edelweissfe"], "_ewb_shimmed", False):
        return
    if not reference_available():
        raise ImportError("reference tree not present at %s" % REFERENCE_ROOT)
    if REFERENCE_ROOT not in sys.path:
        sys.path.insert(0, REFERENCE_ROOT)
    for name in ("h5py",):
        try:
            __import__(name)
        except ImportError:
            _stub(name, File=object)
    try:
        import prettytable  # noqa: F401
    except ImportError:
        _stub("prettytable", PrettyTable=_PrettyTable)
    try:
        import matplotlib  # noqa: F401
    except ImportError:
        pass
    _stub("edelweissfe.elements.marmotelement.element", MarmotElementWrapper=type("MarmotElementWrapper", (), {}))
    _stub(
        "edelweissfe.elements.marmotsingleqpelement.element",
        MarmotMaterialWrappingElement=type("MarmotMaterialWrappingElement", (), {}),
    )
    _stub("edelweissfe.utils.plotter", Plotter=_Anything)
    import scipy.sparse.linalg as spla

    _stub(
        "edelweissfe.linsolve.pardiso.pardiso",
        pardisoSolve=lambda A, b: spla.spsolve(A.tocsc(), b, use_umfpack=False),
    )
    import edelweissfe

    edelweissfe._ewb_shimmed = True


def build_cython_module(relpath: str, modname: str, openmp: bool = False):
    """Compile one dependency-free .pyx of the reference (where it lies, output only to a
    scratch dir under /tmp) and register it in sys.modules under its reference name.
    Flags follow the reference's setup.py:41-46,251 (boundscheck/wraparound off)."""
    import importlib.util
    import subprocess
    import sysconfig

    import numpy

    if modname in sys.modules:
        return sys.modules[modname]
    prebuilt = os.path.join(REFERENCE_ROOT, os.path.splitext(relpath)[0] + sysconfig.get_config_var("EXT_SUFFIX"))
    if os.path.exists(prebuilt):  # installed tree (tools/install_reference.py): the module imports like any other
        import importlib

        return importlib.import_module(modname)
    out = os.environ.get("EWB_REFBUILD_DIR", "/tmp/ewb_refbuild")
    os.makedirs(out, exist_ok=True)
    base = modname.split(".")[-1]
    so = os.path.join(out, base + sysconfig.get_config_var("EXT_SUFFIX"))
    if not os.path.exists(so):
        src = os.path.join(REFERENCE_ROOT, relpath)
        cpp = os.path.join(out, base + ".cpp")
        subprocess.check_call(
            [sys.executable, "-m", "cython", "-3", "--cplus", "-X", "boundscheck=False", "-X", "wraparound=False",
             "-X", "nonecheck=False", "-X", "initializedcheck=False", "-I", REFERENCE_ROOT, src, "-o", cpp]
        )
        cmd = ["/usr/bin/g++", "-O3", "-fPIC", "-shared", "-I" + sysconfig.get_paths()["include"],
               "-I" + numpy.get_include(), cpp, "-o", so]
        if openmp:
            cmd.insert(1, "-fopenmp")
        subprocess.check_call(cmd)
    spec = importlib.util.spec_from_file_location(modname, so)
    mod = importlib.util.module_from_spec(spec)
    sys.modules[modname] = mod
    spec.loader.exec_module(mod)
    return mod


def csr_generator_class():
    bootstrap()
    return build_cython_module("edelweissfe/numerics/csrgenerator.pyx", "edelweissfe.numerics.csrgenerator").CSRGenerator


def build_native_helpers():
    """The two dependency-free Cython modules full jobs need."""
    bootstrap()
    build_cython_module("edelweissfe/numerics/csrgenerator.pyx", "edelweissfe.numerics.csrgenerator")
    build_cython_module("edelweissfe/utils/elementresultcollector.pyx", "edelweissfe.utils.elementresultcollector")
