"""Drive the UNMODIFIED reference (imported from /root/reference through tools/refshim.py)
on its own element objects: BoxGen mesh -> DofManager -> VIJ -> NIST.computeElements loop
(nonlinearimplicitstatic.py:837-844) -> CSRGenerator.updateCSR.  Fixture generation and
CPU-side cross checks only (never on the product path / GPU box)."""
import numpy as np

from . import refshim  # noqa


def _materials():
    from edelweissfe.materials.linearelastic.linearelastic import LinearElasticMaterial
    from edelweissfe.materials.neohooke.neohookepencegouformulationa import NeoHookeanWaMaterial
    from edelweissfe.materials.neohooke.neohookepencegouformulationb import NeoHookeanWbMaterial
    from edelweissfe.materials.neohooke.neohookepencegouformulationc import NeoHookeanWcMaterial
    from edelweissfe.materials.vonmises.vonmises import VonMisesMaterial

    return dict(
        linearelastic=LinearElasticMaterial,
        vonmises=VonMisesMaterial,
        neohookewa=NeoHookeanWaMaterial,
        neohookewb=NeoHookeanWbMaterial,
        neohookewc=NeoHookeanWcMaterial,
    )


class RefModel:
    """Reference FEModel + DofManager + CSRGenerator for one BoxGen (or explicit) mesh."""

    def __init__(self, elType, material, props, box=None, nodes=None, conn=None):
        refshim.bootstrap()
        from edelweissfe.generators.abqmodelconstructor import AbqModelConstructor
        from edelweissfe.generators.boxgen import generateModelData
        from edelweissfe.journal.journal import Journal
        from edelweissfe.models.femodel import FEModel
        from edelweissfe.numerics.dofmanager import DofManager

        CSRGenerator = refshim.csr_generator_class()
        journal = Journal(verbose=False)
        model = FEModel(3)
        if box is not None:
            data = ["%s=%s" % kv for kv in box.items()] + ["elType=%s" % elType, "elProvider=edelweiss"]
            generateModelData({"name": "gen", "data": data}, model, journal)
            defs = {k: [] for k in ("*node", "*element", "*elSet", "*nSet", "*surface")}
        else:
            defs = {k: [] for k in ("*elSet", "*nSet", "*surface")}
            defs["*node"] = [{"data": [", ".join([str(i + 1)] + [repr(float(c)) for c in x]) for i, x in enumerate(nodes)]}]
            defs["*element"] = [
                {"type": elType, "provider": "edelweiss", "data": [", ".join([str(e + 1)] + [str(int(n) + 1) for n in c]) for e, c in enumerate(conn)]}
            ]
        AbqModelConstructor(journal).createGeometryFromInputFile(model, defs)
        model._prepareVariablesAndFields(journal)
        self.material = _materials()[material.lower()](np.asarray(props, dtype=float))
        for el in model.elements.values():
            el.initializeElement()
            el.setMaterial(self.material)
        self.model = model
        self.dm = DofManager(
            model.nodeFields.values(), model.scalarVariables.values(), model.elements.values(),
            model.constraints.values(), model.nodeSets.values(),
        )
        self.K = self.dm.constructVIJSystemMatrix()
        self.csrgen = CSRGenerator(self.K)
        self.elements = list(model.elements.values())

    # --- layout contract -------------------------------------------------------------
    def coords(self):
        return np.array([n.coordinates for n in self.model.nodes.values()], dtype=float)

    def connectivity(self):
        labels = {lab: i for i, lab in enumerate(self.model.nodes.keys())}
        return np.array([[labels[n.label] for n in el.nodes] for el in self.elements], dtype=np.int32)

    def element_dofs(self):
        return np.array([self.dm.idcsOfElementsInDofVector[el] for el in self.elements], dtype=np.int64)

    def set_state(self, stateRef):
        for el, s in zip(self.elements, stateRef):
            el._stateVarsRef[:] = s

    def assemble(self, U, dU, time=(0.0, 0.0), dT=1.0):
        """One pass of the loop at nonlinearimplicitstatic.py:837-844 + updateCSR (:753-769)."""
        dm = self.dm
        K = self.K
        K[:] = 0.0
        Uv, dUv, P, F = (dm.constructDofVector() for _ in range(4))
        Uv[:] = U
        dUv[:] = dU
        t = np.array(time, dtype=float)
        for el in self.elements:
            Ke = K[el]
            Pe = np.zeros(el.nDof)
            el.computeYourself(Ke, Pe, Uv[el], dUv[el], t, dT)
            P[el] += Pe
            F[el] += abs(Pe)
        csr = self.csrgen.updateCSR(K)
        state = np.array([np.array(el._stateVarsTemp) for el in self.elements])
        return dict(
            V=np.array(K), I=np.array(K.I), J=np.array(K.J), indptr=csr.indptr.copy(), indices=csr.indices.copy(),
            data=csr.data.copy(), P=np.array(P), F=np.array(F), stateTemp=state,
        )

    def accept(self):
        for el in self.elements:
            el.acceptLastState()
