"""Dry-run aid for the GPU tests (NOT part of the product and not used by default): an object with the
surface of rapidcfd-dev_b200.capi backed by the CPU oracle and CPU torch tensors.

    B200LDU_DRYRUN_ORACLE=1 python -m pytest tests/test_ref_golden.py tests/test_zz_golden.py tests/test_zzz_fvm_gpu.py -m gpu -q

runs the `-m gpu` tests of those files with this stand-in instead of the CUDA library: it checks the
tests' own logic (indexing of the fixtures, shapes, iteration windows, tolerances) where no GPU is
available.  It says nothing about the CUDA path; on a GPU box the variable is unset and the real library
is used."""
import numpy as np

from oracle import fvm_oracle as fo
from oracle import ldu_oracle as orc


class _Ctx:
    device = "cpu"

    def close(self):
        pass


class LduAddressing:
    def __init__(self, ctx, nCells, lower, upper, patchStart=None, faceCells=None, neighbRank=None, cellCentres=None):
        self.nCells, self.nFaces = int(nCells), len(lower)
        self.o = orc.Addr(nCells, lower, upper) if patchStart is None else orc.Addr(nCells, lower, upper, patchStart,
                                                                                     faceCells, neighbRank=neighbRank)

    def close(self):
        pass


def _np(t):
    return None if t is None else t.detach().cpu().numpy()


class LduMatrix:
    def __init__(self, addr):
        self.addr = addr
        self.m = None

    def set(self, diag, upper, lower=None, bou=None, intc=None):
        self.m = orc.Matrix(self.addr.o, _np(diag), _np(upper), _np(lower), _np(bou), _np(intc))
        self.coeffs = (diag, upper, lower, bou, intc)
        return self

    def _t(self, a):
        import torch
        return torch.from_numpy(np.ascontiguousarray(a))

    def Amul(self, x):
        return self._t(self.m.amul(_np(x)))

    def Tmul(self, x):
        return self._t(self.m.tmul(_np(x)))

    def sumA(self, like):
        return self._t(self.m.sumA())

    def residual(self, x, b):
        return self._t(self.m.residual(_np(x), _np(b)))

    def H(self, x):
        return self._t(self.m.H(_np(x)))

    def H1(self, like):
        return self._t(self.m.H1())

    def faceH(self, x):
        return self._t(self.m.faceH(_np(x)))

    def precondition(self, name, r, transpose=False):
        return self._t(self.m.precondition(name, _np(r), transpose))

    def smooth(self, name, x, b, nSweeps, omega=0.9):
        return self._t(self.m.jacobi(_np(x), _np(b), nSweeps, omega=omega))

    def solve(self, solver, pre, psi, source, gamg=None, histCap=0, **ctl):
        if solver == "GAMG":
            out, perf, hist = gamg.o.solve(self.m, pre, _np(psi), _np(source), **ctl)
        else:
            out, perf, hist = self.m.solve(solver, pre, _np(psi), _np(source), **ctl)
        psi.copy_(self._t(out))
        return perf, (hist if histCap else hist[:0])

    def close(self):
        pass


class FvMatrix:
    """surface of capi.FvMatrix over oracle/fvm_oracle.py (dry runs of tests/test_zzz_fvm_gpu.py)"""

    def __init__(self, matrix, nComp, diag, source, psi, V, internalCoeffs=None, boundaryCoeffs=None):
        self.m, self.nc = matrix, int(nComp)
        self.diag, self.source, self.psi, self.V, self.ic, self.bc = diag, source, psi, V, internalCoeffs, boundaryCoeffs

    def _o(self, diag=None, source=None):
        _, upper, lower, bou, intc = self.m.coeffs
        a = self.m.addr
        return fo.FvMatrix(orc, a.o, self.nc, _np(self.diag if diag is None else diag), _np(upper), _np(lower),
                           _np(self.source if source is None else source), _np(self.psi), _np(self.V), a.bfc,
                           _np(self.ic), _np(self.bc), couInt=_np(intc), couBou=_np(bou))

    def _t(self, a):
        import torch
        return torch.from_numpy(np.ascontiguousarray(a).ravel())

    def _pnf(self, pnf):
        return None if pnf is None else _np(pnf).reshape(-1, self.nc)

    def addBoundaryDiag(self, diag, cmpt):
        d = _np(diag).copy()
        self._o().addBoundaryDiag(d, cmpt)
        diag.copy_(self._t(d))

    def addCmptAvBoundaryDiag(self, diag):
        d = _np(diag).copy()
        self._o().addCmptAvBoundaryDiag(d)
        diag.copy_(self._t(d))

    def addBoundarySource(self, source, pnf=None):
        s = _np(source).reshape(-1, self.nc).copy()
        self._o().addBoundarySource(s, pnf is not None, self._pnf(pnf))
        source.copy_(self._t(s))

    def A(self):
        return self._t(self._o().A())

    def H(self, pnf=None):
        return self._t(self._o().H(pnf=self._pnf(pnf)))

    def flux(self, nBFaces, nCoupledFaces=0, pnf=None):
        return tuple(self._t(x) for x in self._o().flux(self._pnf(pnf)))

    def residual(self, pnf=None):
        return self._t(self._o().residual(self._pnf(pnf)))

    def relax(self, alpha):
        o = self._o()
        o.relax(alpha)
        self.diag.copy_(self._t(o.diag))
        self.source.copy_(self._t(o.source))

    def setReference(self, celli, value):
        o = self._o()
        o.setReference(celli, value)
        self.diag.copy_(self._t(o.diag))
        self.source.copy_(self._t(o.source))

    def solve(self, solver, pre, gamg=None, pnf=None, **ctl):
        psi, perfs, _ = self._o().solve(solver, pre, gamg.o if gamg is not None else None, self._pnf(pnf), **ctl)
        self.psi.copy_(self._t(psi))
        return perfs


class GamgAgglomeration:
    def __init__(self, addr, faceWeights, nCellsInCoarsestLevel=10, mergeLevels=1, forward=1):
        self.o = orc.Gamg(addr.o, faceWeights, nCellsInCoarsestLevel, mergeLevels=mergeLevels, forward=forward)
        self.nLevels = self.o.nLevels

    @property
    def forward(self):
        return self.o.forward

    def level_size(self, lev):
        return self.o.ncells(lev), self.o.nfaces(lev)

    def restrict_addr(self, lev):
        return self.o.restrict_addr(lev)

    def close(self):
        pass


class _Capi:
    LduAddressing = LduAddressing
    LduMatrix = LduMatrix
    FvMatrix = FvMatrix

    @staticmethod
    def fv_boundary_set(addr, bFaceCells):
        addr.bfc = np.asarray(bFaceCells, np.int32).copy()
    GamgAgglomeration = GamgAgglomeration

    @staticmethod
    def Context(dev):
        return _Ctx()

    @staticmethod
    def mesh_to_device(ctx, mesh, with_centres=True):
        ps, fc = mesh.patch_start_facecells()
        nr = [p.neighbRank for p in mesh.coupled_patches()]
        if not nr:
            return LduAddressing(ctx, mesh.nCells, mesh.lower, mesh.upper)
        return LduAddressing(ctx, mesh.nCells, mesh.lower, mesh.upper, ps, fc, nr)

    @staticmethod
    def polymesh_to_device(ctx, pm, cellCentres=None):
        lo, up = pm.ldu()
        ps, fc, nr = pm.coupled_interface_arrays()
        return LduAddressing(ctx, pm.nCells, lo, up, ps, fc, nr)


def fixture():
    """what the `gpu` fixtures yield: (capi, ctx, torch)"""
    import torch
    return _Capi, _Ctx(), torch
