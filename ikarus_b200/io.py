"""Host mirror of Ikarus::ResultFunction (ikarus/io/resultfunction.hh:57-157) over the device's calculateAt.

The reference wraps an assembler for a VTK writer and evaluates `fe.calculateAt<RT>(requirement(), local)` one element
and one local position at a time (resultfunction.hh:143-148).  A writer asks for the same few local positions -- the
element vertices -- on every element, so the first request for a position evaluates ALL elements on the device
(`ikb_calculate_at`) and the following ones are served from that table; the tables are dropped when the bound
requirement's solution or load factor changes."""
import itertools

import numpy as np

from .assembler import ResultTypes


class ResultFunction:
    """`makeResultFunction<RT>(assembler[, userFunction])` (resultfunction.hh:172-195).

    userFunction(resultArray, pos, fe_index, comp) -> float, with optional attributes/methods `ncomps()` and `name()`
    (resultfunction.hh:100-121); the default returns resultArray[comp] (resultfunction.hh:19-37)."""

    def __init__(self, assembler, resultType, userFunction=None):
        self._asm = assembler
        self._rt = ResultTypes(resultType)
        self._uf = userFunction
        self._tables = {}
        self._state = None

    # -- the Dune::VTKFunction interface (resultfunction.hh:77-121)
    def evaluate(self, comp, elementIndex, local):
        local = tuple(float(v) for v in np.asarray(local, float).ravel())
        r = self._table(local)[int(elementIndex)]
        if self._uf is None:
            return float(r[comp])
        return float(self._uf(r, np.asarray(local), int(elementIndex), comp))

    def ncomps(self):
        if self._uf is not None and hasattr(self._uf, "ncomps"):
            return int(self._uf.ncomps())
        dim = self._asm.finiteElements().dim
        full = self._rt in (ResultTypes.linearStressFull, ResultTypes.PK2StressFull)
        return 6 if full else dim * (dim + 1) // 2

    def name(self):
        if self._uf is not None and hasattr(self._uf, "name"):
            return str(self._uf.name())
        return self._rt.name

    # -- what a vertex-data writer consumes in one go
    def vertexData(self):
        """[nElem, 2^dim, ncomps()]: the function at the corners of every element (DUNE corner order, x fastest):
        what `vtkWriter.addVertexData(resultFunction)` samples before averaging per vertex."""
        dim = self._asm.finiteElements().dim
        corners = [tuple(float((c >> k) & 1) for k in range(dim)) for c in range(1 << dim)]
        n = len(self._asm.finiteElements())
        out = np.empty((n, len(corners), self.ncomps()))
        for q, xi in enumerate(corners):
            t = self._table(xi)
            if self._uf is None:
                out[:, q, :] = t
            else:
                for e, c in itertools.product(range(n), range(self.ncomps())):
                    out[e, q, c] = self._uf(t[e], np.asarray(xi), e, c)
        return out

    def _table(self, local):
        req = self._asm.requirement()
        d = req.globalSolution()
        state = (req.parameter(), d.tobytes())
        if state != self._state:
            self._tables.clear()
            self._state = state
        t = self._tables.get(local)
        if t is None:
            t = self._asm.calculateAt(self._rt, req, np.asarray(local))[:, 0, :]
            self._tables[local] = t
        return t


def makeResultFunction(assembler, resultType, userFunction=None):
    return ResultFunction(assembler, resultType, userFunction)
