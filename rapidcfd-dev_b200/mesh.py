"""Synthetic hex-block meshes and matrices in OpenFOAM (lduAddressing) ordering.

Stand-in for blockMesh + decomposePar, which the reference does not ship (SURVEY.md
section 8d): an nx*ny*nz hex cavity with cell c = i + nx*j + nx*ny*k, internal faces
sorted by owner then neighbour (per owner: +x, +y, +z), six boundary patches
(xmin, xmax, ymin, ymax = moving lid, zmin, zmax) and a structured brick decomposition
whose cut faces become processor patches ordered by neighbour rank
(processorPolyPatch ordering).  Pure numpy; host-side harness, not on the GPU path.
"""
from dataclasses import dataclass, field

import numpy as np


@dataclass
class Patch:
    name: str
    faceCells: np.ndarray            # int32 local cell next to each patch face
    Sf: np.ndarray                   # (P,3) outward face area vectors
    kind: str = "wall"               # wall | processor
    neighbRank: int = -1
    # for processor patches: global index of the cell on the other side (test aid)
    nbrGlobalCells: np.ndarray = None


@dataclass
class HexMesh:
    nx: int
    ny: int
    nz: int
    h: float
    lower: np.ndarray                # int32 [F] owner
    upper: np.ndarray                # int32 [F] neighbour
    faceDir: np.ndarray              # int8  [F] 0/1/2 = x/y/z normal
    patches: list = field(default_factory=list)
    cellGlobal: np.ndarray = None    # local cell -> global cell (decomposed meshes)
    origin: tuple = (0, 0, 0)        # brick origin (i0, j0, k0) in the global mesh
    globalDims: tuple = None

    @property
    def nCells(self):
        return self.nx * self.ny * self.nz

    @property
    def nFaces(self):
        return len(self.lower)

    def cell_centres(self):
        i0, j0, k0 = self.origin
        c = np.arange(self.nCells)
        i = c % self.nx
        j = (c // self.nx) % self.ny
        k = c // (self.nx * self.ny)
        return np.stack([(i + i0 + 0.5) * self.h, (j + j0 + 0.5) * self.h, (k + k0 + 0.5) * self.h], axis=1)

    def volumes(self):
        return np.full(self.nCells, self.h ** 3)

    def Sf(self):
        s = np.zeros((self.nFaces, 3))
        s[np.arange(self.nFaces), self.faceDir] = self.h * self.h
        return s

    def magSf(self):
        return np.full(self.nFaces, self.h * self.h)

    def deltaCoeffs(self):
        return np.full(self.nFaces, 1.0 / self.h)

    def weights(self):
        return np.full(self.nFaces, 0.5)

    def face_centres(self):
        cc = self.cell_centres()
        return 0.5 * (cc[self.lower] + cc[self.upper])

    def coupled_patches(self):
        return [p for p in self.patches if p.kind == "processor"]

    def wall_patches(self):
        return [p for p in self.patches if p.kind != "processor"]

    def patch_start_facecells(self, patches=None):
        ps = self.coupled_patches() if patches is None else patches
        start = np.zeros(len(ps) + 1, dtype=np.int32)
        for i, p in enumerate(ps):
            start[i + 1] = start[i] + len(p.faceCells)
        fc = np.concatenate([p.faceCells for p in ps]).astype(np.int32) if ps else np.zeros(0, np.int32)
        return start, fc


def _brick_faces(nx, ny, nz):
    n = nx * ny * nz
    c = np.arange(n, dtype=np.int64)
    i = c % nx
    j = (c // nx) % ny
    k = c // (nx * ny)
    hx = (i < nx - 1)
    hy = (j < ny - 1)
    hz = (k < nz - 1)
    cnt = hx.astype(np.int64) + hy + hz
    start = np.zeros(n + 1, dtype=np.int64)
    np.cumsum(cnt, out=start[1:])
    F = int(start[-1])
    lower = np.empty(F, dtype=np.int32)
    upper = np.empty(F, dtype=np.int32)
    fdir = np.empty(F, dtype=np.int8)
    fx = start[:-1][hx]
    lower[fx] = c[hx]
    upper[fx] = c[hx] + 1
    fdir[fx] = 0
    fy = (start[:-1] + hx)[hy]
    lower[fy] = c[hy]
    upper[fy] = c[hy] + nx
    fdir[fy] = 1
    fz = (start[:-1] + hx + hy)[hz]
    lower[fz] = c[hz]
    upper[fz] = c[hz] + nx * ny
    fdir[fz] = 2
    return lower, upper, fdir


def _side_cells(nx, ny, nz, axis, hi):
    c = np.arange(nx * ny * nz, dtype=np.int64)
    idx = [c % nx, (c // nx) % ny, c // (nx * ny)][axis]
    lim = [nx, ny, nz][axis] - 1 if hi else 0
    return c[idx == lim].astype(np.int32)


_PATCH_NAMES = ["xmin", "xmax", "ymin", "movingWall", "zmin", "zmax"]


def hex_mesh(nx, ny=None, nz=None, length=1.0):
    """Single-domain nx*ny*nz cavity of edge length `length` (cells are cubes of h = length/nx)."""
    ny = nx if ny is None else ny
    nz = nx if nz is None else nz
    h = length / nx
    lower, upper, fdir = _brick_faces(nx, ny, nz)
    m = HexMesh(nx, ny, nz, h, lower, upper, fdir, globalDims=(nx, ny, nz))
    m.cellGlobal = np.arange(m.nCells, dtype=np.int64)
    for axis in range(3):
        for hi in (0, 1):
            fc = _side_cells(nx, ny, nz, axis, hi)
            sf = np.zeros((len(fc), 3))
            sf[:, axis] = (1.0 if hi else -1.0) * h * h
            m.patches.append(Patch(_PATCH_NAMES[2 * axis + hi], fc, sf, "wall"))
    return m


def brick_split(nRanks):
    """2 -> 2x1x1, 4 -> 2x2x1, 8 -> 2x2x2."""
    return {1: (1, 1, 1), 2: (2, 1, 1), 4: (2, 2, 1), 8: (2, 2, 2)}[nRanks]


def decompose(n, nRanks, rank, length=1.0, dims=None):
    """Local mesh of `rank` in a structured brick decomposition of the n^3 (or dims) cavity.

    Processor patches follow the physical wall patches and are ordered by neighbour
    rank; faces inside a processor patch are ordered by the *owner-side* (lower rank)
    global face order, which for these bricks equals ascending local cell index on both
    sides -- so both sides enumerate a patch identically."""
    gx, gy, gz = (n, n, n) if dims is None else dims
    px, py, pz = brick_split(nRanks)
    assert gx % px == 0 and gy % py == 0 and gz % pz == 0
    nx, ny, nz = gx // px, gy // py, gz // pz
    rx, ry, rz = rank % px, (rank // px) % py, rank // (px * py)
    h = length / gx
    lower, upper, fdir = _brick_faces(nx, ny, nz)
    m = HexMesh(nx, ny, nz, h, lower, upper, fdir, origin=(rx * nx, ry * ny, rz * nz), globalDims=(gx, gy, gz))
    c = np.arange(m.nCells, dtype=np.int64)
    gi = c % nx + rx * nx
    gj = (c // nx) % ny + ry * ny
    gk = c // (nx * ny) + rz * nz
    m.cellGlobal = gi + gx * gj + gx * gy * gk
    procs = []
    P = (px, py, pz)
    R = (rx, ry, rz)
    for axis in range(3):
        for hi in (0, 1):
            fc = _side_cells(nx, ny, nz, axis, hi)
            sf = np.zeros((len(fc), 3))
            sf[:, axis] = (1.0 if hi else -1.0) * h * h
            nb = list(R)
            nb[axis] += 1 if hi else -1
            if 0 <= nb[axis] < P[axis]:
                nbRank = nb[0] + px * nb[1] + px * py * nb[2]
                stride = [1, gx, gx * gy][axis]
                nbrG = m.cellGlobal[fc] + (stride if hi else -stride)
                procs.append(Patch(f"procBoundary{rank}to{nbRank}", fc, sf, "processor", nbRank, nbrG))
            else:
                m.patches.append(Patch(_PATCH_NAMES[2 * axis + hi], fc, sf, "wall"))
    procs.sort(key=lambda p: p.neighbRank)
    m.patches.extend(procs)
    return m


# ---------------------------------------------------------------------------
# synthetic matrices
# ---------------------------------------------------------------------------

def _face_field_global(mesh, seed, lo, hi):
    """Per-internal-face random field that is identical for the same *global* face in
    every decomposition (hash of the global owner cell and direction)."""
    g = mesh.cellGlobal[mesh.lower].astype(np.uint64)
    key = g * np.uint64(3) + mesh.faceDir.astype(np.uint64) + _seedmix(seed)
    return lo + (hi - lo) * _hash01(key)


def _seedmix(seed):
    return np.uint64((int(seed) * 0x9E3779B97F4A7C15) & 0xFFFFFFFFFFFFFFFF)


def _hash01(key):
    x = key.astype(np.uint64)
    x ^= x >> np.uint64(33)
    x *= np.uint64(0xFF51AFD7ED558CCD)
    x ^= x >> np.uint64(33)
    x *= np.uint64(0xC4CEB9FE1A85EC53)
    x ^= x >> np.uint64(33)
    return (x >> np.uint64(11)).astype(np.float64) / float(1 << 53)


def cell_field_global(mesh, seed, lo=-1.0, hi=1.0):
    key = mesh.cellGlobal.astype(np.uint64) + _seedmix(seed)
    return lo + (hi - lo) * _hash01(key)


def pressure_laplacian(mesh, seed=1234, vary=True, pin=True):
    """Pressure-equation matrix fvm::laplacian(rAU, p) with zeroGradient walls.

    upper = +rAU_f*|Sf|*delta (OpenFOAM sign: negative semi-definite), diag = -sum, the
    reference cell (global cell 0) pinned as fvMatrix::setReference does
    (FV/fvMatrices/fvMatrix/fvMatrix.C:965-983).  Processor patches get
    boundaryCoeffs = internalCoeffs = -rAU_f*|Sf|*delta (coupled gradient coeffs,
    gaussLaplacianScheme.C:74-79) with internalCoeffs added to the diagonal.
    Returns dict(diag, upper, lower=None, bou, int)."""
    hmag = mesh.h  # |Sf|*delta = h^2 * 1/h
    rAUf = _face_field_global(mesh, seed, 0.5, 1.5) if vary else np.ones(mesh.nFaces)
    upper = rAUf * hmag
    diag = np.zeros(mesh.nCells)
    np.subtract.at(diag, mesh.lower, upper)
    np.subtract.at(diag, mesh.upper, upper)
    bou, intc = [], []
    for p in mesh.coupled_patches():
        # coefficient keyed on the global face: owner = min(global cells), same direction
        gl = mesh.cellGlobal[p.faceCells]
        gown = np.minimum(gl, p.nbrGlobalCells).astype(np.uint64)
        axis = int(np.argmax(np.abs(p.Sf[0])))
        key = gown * np.uint64(3) + np.uint64(axis) + _seedmix(seed)
        r = (0.5 + _hash01(key)) if vary else np.ones(len(gl))
        coeff = r * hmag
        # coupled gradient coeffs: gradientInternalCoeffs = -delta, gradientBoundaryCoeffs = +delta
        # => internalCoeffs = -coeff, boundaryCoeffs = -(coeff) (gaussLaplacianScheme.C:74-79);
        # fvMatrix adds internalCoeffs to diag, Amul subtracts boundaryCoeffs*psi_nbr.
        np.subtract.at(diag, p.faceCells, coeff)
        bou.append(-coeff)
        intc.append(-coeff)
    if pin:
        c0 = np.nonzero(mesh.cellGlobal == 0)[0]
        if len(c0):
            diag[c0[0]] += diag[c0[0]]
    out = dict(diag=diag, upper=upper, lower=None)
    out["bou"] = np.concatenate(bou) if bou else np.zeros(0)
    out["int"] = np.concatenate(intc) if intc else np.zeros(0)
    return out


def momentum_matrix(mesh, nu=0.01, Co=0.5):
    """Asymmetric convection-diffusion matrix fvm::ddt + fvm::div(phi) - fvm::laplacian(nu)
    for the solenoidal field u = (sin(pi x) cos(pi y), -cos(pi x) sin(pi y), 0), linear
    weights w = 0.5 (gaussConvectionScheme.C:95-97, EulerDdtScheme.C:347-357)."""
    h = mesh.h
    fc = mesh.face_centres()
    ux = np.sin(np.pi * fc[:, 0]) * np.cos(np.pi * fc[:, 1])
    uy = -np.cos(np.pi * fc[:, 0]) * np.sin(np.pi * fc[:, 1])
    un = np.where(mesh.faceDir == 0, ux, np.where(mesh.faceDir == 1, uy, 0.0))
    phi = un * h * h
    dt = Co * h / 1.0
    lower = -0.5 * phi - nu * h
    upper = lower + phi
    diag = np.full(mesh.nCells, h ** 3 / dt)
    np.subtract.at(diag, mesh.lower, lower)
    np.subtract.at(diag, mesh.upper, upper)
    # wall patches: fixedValue U -> diffusion adds nu*|Sf|*2/h to the diagonal
    for p in mesh.wall_patches():
        np.add.at(diag, p.faceCells, nu * h * 2.0)
    bou, intc = [], []
    for p in mesh.coupled_patches():
        pc = mesh.cell_centres()[p.faceCells]
        axis = int(np.argmax(np.abs(p.Sf[0])))
        sgn = np.sign(p.Sf[0, axis])
        pf = pc.copy()
        pf[:, axis] += sgn * 0.5 * h
        uxp = np.sin(np.pi * pf[:, 0]) * np.cos(np.pi * pf[:, 1])
        uyp = -np.cos(np.pi * pf[:, 0]) * np.sin(np.pi * pf[:, 1])
        unp = [uxp, uyp, np.zeros(len(pf))][axis] * sgn
        pphi = unp * h * h   # outward flux
        # coupled: valueInternalCoeffs = w, valueBoundaryCoeffs = 1-w ; gradient coeffs +-delta
        internal = pphi * 0.5 + nu * h
        boundary = -pphi * 0.5 + nu * h
        np.add.at(diag, p.faceCells, internal)
        bou.append(boundary)
        intc.append(internal)
    out = dict(diag=diag, upper=upper, lower=lower)
    out["bou"] = np.concatenate(bou) if bou else np.zeros(0)
    out["int"] = np.concatenate(intc) if intc else np.zeros(0)
    return out


def face_area_pair_weights(mesh):
    """faceAreaPairGAMGAgglomeration face weights |Sf/sqrt|Sf| * (1, 1.01, 1.02)|
    (FV/fvMatrices/solvers/GAMGSymSolver/GAMGAgglomerations/faceAreaPairGAMGAgglomeration/
    faceAreaPairGAMGAgglomeration.C:56-78)."""
    sf = mesh.Sf()
    mag = np.sqrt((sf ** 2).sum(1))
    t = sf / np.sqrt(mag)[:, None] * np.array([1.0, 1.01, 1.02])
    return np.sqrt((t ** 2).sum(1))
