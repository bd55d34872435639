"""Helpers for the multi-rank tests: run the ORACLE's solvers over a brick decomposition
with a pluggable exchange layer (in-process threads, or torch.distributed gloo/nccl), so
the decomposition + processor-patch logic used by the GPU path is checked against the
single-domain solve."""
import importlib
import threading

import numpy as np


def local_case(meshmod, n, nRanks, rank, kind="P", dims=None):
    mesh = meshmod.decompose(n, nRanks, rank, dims=dims)
    coef = meshmod.pressure_laplacian(mesh) if kind == "P" else meshmod.momentum_matrix(mesh)
    return mesh, coef


def global_case(meshmod, n, kind="P", dims=None):
    mesh = meshmod.hex_mesh(*(dims or (n, n, n)))
    coef = meshmod.pressure_laplacian(mesh) if kind == "P" else meshmod.momentum_matrix(mesh)
    return mesh, coef


def oracle_matrix(orc, mesh, coef):
    ps, fc = mesh.patch_start_facecells()
    nr = [p.neighbRank for p in mesh.coupled_patches()]
    a = orc.Addr(mesh.nCells, mesh.lower, mesh.upper, ps, fc, neighbRank=nr if nr else None)
    m = orc.Matrix(a, coef["diag"], coef["upper"], coef["lower"], coef["bou"], coef["int"])
    return a, m


class ThreadExchange:
    """All ranks live in one process (one thread each); halo and sums go through shared
    slots guarded by a barrier.  Sums are rank-ordered => reproducible."""

    def __init__(self, nRanks):
        self.n = nRanks
        self.bar = threading.Barrier(nRanks)
        self.box = [None] * nRanks
        self.sums = [None] * nRanks

    def comm(self, orc, rank, mesh, nCellsGlobal):
        nbrs = [p.neighbRank for p in mesh.coupled_patches()]

        def halo(send, starts):
            self.box[rank] = {nb: send[starts[i]:starts[i + 1]].copy() for i, nb in enumerate(nbrs)}
            self.bar.wait()
            recv = np.empty_like(send)
            for i, nb in enumerate(nbrs):
                recv[starts[i]:starts[i + 1]] = self.box[nb][rank]
            self.bar.wait()
            return recv

        def allsum(v):
            self.sums[rank] = v.copy()
            self.bar.wait()
            tot = np.zeros_like(v)
            for r in range(self.n):
                tot = tot + self.sums[r]
            self.bar.wait()
            return tot

        def gather(mine):
            self.sums[rank] = mine.copy()
            self.bar.wait()
            out = np.stack([self.sums[r] for r in range(self.n)])
            self.bar.wait()
            return out
        return orc.PyComm(halo, allsum, nCellsGlobal, gather, rank, self.n)


def run_threads(nRanks, fn):
    """fn(rank) -> result, one thread per rank; re-raises the first exception."""
    out = [None] * nRanks
    errs = []

    def work(r):
        try:
            out[r] = fn(r)
        except BaseException as e:  # noqa: BLE001
            errs.append(e)
            raise
    th = [threading.Thread(target=work, args=(r,)) for r in range(nRanks)]
    for t in th:
        t.start()
    for t in th:
        t.join()
    if errs:
        raise errs[0]
    return out


def torch_comm(orc, mesh, nCellsGlobal, device="cpu"):
    """Exchange layer over torch.distributed (gloo on CPU)."""
    import torch
    import torch.distributed as dist
    rank = dist.get_rank()
    world = dist.get_world_size()
    nbrs = [p.neighbRank for p in mesh.coupled_patches()]

    def halo(send, starts):
        recv = np.empty_like(send)
        reqs, bufs = [], []
        for i, nb in enumerate(nbrs):
            s = torch.from_numpy(send[starts[i]:starts[i + 1]].copy())
            r = torch.empty(int(starts[i + 1] - starts[i]), dtype=torch.float64)
            reqs.append(dist.isend(s, nb, tag=rank))
            reqs.append(dist.irecv(r, nb, tag=nb))
            bufs.append((i, r, s))
        for q in reqs:
            q.wait()
        for i, r, _ in bufs:
            recv[starts[i]:starts[i + 1]] = r.numpy()
        return recv

    def _allgather(v):
        t = torch.from_numpy(np.ascontiguousarray(v).copy())
        parts = [torch.empty_like(t) for _ in range(world)]
        dist.all_gather(parts, t)
        return [p.numpy() for p in parts]

    def allsum(v):
        tot = np.zeros_like(v)
        for p in _allgather(v):  # rank-ordered sum
            tot = tot + p
        return tot

    def gather(mine):
        return np.stack(_allgather(mine))
    return orc.PyComm(halo, allsum, nCellsGlobal, gather, rank, world)
