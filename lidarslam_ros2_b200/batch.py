"""Batched loop-closure sweep: independent (scan, submap) registrations sharded one pair per GPU at a time, with ONE
all-gather of the resulting poses.

The reference evaluates a single revisit candidate per timer tick (graph_based_slam_component.cpp:187-233: nearest
submap that passes the travelled-distance / range gates → align → getFitnessScore → threshold). Registering ALL gated
candidates is the same computation repeated on independent inputs, so it shards with no data-path collective: pair i
goes to rank i mod world; every rank builds its own targets, runs its own solves, and the only exchange is the
all-gather of the fixed-size result rows so that every rank (and the host feeding the pose graph) sees all poses.
"""
from __future__ import annotations

import numpy as np

ROW = 20  # 16 pose floats (row-major 4x4) + fitness + converged + iterations + pair index


def shard_pairs(n_pairs: int, rank: int, world: int) -> list[int]:
    """Pair indices owned by `rank`: i mod world == rank (round-robin keeps per-rank work balanced)."""
    return [i for i in range(n_pairs) if i % world == rank]


def pack_row(index: int, T: np.ndarray, fitness: float, converged: bool, iterations: int) -> np.ndarray:
    row = np.zeros(ROW, dtype=np.float32)
    row[:16] = np.asarray(T, dtype=np.float32).reshape(16)
    row[16] = fitness
    row[17] = 1.0 if converged else 0.0
    row[18] = iterations
    row[19] = index
    return row


def unpack_rows(rows: np.ndarray):
    rows = np.asarray(rows, dtype=np.float32).reshape(-1, ROW)
    order = np.argsort(rows[:, 19], kind="stable")
    rows = rows[order]
    return {
        "index": rows[:, 19].astype(np.int64),
        "pose": rows[:, :16].reshape(-1, 4, 4).copy(),
        "fitness": rows[:, 16].copy(),
        "converged": rows[:, 17] > 0.5,
        "iterations": rows[:, 18].astype(np.int64),
    }


class RowComm:
    """The sweep's collective from C (include/b200comm.h): ncclGetUniqueId on rank 0, the 128-byte id handed to the other
    ranks out of band (here: torch.distributed's rendezvous — only as the bootstrap channel), ncclCommInitRank, and then
    ONE ncclAllGather of the packed result rows per sweep, issued by libb200reg.so itself."""

    def __init__(self, rank: int, world: int, device: int):
        import ctypes as C

        import torch
        import torch.distributed as dist

        from . import _capi

        self._lib, self._C = _capi.lib(), C
        ident = (C.c_ubyte * 128)()
        if rank == 0:
            rc = self._lib.b200comm_unique_id(ident)
            if rc != 0:
                raise RuntimeError("b200comm_unique_id: " + self._lib.b200comm_last_error().decode())
        if world > 1:
            t = torch.tensor(list(ident), dtype=torch.uint8, device=torch.device("cuda", device) if dist.get_backend() == "nccl" else None)
            dist.broadcast(t, src=0)
            ident = (C.c_ubyte * 128)(*[int(v) for v in t.cpu().tolist()])
        h = C.c_void_p()
        rc = self._lib.b200comm_create(ident, int(rank), int(world), int(device), C.byref(h))
        if rc != 0:
            raise RuntimeError("b200comm_create: " + self._lib.b200comm_last_error().decode())
        self._h, self.rank, self.world = h, rank, world

    def __del__(self):
        if getattr(self, "_h", None):
            self._lib.b200comm_destroy(self._h)
            self._h = None

    def create_board(self, max_rows: int) -> "PoseBoard":
        """Collective: one pose board per rank, every rank maps all of them (b200comm_board_create)."""
        return PoseBoard(self, max_rows)

    def all_gather_rows(self, rows: np.ndarray) -> np.ndarray:
        rows = np.ascontiguousarray(rows, dtype=np.float32)
        per, width = rows.shape
        out = np.empty((self.world * per, width), dtype=np.float32)
        rc = self._lib.b200comm_all_gather_rows(self._h, rows.ctypes.data, int(per), int(width), out.ctypes.data)
        if rc != 0:
            raise RuntimeError("b200comm_all_gather_rows: " + self._lib.b200comm_last_error().decode())
        return out


class PoseBoard:
    """include/b200comm.h's pose board: attached to an NDT handle (NormalDistributionsTransform.attachPoseBoard) it makes
    that handle's batch calls exchange their poses from inside the solver kernel (peer-memory stores over NVLink)."""

    def __init__(self, comm: RowComm, max_rows: int):
        C = comm._C
        self._lib, self._comm = comm._lib, comm  # the communicator must outlive the board's creation only; kept for clarity
        h = C.c_void_p()
        rc = self._lib.b200comm_board_create(comm._h, int(max_rows), C.byref(h))
        if rc != 0:
            raise RuntimeError("b200comm_board_create: " + self._lib.b200comm_last_error().decode())
        self._h, self.rank, self.world, self.max_rows = h, comm.rank, comm.world, int(max_rows)

    def close(self):
        if getattr(self, "_h", None):
            self._lib.b200comm_board_destroy(self._h)
            self._h = None

    def __del__(self):
        self.close()


def gather_rows(local_rows: np.ndarray, n_pairs: int, rank: int, world: int, device=None, comm: "RowComm | None" = None):
    """All-gather the per-rank result rows. Ranks own ceil/floor(n_pairs/world) pairs; rows are padded to the maximum
    count with index -1 and dropped after the gather. With `comm` the collective is the C-side ncclAllGather
    (b200comm_all_gather_rows); without it torch.distributed carries it (gloo in the CPU tests)."""
    per = (n_pairs + world - 1) // world
    if comm is not None and world > 1:
        buf = np.full((per, ROW), -1.0, dtype=np.float32)
        if len(local_rows):
            buf[:len(local_rows)] = np.asarray(local_rows, dtype=np.float32).reshape(-1, ROW)
        allr = comm.all_gather_rows(buf)
        return unpack_rows(allr[allr[:, 19] >= 0])
    import torch
    import torch.distributed as dist

    buf = torch.full((per, ROW), -1.0, dtype=torch.float32)
    if len(local_rows):
        buf[:len(local_rows)] = torch.from_numpy(np.asarray(local_rows, dtype=np.float32).reshape(-1, ROW))
    if device is not None:
        buf = buf.to(device)
    if world == 1:
        allr = buf.cpu().numpy()
    else:
        out = [torch.empty_like(buf) for _ in range(world)]
        dist.all_gather(out, buf)
        allr = torch.cat(out).cpu().numpy()
    allr = allr[allr[:, 19] >= 0]
    return unpack_rows(allr)


def register_pair(engine, source, target, guess=None, fitness_max_range=None):
    """One loop-closure candidate: setInputTarget → setInputSource → align → getFitnessScore
    (graph_based_slam_component.cpp:181, 227-231)."""
    engine.setInputTarget(target)
    engine.setInputSource(source)
    T = engine.align(guess)
    fit = engine.getFitnessScore() if fitness_max_range is None else engine.getFitnessScore(fitness_max_range)
    return T, fit, engine.hasConverged(), engine.getFinalNumIteration()


class LoopSweep:
    """The loop-closure candidate sweep of one rank: its share of the (scan, submap) pairs, one after the other through the
    node's own call sequence (graph_based_slam_component.cpp:181, 227-231). Returns the packed result rows."""

    def __init__(self, m, device: int = 0, resolution: float = 2.0, max_iterations: int = 100, transformation_epsilon: float = 0.01):
        self.ndt = m.NormalDistributionsTransform(device=device)
        self.ndt.setResolution(resolution)
        self.ndt.setTransformationEpsilon(transformation_epsilon)
        self.ndt.setMaximumIterations(max_iterations)  # graph_based_slam_component.cpp:66
        self.ndt.setNeighborhoodSearchMethod(m.DIRECT7)

    def kernel_launches(self) -> int:
        return int(self.ndt.stats()["kernel_launches"])

    def run(self, sources, targets, indices) -> np.ndarray:
        """b200reg_ndt_sweep: the pairs dealt round-robin to up to four internal engines (one host thread and one stream each)."""
        if len(indices) == 0:
            return np.zeros((0, ROW), dtype=np.float32)
        r = self.ndt.sweep(sources, targets)
        rows = np.zeros((len(indices), ROW), dtype=np.float32)
        rows[:, :16] = r["pose"].reshape(-1, 16)
        rows[:, 16] = r["fitness"]
        rows[:, 17] = r["converged"] != 0
        rows[:, 18] = r["iterations"]
        rows[:, 19] = np.asarray(indices, dtype=np.float32)
        return rows

    def run_sequential(self, sources, targets, indices) -> np.ndarray:
        """The same pairs one after the other through the public single-pair calls (reference for the parity test)."""
        rows = [pack_row(i, *register_pair(self.ndt, s, t)) for s, t, i in zip(sources, targets, indices)]
        return np.array(rows, dtype=np.float32).reshape(-1, ROW)
