"""Host side of the multi-GPU trace solve: partitioning the mesh's blocks across the GPUs of one node (one process
per GPU) and the tables the library needs about the cut.

Blocks are independent in M-tilde u and in the local solves; coupling is only through faces, each shared by two
blocks (global_curved.jl:525-554).  A face whose two blocks live on different ranks (a *cut face*) carries lambda on
both ranks, kept bitwise identical.  Everything that moves data -- the point-to-point exchange of the partial
Fbar^T contributions (one message per partner), the all-reduces of the CG scalars and of the coarse-level data, the
completion of D and of the face blocks at setup -- happens inside libhsbp on the context's NCCL communicator
(hsbp_comm_init, hsbp_trace_set_partition, hsbp_trace_solve); this file only computes who owns what.

tests/dist_model.py holds a numpy model of the library's algorithm that runs the same tables over gloo on CPU.
"""
from dataclasses import dataclass
from typing import Dict, List

import numpy as np

from . import host


# ---- partitioning ----------------------------------------------------------------------------------
def partition_contiguous(nblocks, world):
    """owner[e] for contiguous, equally sized ranges of blocks."""
    per = -(-nblocks // world)
    return np.minimum(np.arange(nblocks) // per, world - 1).astype(np.int64)


@dataclass
class LocalMesh:
    blocks: np.ndarray          # global ids (0-based) of the local blocks, increasing
    faces: np.ndarray           # global ids (0-based) of the faces touched by local blocks, increasing
    EToF: np.ndarray            # 4 x nlocal, local face ids, 1-based
    FToB: np.ndarray
    FToE: np.ndarray            # 2 x nfaces_local, local block ids 1-based, 0 = on another rank
    FToLF: np.ndarray
    EToO: np.ndarray
    EToS: np.ndarray
    cut: Dict[int, List[int]]   # partner rank -> local face ids (0-based) of the cut faces, by increasing global id
    owned: np.ndarray           # per local face: this rank counts it in inner products
    gamma: np.ndarray           # per local face: index among ALL cut faces of the mesh (by global face id), -1 if not cut
    n_gamma: int                # number of cut faces of the whole mesh

    def partition_arrays(self):
        """(faces, partner, gamma, n_gamma) as hsbp_trace_set_partition takes them: 1-based local ids of the cut faces"""
        faces, partner = [], []
        for q, fl in sorted(self.cut.items()):
            faces += [i + 1 for i in fl]
            partner += [q] * len(fl)
        faces = np.asarray(faces, dtype=np.int64)
        return faces, np.asarray(partner, dtype=np.int64), self.gamma[faces - 1].astype(np.int64), int(self.n_gamma)


def localize(rank, owner, EToF, FToB, FToE, FToLF, EToO, EToS):
    """Connectivity of one rank's blocks in local numbering (reference conventions, 1-based ids in arrays)."""
    owner = np.asarray(owner)
    blocks = np.where(owner == rank)[0]
    g2l = -np.ones(len(owner), dtype=np.int64)
    g2l[blocks] = np.arange(len(blocks))
    faces = np.unique(EToF[:, blocks]) - 1
    f2l = {int(f): i for i, f in enumerate(faces)}
    lEToF = np.vectorize(lambda f: f2l[int(f) - 1] + 1)(EToF[:, blocks]).astype(np.int64)
    lFToE = np.zeros((2, len(faces)), dtype=np.int64)
    cut: Dict[int, List[int]] = {}
    owned = np.ones(len(faces), dtype=bool)
    has_lambda = lambda f: FToB[f] == host.BC_LOCKED_INTERFACE or FToB[f] >= host.BC_JUMP_INTERFACE
    # the cut faces of the whole mesh, numbered by increasing global face id: every rank computes the same table
    FToB_a = np.asarray(FToB)
    lam_mask = (FToB_a == host.BC_LOCKED_INTERFACE) | (FToB_a >= host.BC_JUMP_INTERFACE)
    both = lam_mask & (FToE[0] > 0) & (FToE[1] > 0)
    is_cut = np.zeros(len(FToB_a), dtype=bool)
    is_cut[both] = owner[FToE[0, both] - 1] != owner[FToE[1, both] - 1]
    gamma_of = -np.ones(len(FToB_a), dtype=np.int64)
    gamma_of[is_cut] = np.arange(int(is_cut.sum()))
    for i, f in enumerate(faces):
        for side in range(2):
            e = FToE[side, f] - 1
            if e >= 0 and owner[e] == rank:
                lFToE[side, i] = g2l[e] + 1
        if has_lambda(f):
            em, ep = FToE[0, f] - 1, FToE[1, f] - 1
            if owner[em] != owner[ep]:
                partner = int(owner[ep] if owner[em] == rank else owner[em])
                cut.setdefault(partner, []).append(i)
                owned[i] = owner[em] == rank                 # the minus side's rank counts the face
    return LocalMesh(blocks, faces, lEToF, FToB[faces].copy(), lFToE, FToLF[:, faces].copy(),
                     EToO[:, blocks].copy(), EToS[:, blocks].copy(), cut, owned, gamma_of[faces].copy(), int(is_cut.sum()))
