"""Multi-GPU plumbing (one process per GPU, torch.distributed): SURVEY.md section 8e.

Every MSM's (scalar, base) pairs are dealt round-robin to the ranks (pair i -> rank i mod world); each rank's context keeps
only its share of every query resident (g16_pk_load(rank, world)).  EC addition is exactly associative and commutative, so
the proof is bit-identical for any world size; NCCL has no user-defined reduction for curve points, hence gather-then-add.

Two exchanges exist:
  * in the library (default on GPUs, `ShardedProver(native=True)`): g16_comm_init + g16_prove_sharded -- the library issues
    ONE ncclAllGather of three XYZZ points per rank (A_k, B2_k, C_k = s A_k + r B1_k + L_k + H_k; 768 B on BLS12-381) on
    its own high-priority stream and finishes the proof; torch.distributed only carries the two NCCL unique ids (256 bytes) once.
  * host-plumbed (`native=False`; the gloo CPU tests, or a launcher without NCCL): g16_prove_partial -> all_gather of five
    affine points per rank through torch.distributed -> g16_prove_assemble.
"""
from __future__ import annotations

import numpy as np


def shard_indices(pairs: int, rank: int, world: int) -> slice:
    """Pairs of an MSM of `pairs` pairs owned by `rank`: rank, rank + world, ... -- must match Engine::shard in
    csrc/engine.cuh.  The split is interleaved rather than by contiguous range because the density of a query usually
    varies with the variable index (early variables are used more often): contiguous ranges would be unbalanced."""
    return slice(rank, pairs, world)


def all_gather_partials(partial: np.ndarray, device=None) -> np.ndarray:
    """partial: uint64 limbs of this rank's five partial points -> (world, limbs) array, rank order."""
    import torch
    import torch.distributed as dist
    world = dist.get_world_size()
    t = torch.from_numpy(np.ascontiguousarray(partial, dtype=np.uint64).view(np.int64))
    if device is not None:
        t = t.to(device)
    out = torch.empty(world * t.numel(), dtype=torch.int64, device=t.device)
    dist.all_gather_into_tensor(out, t)
    return out.cpu().numpy().view(np.uint64).reshape(world, -1)


def broadcast_unique_id(groth16, rank: int, device=None) -> np.ndarray:
    """rank 0 asks the library for an NCCL unique id (g16_comm_unique_id); torch.distributed broadcasts the 256 bytes"""
    import torch
    import torch.distributed as dist
    uid = groth16.comm_unique_id() if rank == 0 else np.zeros(256, dtype=np.uint8)
    t = torch.from_numpy(uid.copy())
    if device is not None:
        t = t.to(device)
    dist.broadcast(t, src=0)
    return t.cpu().numpy()


class ShardedProver:
    """Groth16 prover over `world` GPUs for one resident circuit + key."""

    def __init__(self, groth16, pk, matrices, rank: int, world: int, device=None, native: bool = True):
        self.g = groth16
        self.rank, self.world, self.device = rank, world, device
        self.native = native and world > 1
        if matrices is not None:
            self.g.load_matrices(matrices)
        if self.native:
            self.g.comm_init(broadcast_unique_id(self.g, rank, device), rank, world)
        self.g.load_proving_key(pk, rank, world)
        self._partial = np.zeros(self.g.partial_limbs(), dtype=np.uint64)
        self._proof = np.zeros(8 * self.g.nq, dtype=np.uint64)

    def _as_proof(self):
        from .api import Proof
        nq = self.g.nq
        return Proof(self._proof[:2 * nq].copy(), self._proof[2 * nq:6 * nq].copy(), self._proof[6 * nq:].copy())

    # pipelined form: submit(slot) returns at once; finish(slot) waits, gathers and assembles.  Submitting proof i+1
    # before finishing proof i overlaps the all_gather / host assembly of one proof with the GPU work of the next.
    def submit(self, slot: int, r, z_ptr: int, flags: int = 0, s=None):
        self._r_keep = getattr(self, "_r_keep", {})
        if self.native:
            if s is None:
                raise ValueError("the in-library exchange needs s at submission (it forms s*A_k + r*B1_k per rank)")
            self._r_keep[slot] = (self.g._fr_arg(r), self.g._fr_arg(s))
            self.g.prove_sharded_submit_raw(slot, self._r_keep[slot][0], self._r_keep[slot][1], z_ptr, flags)
            return
        self._r_keep[slot] = self.g._fr_arg(r)
        self.g.prove_partial_submit_raw(slot, self._r_keep[slot], z_ptr, flags)

    def finish(self, slot: int, r, s):
        if self.native:
            self.g.prove_sharded_wait_raw(slot, self._proof)
            return self._as_proof()
        self.g.prove_partial_wait_raw(slot, self._partial)
        allp = all_gather_partials(self._partial, self.device) if self.world > 1 else self._partial[None, :]
        return self.g.prove_assemble(self.g._fr_arg(r), self.g._fr_arg(s), allp)

    def prove(self, r, s, z_ptr: int, flags: int = 0):
        """r, s: Montgomery limbs; z_ptr: address of the full assignment (host or device per `flags`)."""
        rl = self.g._fr_arg(r)
        if self.native:
            self.g.prove_sharded_raw(rl, self.g._fr_arg(s), z_ptr, flags, self._proof)
            return self._as_proof()
        self.g.prove_assemble_prepare(rl, s)     # (r, s)-only scalar multiplications overlap the GPU work and the gather
        self.g.prove_partial_raw(rl, z_ptr, flags, self._partial)
        allp = all_gather_partials(self._partial, self.device) if self.world > 1 else self._partial[None, :]
        return self.g.prove_assemble(rl, self.g._fr_arg(s), allp)
