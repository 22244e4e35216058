"""Host plumbing of a multi-GPU run: one process per GPU (torchrun), `torch.distributed` only moves the CUDA IPC
handles and provides host barriers -- the basket exchange itself happens inside the kernels over NVLink peer
mappings (csrc/device_common.cuh: publish_slot, cross_gpu_barrier)."""
from __future__ import annotations

import ctypes as C

import numpy as np

from . import _lib
from .api import Session

IPC_BYTES = 192


def shard_owner(g, colour=None, hub_threshold: int = 0, world: int = 1) -> np.ndarray:
    """owner[v] = rank that updates node v (-1: sink). Host only (pprb200_shard_owner)."""
    lib = _lib.load()
    owner = np.full(max(g.n, 1), -1, dtype=np.int32)
    col_arr = None if colour is None else np.ascontiguousarray(colour, dtype=np.uint8)
    _lib.check(lib.pprb200_shard_owner(_lib.ptr(g.row_ptr), _lib.ptr(g.col), g.n, _lib.ptr(col_arr), hub_threshold, world,
                                       _lib.ptr(owner)))
    return owner[:g.n]


def export_handles(sess: Session) -> np.ndarray:
    blob = np.zeros(IPC_BYTES, dtype=np.uint8)
    _lib.check(sess.lib.pprb200_session_ipc_export(sess.handle, _lib.ptr(blob)))
    return blob


def attach_handles(sess: Session, all_blobs: np.ndarray) -> None:
    all_blobs = np.ascontiguousarray(all_blobs, dtype=np.uint8)
    _lib.check(sess.lib.pprb200_session_ipc_attach(sess.handle, _lib.ptr(all_blobs)))


def connect(sess: Session, dist) -> None:
    """all-gather the IPC handles over the default process group (gloo or nccl) and attach the peers"""
    import torch
    world = dist.get_world_size()
    if world == 1:
        return
    mine = torch.from_numpy(export_handles(sess).copy())
    dev = torch.device("cuda", torch.cuda.current_device()) if dist.get_backend() == "nccl" else torch.device("cpu")
    mine = mine.to(dev)
    out = [torch.empty_like(mine) for _ in range(world)]
    dist.all_gather(out, mine)
    attach_handles(sess, torch.stack(out).cpu().numpy())
    dist.barrier()


def need_mask(g, colour=None, hub_threshold: int = 0, world: int = 2):
    """(owner[v], need[v]): the rank that updates node v (-1: sink) and the ranks that read v's basket during the iterations
    (bit r: rank r owns a predecessor of v) -- where v's owner stores the basket. Host only (pprb200_debug_need_mask)."""
    lib = _lib.load()
    owner = np.zeros(max(g.n, 1), dtype=np.int32)
    need = np.zeros(max(g.n, 1), dtype=np.uint8)
    col_arr = None if colour is None else np.ascontiguousarray(colour, dtype=np.uint8)
    _lib.check(lib.pprb200_debug_need_mask(_lib.ptr(g.row_ptr), _lib.ptr(g.col), g.n, _lib.ptr(col_arr), hub_threshold, world,
                                           _lib.ptr(owner), _lib.ptr(need)))
    return owner[:g.n], need[:g.n]
