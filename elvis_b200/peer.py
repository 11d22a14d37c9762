"""Peer-memory transport of the frame-sharded path: the halo exchange as copy-engine
peer copies over NVLink and the global min / max normalisations as one-shot mailbox all-reduces, both
through the C ABI (include/elvis_b200.h, csrc/peer.cu).  torch.distributed is used once, to pass the
CUDA IPC handles around; after that no NCCL kernel runs on the data path.

    group = PeerGroup(rank, world, device)            # collective: every rank constructs it
    clip = group.halo_clip(n_local, H, W)             # collective: a HaloClip whose buffer the neighbours can write
    ... fill clip.owned ...
    scores = sharding.sharded_removability(clip, T, bs, alpha, beta, rank, world, transport=group)

Protocol of one halo exchange (sequence number k = 1, 2, ... per clip):
    sender    wait until the neighbour has consumed halo k-1 (ack flag in MY memory), then copy my edge frame
              into the neighbour's halo slot and store k into its "arrived" flag;
    receiver  wait for "arrived" >= k before scoring; after the scoring kernel store k into the sender's ack flag.
Everything is enqueued on CUDA streams; nothing blocks the host.
"""
from __future__ import annotations

import ctypes as C
from typing import List, Optional

import torch
import torch.distributed as dist

from . import _lib
from ._lib import F32, F64, call
from .sharding import HaloClip

_FLAG_BYTES = 256          # flag words of a halo clip, kept apart from the pixels


def _flag_offset(n_local: int, height: int, width: int) -> int:
    """Byte offset of a halo clip's flag words: behind the (n + 2) frames, 16-byte aligned."""
    return ((n_local + 2) * height * width + 15) & ~15


class _DeviceMemory:
    """__cuda_array_interface__ view of a raw device allocation, so that torch can alias it."""

    def __init__(self, ptr: int, nbytes: int):
        self.__cuda_array_interface__ = {"shape": (nbytes,), "typestr": "|u1", "data": (ptr, False), "version": 3, "strides": None}


def _stream() -> C.c_void_p:
    return C.c_void_p(torch.cuda.current_stream().cuda_stream)


class PeerBuffer:
    """A cudaMalloc'ed, zero-filled buffer with an IPC handle, plus the peers' mappings of theirs."""

    def __init__(self, nbytes: int, device: torch.device):
        self.device, self.nbytes = device, nbytes
        ptr, handle = C.c_void_p(), C.create_string_buffer(64)
        with torch.cuda.device(device):
            call("elvis_peer_alloc", nbytes, C.byref(ptr), handle)
        self.ptr, self.handle = ptr.value, handle.raw
        self.tensor = torch.as_tensor(_DeviceMemory(self.ptr, nbytes), device=device)     # uint8 alias, not owning
        self.peers: dict = {}                                                            # rank -> mapped device pointer

    def open_peer(self, rank: int, handle: bytes) -> int:
        ptr = C.c_void_p()
        with torch.cuda.device(self.device):
            call("elvis_peer_open", C.c_char_p(handle), C.byref(ptr))
        self.peers[rank] = ptr.value
        return ptr.value

    def close(self) -> None:
        with torch.cuda.device(self.device):
            torch.cuda.synchronize(self.device)
            for p in self.peers.values():
                _lib.lib.elvis_peer_close(C.c_void_p(p))
            self.peers.clear()
            if self.ptr:
                self.tensor = None
                _lib.lib.elvis_peer_free(C.c_void_p(self.ptr))
                self.ptr = 0


class PeerHaloClip(HaloClip):
    """HaloClip whose buffer lives in IPC-shareable memory: the neighbours copy their edge frames straight into
    the halo slots.  Flag words behind the pixels: [0] left halo arrived, [1] right halo arrived (written by the
    neighbours), [2] my first frame was consumed by the left neighbour, [3] my last frame by the right one."""

    def __init__(self, n_local: int, height: int, width: int, device, buffer: PeerBuffer):
        frame = height * width
        pixels = buffer.tensor[:(n_local + 2) * frame].view(n_local + 2, height, width)
        super().__init__(n_local, height, width, device, buf=pixels)
        self.buffer, self.frame_bytes = buffer, frame
        self.flag_offset = _flag_offset(n_local, height, width)
        self.seq = 0

    def flag_ptr(self, base: int, index: int) -> C.c_void_p:
        return C.c_void_p(base + self.flag_offset + 4 * index)


class PeerGroup:
    """The ranks of one box that share memory over NVLink.  Construction and halo_clip() are collective."""

    def __init__(self, rank: int, world: int, device, group=None):
        self.rank, self.world, self.group = rank, world, group
        self.device = torch.device(device)
        if world > 16:
            raise ValueError("the mailbox all-reduce supports at most 16 ranks")
        self.buffers: List[PeerBuffer] = []
        self.mailbox = self._collective(lambda: PeerBuffer(int(_lib.lib.elvis_peer_mailbox_bytes()), self.device), "allocating the mailbox")
        self.buffers.append(self.mailbox)
        handles = self._all_gather(self.mailbox.handle)
        ptrs = self._collective(lambda: [self.mailbox.ptr if r == rank else self.mailbox.open_peer(r, handles[r]) for r in range(world)],
                                "mapping the peers' mailboxes (CUDA IPC)")
        self._mail_ptrs = (C.c_void_p * world)(*ptrs)
        self.error = torch.zeros(1, dtype=torch.int32, device=self.device)
        self.calls = 0
        self.barrier()

    def _collective(self, fn, what: str):
        """Run a local step that may fail and agree on the outcome, so that no rank is left waiting in the next
        collective: every rank raises if any rank failed."""
        result, ok = None, 1
        try:
            result = fn()
        except Exception as e:        # noqa: BLE001
            ok, result = 0, e
        flag = torch.tensor([ok], dtype=torch.int32, device=self.device)
        dist.all_reduce(flag, op=dist.ReduceOp.MIN, group=self.group)
        if int(flag.item()) != 1:
            raise RuntimeError(f"peer-memory setup failed while {what}" + (f": {result}" if ok == 0 else " on another rank"))
        return result

    # ------------------------------------------------------------------ plumbing over torch.distributed
    def _all_gather(self, obj):
        out = [None] * self.world
        dist.all_gather_object(out, obj, group=self.group)
        return out

    def barrier(self) -> None:
        torch.cuda.synchronize(self.device)
        dist.barrier(group=self.group)

    # ------------------------------------------------------------------ buffers
    def halo_clip(self, n_local: int, height: int, width: int) -> PeerHaloClip:
        if n_local < 1:
            raise ValueError("every rank needs at least one frame")
        buf = self._collective(lambda: PeerBuffer(_flag_offset(n_local, height, width) + _FLAG_BYTES, self.device), "allocating a halo clip")
        self.buffers.append(buf)
        clip = PeerHaloClip(n_local, height, width, self.device, buf)
        info = self._all_gather((buf.handle, n_local))
        clip.left = clip.right = None

        def map_neighbours():
            for nb, side in ((self.rank - 1, "left"), (self.rank + 1, "right")):
                if 0 <= nb < self.world:
                    handle, n_nb = info[nb]
                    base = buf.open_peer(nb, handle)
                    setattr(clip, side, {"base": base, "n": n_nb, "flag_offset": _flag_offset(n_nb, height, width)})
        self._collective(map_neighbours, "mapping the neighbours' halo clips (CUDA IPC)")
        self.barrier()
        return clip

    # ------------------------------------------------------------------ halo exchange
    def exchange_halo(self, clip: PeerHaloClip) -> None:
        """Push my edge frames into the neighbours' halo slots (current stream).  The matching wait is
        wait_halo(); acknowledge with release_halo() once the frames have been consumed."""
        clip.seq += 1
        k, st, fb = clip.seq, _stream(), clip.frame_bytes
        err = C.c_void_p(self.error.data_ptr())
        own = clip.buffer.ptr
        if clip.left is not None:       # my first frame -> the left neighbour's right halo slot (index n+1), its flag [1]
            nb = clip.left
            if k > 1:
                call("elvis_peer_wait", clip.flag_ptr(own, 2), k - 1, err, st)
            call("elvis_peer_put", C.c_void_p(nb["base"] + (nb["n"] + 1) * fb), C.c_void_p(own + fb), fb,
                 C.c_void_p(nb["base"] + nb["flag_offset"] + 4), k, st)
        if clip.right is not None:      # my last frame -> the right neighbour's left halo slot (index 0), its flag [0]
            nb = clip.right
            if k > 1:
                call("elvis_peer_wait", clip.flag_ptr(own, 3), k - 1, err, st)
            call("elvis_peer_put", C.c_void_p(nb["base"]), C.c_void_p(own + clip.n * fb), fb,
                 C.c_void_p(nb["base"] + nb["flag_offset"]), k, st)

    def wait_halo(self, clip: PeerHaloClip) -> None:
        """Make the current stream wait until both halo frames of the latest exchange have landed."""
        k, st, err, own = clip.seq, _stream(), C.c_void_p(self.error.data_ptr()), clip.buffer.ptr
        if clip.left is not None:
            call("elvis_peer_wait", clip.flag_ptr(own, 0), k, err, st)
        if clip.right is not None:
            call("elvis_peer_wait", clip.flag_ptr(own, 1), k, err, st)

    def release_halo(self, clip: PeerHaloClip) -> None:
        """Tell the neighbours (stream ordered, i.e. after the kernels that read the halo slots) that their
        frames of the latest exchange have been consumed and the slots may be overwritten."""
        k, st = clip.seq, _stream()
        if clip.left is not None:       # the left neighbour's LAST frame sits in my slot 0: its ack flag is [3]
            call("elvis_peer_signal", C.c_void_p(clip.left["base"] + clip.left["flag_offset"] + 12), k, st)
        if clip.right is not None:      # the right neighbour's FIRST frame sits in my slot n+1: its ack flag is [2]
            call("elvis_peer_signal", C.c_void_p(clip.right["base"] + clip.right["flag_offset"] + 8), k, st)

    # ------------------------------------------------------------------ min / max all-reduce
    def allreduce_minmax_(self, mm: torch.Tensor) -> torch.Tensor:
        """In-place exact all-reduce of interleaved {min, max, ...} (float32 or float64, <= 8 values, CUDA,
        contiguous) on the current stream.  Calls must be stream ordered on every rank."""
        if not mm.is_cuda or not mm.is_contiguous() or mm.numel() > 8:
            raise ValueError("mm must be a contiguous CUDA tensor of at most 8 values")
        dt = F32 if mm.dtype == torch.float32 else F64 if mm.dtype == torch.float64 else None
        if dt is None:
            raise TypeError("mm must be float32 or float64")
        self.calls += 1
        call("elvis_peer_allreduce_minmax", C.c_void_p(mm.data_ptr()), dt, mm.numel(), self.rank, self.world, self._mail_ptrs,
             self.calls % 4, self.calls, C.c_void_p(self.error.data_ptr()), _stream())
        return mm

    def check(self) -> None:
        """Raise if any wait of this rank timed out (synchronises the device)."""
        if int(self.error.item()):
            raise RuntimeError("a peer-memory wait timed out: a neighbouring rank never arrived")

    def close(self) -> None:
        self.barrier()
        for b in reversed(self.buffers):
            b.close()
        self.buffers = []


def try_create(rank: int, world: int, device, group=None) -> Optional[PeerGroup]:
    """PeerGroup if every rank can set it up (CUDA IPC between the GPUs of the box), else None on every rank."""
    try:
        return PeerGroup(rank, world, device, group)      # its setup steps agree across the ranks: all succeed or all raise
    except RuntimeError:
        return None
