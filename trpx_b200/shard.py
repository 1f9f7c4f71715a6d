"""Frame sharding across GPUs: frames are independent (Terse.hpp:505, :547), so rank r
encodes a contiguous range of frames and the only cross-rank work is host-side bookkeeping --
concatenate the payload slabs in rank order, concatenate the per-frame sizes, take the max of
prolix_bits.  No collective touches pixel or payload data on the device; torch.distributed (gloo on
CPU, nccl on GPUs) only carries these small host objects."""
import numpy as np


def frame_range(n_frames, rank, world):
    """Contiguous frame range [lo, hi) of `rank` (GPU g gets frames [g*F/G, (g+1)*F/G))."""
    return n_frames * rank // world, n_frames * (rank + 1) // world


def merge_encoded(parts):
    """parts: [(payload uint8[], frame_bytes uint64[], prolix_bits)] in rank order -> one stack."""
    payload = np.concatenate([np.asarray(p[0], np.uint8) for p in parts]) if parts else np.zeros(0, np.uint8)
    frame_bytes = np.concatenate([np.asarray(p[1], np.uint64) for p in parts]) if parts else np.zeros(0, np.uint64)
    prolix_bits = max([int(p[2]) for p in parts], default=0)
    return payload, frame_bytes, prolix_bits


def payload_slab(frame_bytes, lo, hi):
    """Byte range [b0, b1) of frames [lo, hi) inside a stack payload with the given per-frame sizes."""
    ends = np.concatenate([[0], np.cumsum(np.asarray(frame_bytes, np.uint64))]).astype(np.uint64)
    return int(ends[lo]), int(ends[hi])


def encode_sharded(encode_fn, stack, dist=None):
    """Encode this rank's frames with encode_fn(stack[lo:hi]) -> (payload, frame_bytes, prolix_bits) and gather
    every rank's result on every rank (all_gather_object).  dist=None: single process."""
    if dist is None or not dist.is_initialized() or dist.get_world_size() == 1:
        return encode_fn(stack)
    world, rank = dist.get_world_size(), dist.get_rank()
    lo, hi = frame_range(stack.shape[0], rank, world)
    mine = encode_fn(stack[lo:hi]) if hi > lo else (np.zeros(0, np.uint8), np.zeros(0, np.uint64), 0)
    parts = [None] * world
    dist.all_gather_object(parts, mine)
    return merge_encoded(parts)


def decode_sharded(decode_fn, payload, frame_bytes, n_frames, dist=None):
    """Decode this rank's frames: decode_fn(slab, frame_bytes[lo:hi]) -> (hi-lo, N) array; gathers all frames."""
    if dist is None or not dist.is_initialized() or dist.get_world_size() == 1:
        return decode_fn(payload, frame_bytes)
    world, rank = dist.get_world_size(), dist.get_rank()
    lo, hi = frame_range(n_frames, rank, world)
    b0, b1 = payload_slab(frame_bytes, lo, hi)
    mine = decode_fn(payload[b0:b1], frame_bytes[lo:hi]) if hi > lo else None
    parts = [None] * world
    dist.all_gather_object(parts, mine)
    return np.concatenate([p for p in parts if p is not None])
