"""Frame-chunk partitioning of one long video across ranks.

Frame pairs are independent, so a video splits into contiguous chunks, one per GPU, with no
collective on the data path.  What must be preserved is the reference's keyframe alternation:
odd frames (0-based, counted from the start of the video) are keyframes
(alignment.cpp:357,396-397), so chunks start on even frame indices and every chunk but the
first re-reads one halo frame — the last frame of the previous chunk — to align its first
frame against.  Only the 40 bytes per pair (transform + status) are exchanged: they are
all-gathered so that the sequential trajectory (L1 smoother, accumulate, decay) can run over
the whole video; that exchange is the only communication and it is host-side plumbing.
"""
from __future__ import annotations

import numpy as np


def frame_chunks(n_frames: int, world: int):
    """[(first, last_exclusive)] per rank: contiguous, even-aligned starts, sizes differ by <= 2."""
    bounds = [0]
    for r in range(1, world):
        b = (n_frames * r // world) & ~1           # even boundary: frame parity is global
        bounds.append(max(b, bounds[-1]))
    bounds.append(n_frames)
    return [(bounds[r], bounds[r + 1]) for r in range(world)]


def chunk_upload_range(chunk, rank: int):
    """Frames a rank must hold: its chunk plus the halo frame before it (ranks > 0)."""
    first, last = chunk
    return (first - 1 if rank > 0 and first > 0 and last > first else first), last


def chunk_pairs(chunk, rank: int, slot_of=None):
    """Alignment jobs (template_slot, keyframe_slot, invert) and keyframe slots of one chunk.
    Frame f of the video sits in slot f - upload_first unless slot_of says otherwise."""
    first, last = chunk
    up_first, _ = chunk_upload_range(chunk, rank)
    slot_of = slot_of or (lambda f: f - up_first)
    pairs, keyframes = [], set()
    for f in range(max(first, 1), last):
        if f % 2 == 1:
            pairs.append((slot_of(f - 1), slot_of(f), 0))
            keyframes.add(slot_of(f))
        else:
            pairs.append((slot_of(f), slot_of(f - 1), 1))
            keyframes.add(slot_of(f - 1))
    return pairs, sorted(keyframes)


def gather_measurements(local: np.ndarray, chunk, n_frames: int, group=None) -> np.ndarray:
    """All-gather the per-frame records (n_local, 5) = {A, B, TX, TY, ok} of every rank into the
    (n_frames, 5) table of the whole video.  `local` covers frames chunk[0]..chunk[1]-1
    (frame 0's record is the identity / not-ok of a first frame).  Uses torch.distributed when
    it is initialised (gloo or nccl), otherwise returns the local table."""
    import torch
    import torch.distributed as dist
    table = np.zeros((n_frames, 5), np.float64)
    if not (dist.is_available() and dist.is_initialized()):
        table[chunk[0]:chunk[1]] = local
        return table
    world = dist.get_world_size(group)
    sizes = [b - a for (a, b) in frame_chunks(n_frames, world)]
    width = max(sizes)
    pad = np.zeros((width, 5), np.float64)
    pad[: local.shape[0]] = local
    device = "cuda" if dist.get_backend(group) == "nccl" else "cpu"
    mine = torch.from_numpy(pad).to(device)
    out = [torch.empty_like(mine) for _ in range(world)]
    dist.all_gather(out, mine, group=group)
    pos = 0
    for r in range(world):
        table[pos: pos + sizes[r]] = out[r].cpu().numpy()[: sizes[r]]
        pos += sizes[r]
    return table


def corrections_for_video(table: np.ndarray, width: int, height: int, params=None):
    """Run the sequential host trajectory over the gathered measurements.  Returns
    (corrections (n_out, 4), first_output_frame = 0): correction i belongs to frame i."""
    from . import host
    traj = host.StabilizerTrajectory(params)
    out = []
    for row in table:
        due, corr = traj.push(row[:4], bool(row[4]), width, height)
        if due:
            out.append(corr)
    return np.array(out, np.float64).reshape(-1, 4)
