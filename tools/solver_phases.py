#!/usr/bin/env python
"""Where a frame pair's solve spends its SM cycles (debug taps): averages over the pairs of one launch."""
import ctypes as C
import os
import sys

import numpy as np

REPO = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, REPO)
from video_stabilizer_b200 import _capi as capi, synth  # noqa: E402
from video_stabilizer_b200.clip import Clip, pairs_for_frames  # noqa: E402
from video_stabilizer_b200.imgproc import Context  # noqa: E402

w, h, n = (int(sys.argv[1]), int(sys.argv[2]), int(sys.argv[3])) if len(sys.argv) > 3 else (1920, 1080, 300)
ctx = Context(0)
frames, _ = synth.make_clip_gpu(ctx, w, h, n, 100, chunk=50)
clip = Clip(w, h, n, debug=True, ctx=ctx)
clip.upload(0, frames)
clip.build_pyramids(0, n)
pairs, keys = pairs_for_frames(0, n)
clip.build_keyframes(keys)
for _ in range(2):
    T, st, it = clip.align(pairs)
cyc = np.zeros((n - 1, 6), np.int64)
for p in range(n - 1):
    capi.check(ctx.handle, ctx.lib.vs_clip_get_solver_cycles(clip.handle, p, capi.ptr(cyc[p:p + 1])), "cycles")
rounds = cyc[:, 5].copy()       # slot 5 counts the partition rounds of the selection (both axes share a round)
cyc[:, 5] = 0
tot = cyc.sum(1)
names = ["warpdiff", "select", "hessian sums", "svd || first gather, gn gathers", "gn reduce+update", "-"]
print("pairs %d, converged %d, mean iterations per level %s" % (n - 1, st.sum(), np.round(it.mean(0), 2)))
print("mean cycles per pair %.0f (max %.0f) = %.3f ms at 1.965 GHz" % (tot.mean(), tot.max(), tot.mean() / 1.965e6))
print("selection: %.1f parallel partition rounds per pair (all levels), %.0f cycles per round" % (rounds.mean(), cyc[:, 1].sum() / max(rounds.sum(), 1)))
for i, nm in enumerate(names[:5]):
    print("  %-32s %9.0f cycles  %5.1f%%" % (nm, cyc[:, i].mean(), 100 * cyc[:, i].sum() / tot.sum()))
# per level: {warp-diff, selection, Hessian sums, SVD beside the first iteration, reduce + update, rounds, later gathers, serial steps}
L = clip.levels if hasattr(clip, "levels") else len(it[0])
lv = np.zeros((n - 1, L, 8), np.int64)
for p in range(n - 1):
    capi.check(ctx.handle, ctx.lib.vs_clip_get_solver_level_cycles(clip.handle, p, capi.ptr(lv[p])), "level cycles")
m = lv.mean(0)
print("level  iters  warpdiff   select (rounds, serial steps)   hessian  svd||iter1  later gathers  reduce+update")
for l in range(L - 1, -1, -1):
    print("  L%d   %5.2f  %8.0f  %8.0f (%4.1f, %7.0f)        %7.0f    %7.0f       %7.0f        %7.0f"
          % (l, it[:, l].mean(), m[l, 0], m[l, 1], m[l, 5], m[l, 7], m[l, 2], m[l, 3], m[l, 6], m[l, 4]))
