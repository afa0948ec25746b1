"""video_stabilizer_b200 — B200 (sm_100a) implementation of catid/video_stabilizer's
per-frame alignment-and-warp hot path behind the reference's operator interface.

  _capi     ctypes binding of the C ABI (include/vstab.h -> libvstab.so)
  imgproc   Python mirror of imgproc.hpp (single operators, host arrays)
  clip      device-resident batched pipeline (vs_clip_*)
  host      binding of the C++ host layer (VideoAligner / VideoStabilizer drop-ins)
  synth     synthetic jittered clips (BASELINE.json configs)
  build     in-tree nvcc / g++ build

Importing the package does not load CUDA; creating a Context does, and raises when the
library or a GPU is missing (there is no CPU fallback).
"""
from ._capi import (VS_BORDER_CONSTANT0, VS_BORDER_REPEAT_EDGE, VS_MEM_DEVICE, VS_MEM_HOST,  # noqa: F401
                    VS_WARP_CV_EXACT_BILINEAR, VS_WARP_FLOAT_BILINEAR, VS_WARP_LANCZOS2, VsError)

__all__ = ["imgproc", "clip", "synth", "build"]
