#!/usr/bin/env python
"""NV12 video through the partitioned pipeline (bench.py's measure_partitioned, nv12=True): device-resident value and the
host-streamed end-to-end rate for several sub-chunk sizes / videos in flight.  One JSON line per setting."""
import argparse
import json
import os
import sys
import types

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--width", type=int, default=1920)
    ap.add_argument("--height", type=int, default=1080)
    ap.add_argument("--frames", type=int, default=300)
    ap.add_argument("--subs", default="16,32,48")
    ap.add_argument("--inflight", default="2")
    ap.add_argument("--bgr", action="store_true", help="the BGR pipeline with the same settings, for comparison")
    a = ap.parse_args()
    import bench
    R = bench.Ranks()
    ceiling = bench.host_copy_ceiling(R)
    ceil_sum = ceiling["h2d_duplex_gbs"] + ceiling["d2h_duplex_gbs"]
    for infl in [int(x) for x in a.inflight.split(",")]:
        for sub in [int(x) for x in a.subs.split(",")]:
            args = types.SimpleNamespace(crop=0, inflight=infl, e2e_inflight=infl, width=a.width, height=a.height, frames=a.frames)
            r = bench.measure_partitioned(R, args, a.width, a.height, a.frames, 2, 1, 4, 4, sub, False, "t%d_%d" % (infl, sub),
                                          nv12=not a.bgr)
            e = r["e2e"]
            print(json.dumps({"format": "bgr" if a.bgr else "nv12", "size": "%dx%d" % (a.width, a.height), "inflight": infl,
                              "e2e_sub": sub, "resident_fps": r["value"], "e2e_fps": e["value"], "h2d_gbs": e["h2d_gbs"],
                              "d2h_gbs": e["d2h_gbs"], "frac_of_host_ceiling": (e["h2d_gbs"] + e["d2h_gbs"]) / ceil_sum}), flush=True)
    R.close()


if __name__ == "__main__":
    main()
