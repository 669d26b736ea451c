"""A/B aid: single-stream and 8-stream frames/s of the C2 workload for whatever library / env the process was started with."""
import os, sys, json
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import bench
wl = bench.WORKLOADS["C2"]
ring = bench.Ring(0, wl, torch)
def barrier(): torch.cuda.synchronize()
ident = lambda v: v
r1 = bench.fe_line(wl, [ring], 0, 1, 10, 3, barrier, None, ident)
r8 = bench.fe_line(wl, [ring], 0, 8, 6, 2, barrier, None, ident)
r8t = bench.fe_line(wl, [ring], 0, 8, 6, 2, barrier, None, ident, threads=True)
print(sys.argv[1], "single %d / %d   8 streams one thread %d / %d   8 threads %d / %d" % (r1["value"], r1["e2e"], r8["value"], r8["e2e"], r8t["value"], r8t["e2e"]))
