"""Gather microbenchmarks under a few load policies (see mp_bench.cu load16); used under ncu to read sector amplification."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import megapath_b200 as mp
c = mp.Context(0)
for k in ([int(x) for x in sys.argv[1:]] or (0, 11, 1)):
    print(k, round(c.microbench(k), 1), flush=True)
