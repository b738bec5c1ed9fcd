#!/usr/bin/env python
"""One micro-batch of the hot path (for ncu launch lists): prof_step.py [micro]"""
import os, sys
import torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import bench
from vspbfr_b200 import fastpath
micro = int(sys.argv[1]) if len(sys.argv) > 1 else 8
dev = torch.device("cuda", 0)
net, dec = bench.build_models(dev)
low, codes, z = (t.to(dev) for t in bench.synth_inputs(micro, 1))
for _ in range(2):
    fastpath.restore_faces(net, dec, low, codes, [z])
torch.cuda.synchronize()
torch.cuda.profiler.start()
fastpath.restore_faces(net, dec, low, codes, [z])
torch.cuda.synchronize()
torch.cuda.profiler.stop()
print("done")
