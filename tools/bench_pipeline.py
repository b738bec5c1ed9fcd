#!/usr/bin/env python
"""restoration_test.py:125-131 end to end (e4e encoder + code diffuser + hot path) as one CUDA graph per 32-face micro-batch,
and the front end alone (eager + graph): faces/s on one GPU.  VSP_FRONT_OWN_CONVS=0 keeps the encoder's convolutions on cuDNN."""
import json, os, sys
import torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import bench
from vspbfr_b200 import frontend

def main():
    dev = torch.device("cuda", 0)
    micro = int(os.environ.get("MICRO", "32"))
    net, dec = bench.build_models(dev)
    low, codes, z = (t.to(dev) for t in bench.synth_inputs(micro, 1))
    torch.manual_seed(1)
    front = frontend.WPlusFrontEnd(frontend.Encoder4Editing(50, "ir_se"), n_latent=18).to(dev).eval().half_precision_()
    ddpm = frontend.My_DDPM(frontend.Code_diffuser(timesteps=4), timesteps=4, linear_start=0.1, linear_end=0.99).to(dev).eval()
    g = frontend.GraphedPipeline(front, ddpm, dec, net, micro, device=dev)

    def t(fn, n=5):
        fn(); torch.cuda.synchronize()
        s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        s.record()
        for _ in range(n):
            fn()
        e.record(); torch.cuda.synchronize()
        return s.elapsed_time(e) / n
    with torch.no_grad():
        ms_pipe = t(lambda: g(low, z, clone=False))
        ms_front = t(lambda: ddpm(condi_in=front(low), tf32=True))
        fg = torch.cuda.CUDAGraph()
        with torch.cuda.graph(fg):
            out = ddpm(condi_in=front(low), tf32=True)
        ms_front_g = t(fg.replay)
    row = {"own_convs": os.environ.get("VSP_FRONT_OWN_CONVS", "1"), "pipeline_ms": ms_pipe, "pipeline_faces_per_s": micro / ms_pipe * 1e3,
           "front_end_eager_ms": ms_front, "front_end_graph_ms": ms_front_g, "launches_per_micro_batch": g.launches}
    print(json.dumps(row))

if __name__ == "__main__":
    main()
