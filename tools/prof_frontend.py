"""One micro-batch of the PyTorch W+ front end (for ncu launch lists): prof_frontend.py [micro]"""
import os, sys
import torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from vspbfr_b200 import frontend
micro = int(sys.argv[1]) if len(sys.argv) > 1 else 32
dev = torch.device("cuda", 0)
torch.manual_seed(1)
front = frontend.WPlusFrontEnd(frontend.Encoder4Editing(50, "ir_se"), n_latent=18).to(dev).eval().half_precision_()
ddpm = frontend.My_DDPM(frontend.Code_diffuser(timesteps=4), timesteps=4, linear_start=0.1, linear_end=0.99).to(dev).eval()
low = torch.rand(micro, 3, 512, 512, device=dev) * 2 - 1
for _ in range(2):
    ddpm(condi_in=front(low), tf32=True)
torch.cuda.synchronize()
torch.cuda.profiler.start()
ddpm(condi_in=front(low), tf32=True)
torch.cuda.synchronize()
torch.cuda.profiler.stop()
print("done")
