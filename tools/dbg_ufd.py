import sys, os, subprocess
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
if len(sys.argv) > 1:
    import torch
    from vspbfr_b200.op.upfirdn2d import upfirdn2d_raw
    k1 = torch.tensor([1.0, 3.0, 3.0, 1.0]); k = (k1[None] * k1[:, None] / 64).cuda()
    n, c, h, w, up, pad = [int(v) for v in sys.argv[1:7]]
    x = torch.randn(n, c, h, w, device="cuda")
    y = upfirdn2d_raw(x, k, (up, up), (1, 1), (pad, pad, pad, pad))
    torch.cuda.synchronize()
    os.environ["VSP_NO_TMA"] = "1"
    print("ok", tuple(y.shape), float(y.abs().sum()))
else:
    for args in ["2 3 8 8 2 2", "1 1 64 64 1 2", "1 1 256 256 1 2", "1 8 256 256 1 2", "1 1 512 512 1 2", "4 32 16 16 1 2", "64 1 16 16 1 2", "64 1 32 32 1 2"]:
        r = subprocess.run([sys.executable, __file__] + args.split(), capture_output=True, text=True)
        print(args, "->", (r.stdout.strip() or r.stderr.strip().splitlines()[-1][:150]))
