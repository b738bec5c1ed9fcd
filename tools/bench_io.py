#!/usr/bin/env python
"""Output side of restoration_test.py:133-157 on 512^2 images: the reference's per-image torchvision.utils.save_image
(+ torch.cuda.empty_cache() per batch) vs vspbfr_b200.imageio.ImageWriter (device-side quantisation, one uint8 D2H copy per
batch, PNG encoding on host threads), with the GPU busy on the hot path in between.  One JSON row per variant."""
import json, os, shutil, sys, tempfile, time
import torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import bench
from vspbfr_b200 import fastpath
from vspbfr_b200.imageio import ImageWriter


def main():
    dev = torch.device("cuda", 0)
    micro, n_batches = int(os.environ.get("MICRO", "32")), int(os.environ.get("BATCHES", "4"))
    net, dec = bench.build_models(dev)
    low, codes, z = (t.to(dev) for t in bench.synth_inputs(micro, 1))
    graphed = fastpath.GraphedRestorer(net, dec, micro, device=dev)
    rows = []

    def hot():
        return graphed(low, codes, z)[0]

    for _ in range(2):
        hot()
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    for _ in range(n_batches):
        hot()
    torch.cuda.synchronize()
    t_hot = (time.perf_counter() - t0) / n_batches

    out = tempfile.mkdtemp(prefix="vsp_io_")
    try:
        from torchvision import utils
        t0 = time.perf_counter()
        for i in range(n_batches):
            restored = hot()
            torch.cuda.empty_cache()
            for j in range(micro):
                utils.save_image(restored[j], f"{out}/{i * micro + j:06d}_0_ref_restore.png", nrow=1, normalize=True,
                                 value_range=(-1, 1))
        torch.cuda.synchronize()
        t_ref = (time.perf_counter() - t0) / n_batches
        shutil.rmtree(out); os.makedirs(out)
        t0 = time.perf_counter()
        with ImageWriter(out, name="own", workers=int(os.environ.get("WORKERS", "16"))) as wr:
            for i in range(n_batches):
                wr.save(i * micro, restore=hot())
        torch.cuda.synchronize()
        t_own = (time.perf_counter() - t0) / n_batches
    finally:
        shutil.rmtree(out, ignore_errors=True)
    for name, t in (("hot path only (no saving)", t_hot), ("reference loop: empty_cache + save_image per image", t_ref),
                    ("ImageWriter: device quantisation + threaded PNG", t_own)):
        rows.append({"variant": name, "ms_per_batch": 1e3 * t, "faces_per_s": micro / t, "micro": micro,
                     "host_cores": os.cpu_count()})
        print(json.dumps(rows[-1]), flush=True)
    os.makedirs("gpurun_out", exist_ok=True)
    json.dump(rows, open("gpurun_out/bench_io.json", "w"), indent=1)


if __name__ == "__main__":
    main()
