"""W+ front end on the GPU against the unmodified reference's outputs (tests/golden/frontend.npz, generated on the reference's
CPU fp32 path by tests/golden/make_golden_frontend.py): the e4e IR-SE50 encoder in every form the pipeline runs it —
fp32 module, folded (conv+BatchNorm) fp32, and the inference form (bf16 channels-last, fused SE-tail kernel) — and the code
diffuser + reverse-diffusion sampler with TF32 matmuls as bench.py runs them."""
import math

import numpy as np
import pytest
import torch

from conftest import load_golden
from test_frontend_cpu import _build_encoder
from vspbfr_b200 import frontend as fe

pytestmark = pytest.mark.gpu
G = load_golden("frontend")
DEV = "cuda"


def _psnr(got, want):
    peak = float(want.max() - want.min())
    return 10 * math.log10(peak * peak / max(float(((got - want) ** 2).mean()), 1e-30))


def test_encoder_fp32_and_folded_match_reference_on_gpu():
    enc = _build_encoder().to(DEV)
    x = torch.from_numpy(G["enc_x"]).to(DEV)
    want = torch.from_numpy(G["enc_w"])
    prev = torch.backends.cudnn.allow_tf32, torch.backends.cuda.matmul.allow_tf32
    torch.backends.cudnn.allow_tf32 = torch.backends.cuda.matmul.allow_tf32 = False
    try:
        with torch.no_grad():
            got = enc(x).cpu()
            fe.fold_for_inference_(enc)
            folded = enc(x).cpu()
    finally:
        torch.backends.cudnn.allow_tf32, torch.backends.cuda.matmul.allow_tf32 = prev
    torch.testing.assert_close(got, want, rtol=1e-3, atol=1e-4 * float(want.abs().max()))
    torch.testing.assert_close(folded, want, rtol=1e-3, atol=1e-4 * float(want.abs().max()))


def test_encoder_inference_form_bf16_matches_reference():
    """The form bench.py / GraphedPipeline run: folded, bf16 channels-last, SE scale + shortcut + next BatchNorm in
    vsp_se_tail_nhwc_bf16.  Tolerance = north_star's for bf16 paths: max-abs <= 1e-2 of the range, PSNR > 45 dB."""
    from vspbfr_b200 import _lib
    enc = _build_encoder()
    front = fe.WPlusFrontEnd(enc, n_latent=18).to(DEV).eval().half_precision_()
    x = torch.from_numpy(G["enc_x"]).to(DEV)
    want = torch.from_numpy(G["enc_w"])
    n0 = _lib.launch_count()
    with torch.no_grad():
        got = front.encoder(x.to(torch.bfloat16).contiguous(memory_format=torch.channels_last)).float().cpu()
    assert _lib.launch_count() - n0 >= 24                     # one fused tail per IR-SE unit ran
    peak = float(want.max() - want.min())
    err = float((got - want).abs().max())
    print(f"e4e encoder bf16 inference form: max-abs {err / peak:.3e} of range, psnr {_psnr(got, want):.1f} dB")
    assert err <= 1e-2 * peak and _psnr(got, want) > 45.0, (err / peak, _psnr(got, want))


def test_code_diffuser_matches_reference_fp32_and_tf32():
    """The denoiser (CodeDiffuser.py:127-146) on the GPU against the reference's CPU output, in fp32 and with the TF32 matmuls
    bench.py enables; the 4-step sampler draws its own noise (a CUDA generator cannot replay the CPU golden's stream), so it is
    compared with itself across the two precisions under one seed."""
    torch.manual_seed(79)
    den = fe.Code_diffuser(timesteps=4).eval()
    ddpm = fe.My_DDPM(denoise=den, timesteps=4, linear_start=0.1, linear_end=0.99).eval().to(DEV)
    cond, x_t, t = (torch.from_numpy(G[k]).to(DEV) for k in ("den_cond", "den_xt", "den_t"))
    want = torch.from_numpy(G["den_out"])
    prev = torch.backends.cuda.matmul.allow_tf32
    try:
        with torch.no_grad():
            torch.backends.cuda.matmul.allow_tf32 = False
            exact = ddpm.model(x_t, cond, t).cpu()
            torch.backends.cuda.matmul.allow_tf32 = True
            fast = ddpm.model(x_t, cond, t).cpu()
    finally:
        torch.backends.cuda.matmul.allow_tf32 = prev
    peak = float(want.max() - want.min())
    print(f"denoiser vs reference: fp32 max-abs {float((exact - want).abs().max()) / peak:.2e} of range, "
          f"tf32 {float((fast - want).abs().max()) / peak:.2e}, psnr {_psnr(fast, want):.1f} dB")
    assert float((exact - want).abs().max()) <= 1e-3 * peak
    assert float((fast - want).abs().max()) <= 1e-2 * peak and _psnr(fast, want) > 45.0
    # the 4-step sampler from the golden's x_T (deterministic: the reference's p_sample never adds its noise): fp32 on the
    # GPU vs the reference's CPU run.  (With TF32 the random-init denoiser's 4-step map is not contractive enough for a
    # point-wise bound — per-step parity is asserted above.)
    with torch.no_grad():
        sampled = ddpm(condi_in=cond, x_T=x_t).cpu()
    ws = torch.from_numpy(G["ddpm_out"])
    assert float((sampled - ws).abs().max()) <= 2e-3 * float(ws.max() - ws.min())
