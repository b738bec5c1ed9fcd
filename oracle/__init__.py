"""CPU oracle for the VSPBFR synthesis hot path — TEST INFRASTRUCTURE ONLY.

A restatement (numpy / torch-CPU, fp32 or fp64) of the reference's algorithms for
upfirdn2d, fused bias+leaky-ReLU and the modulated convolution.  Only ``tests/``,
``__graft_entry__.smoke()`` and ``bench.py``'s cpu_baseline / ``--impl reference``
legs may import this package; the product path (``vspbfr_b200``) never does and
fails loudly when its CUDA library is missing.

``oracle/c/vsp_oracle.c`` restates the two memory-bound operators in plain C (built by
``oracle/build_c.py`` with gcc; ``build_c.upfirdn2d_c`` / ``build_c.bias_act_c``).

Parity pinning: the reference ships NO tests or golden vectors, so
the oracle is pinned against outputs of the reference's own Python/CPU code run in
the build container (``tests/golden/make_golden.py`` -> ``tests/golden/*.npz``);
``tests/test_oracle_golden.py`` checks every function here against them.
"""
from .upfirdn2d_ref import upfirdn2d_ref, upfirdn2d_native_port, upfirdn2d_out_size, upfirdn2d_grad_pads
from .bias_act_ref import bias_act_ref, fused_leaky_relu_ref, fused_leaky_relu_grads_ref
from .modconv_ref import modulated_conv2d_ref, conv2d_ref

__all__ = [
    "upfirdn2d_ref", "upfirdn2d_native_port", "upfirdn2d_out_size", "upfirdn2d_grad_pads",
    "bias_act_ref", "fused_leaky_relu_ref", "fused_leaky_relu_grads_ref",
    "modulated_conv2d_ref", "conv2d_ref",
]
from .network_ref import generator_ref, restoration_ref, restore_faces_ref  # noqa: E402

__all__ += ["generator_ref", "restoration_ref", "restore_faces_ref"]
from .discriminator_ref import discriminator_ref, r1_step_ref  # noqa: E402

__all__ += ["discriminator_ref", "r1_step_ref"]
