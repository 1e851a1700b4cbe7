"""The UNMODIFIED reference model code imports and builds against the drop-in mamba_ssm / causal_conv1d packages.

Runs only where /root/reference exists (the build container; it has no GPU, so this is import + construction + the
"fails loudly without CUDA" check - the numerics of the same modules are covered on the GPU by tests/test_gpu_parity.py).
Stub order follows SURVEY.md 8(c): transformers first, then the accelerate stub."""
import os
import sys
import types

import pytest
import torch

REF = "/root/reference"
pytestmark = pytest.mark.skipif(not os.path.isdir(os.path.join(REF, "models", "stage2")), reason="reference tree not present")


@pytest.fixture(scope="module")
def ref_modules():
    import omnimamba_b200
    omnimamba_b200.install_dropin()
    import transformers.generation as tg
    import transformers.integrations  # noqa: F401
    for name in ("GreedySearchDecoderOnlyOutput", "SampleDecoderOnlyOutput"):  # removed in transformers 5 (reference pins 4.46.1)
        if not hasattr(tg, name):
            setattr(tg, name, tg.GenerateDecoderOnlyOutput)
    if "accelerate" not in sys.modules:  # not installed here; the reference only needs add_hook_to_module to import
        acc, hooks = types.ModuleType("accelerate"), types.ModuleType("accelerate.hooks")
        hooks.add_hook_to_module = lambda module, hook, append=False: module
        acc.hooks = hooks
        sys.modules["accelerate"], sys.modules["accelerate.hooks"] = acc, hooks
    sys.path.insert(0, REF)
    try:
        from models.stage2 import block, mixer_seq_simple
        yield mixer_seq_simple, block
    finally:
        sys.path.remove(REF)


def test_reference_imports_resolve_to_the_dropin(ref_modules):
    mixer_seq_simple, block = ref_modules
    import mamba_ssm
    import causal_conv1d
    here = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "omnimamba_b200", "dropin")
    assert os.path.abspath(mamba_ssm.__file__).startswith(here)
    assert os.path.abspath(causal_conv1d.__file__).startswith(here)
    from omnimamba_b200.modules import Mamba2
    assert mixer_seq_simple.Mamba2 is Mamba2
    assert block.RMSNorm is mixer_seq_simple.RMSNorm


def test_reference_backbone_builds_on_the_shim(ref_modules):
    mixer_seq_simple, _ = ref_modules
    from omnimamba_b200.modules import Mamba2
    blk = mixer_seq_simple.create_block(256, d_intermediate=0, ssm_cfg={"layer": "Mamba2"}, rms_norm=True, residual_in_fp32=True,
                                        fused_add_norm=True, layer_idx=0)
    assert isinstance(blk.mixer, Mamba2)
    sd = blk.state_dict()
    # parameter contract of SURVEY.md Appendix C at d_model=256
    assert sd["mixer.in_proj.weight"].shape == (1288, 256)
    assert sd["mixer.conv1d.weight"].shape == (768, 1, 4)
    assert sd["mixer.out_proj.weight"].shape == (256, 512)
    assert sd["mixer.norm.weight"].shape == (512,)
    for k in ("mixer.dt_bias", "mixer.A_log", "mixer.D"):
        assert sd[k].shape == (8,)
    # no CPU fallback: the product path must fail loudly without a CUDA device
    if not torch.cuda.is_available():
        with pytest.raises(RuntimeError):
            blk.mixer(torch.randn(1, 8, 256))
