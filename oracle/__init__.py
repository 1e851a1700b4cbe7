"""CPU oracle for the Mamba-2 selective-scan hot path.

TEST INFRASTRUCTURE ONLY.  Nothing under ``omnimamba_b200/`` may import this
package: only ``tests/``, ``__graft_entry__.smoke()`` and ``bench.py``'s
``cpu_baseline`` / ``--impl reference`` legs do, and there only as the checker
(or the CPU arm being timed), never as the product path.

PARITY UNPINNED BY THE REFERENCE: hustvl/OmniMamba ships no tests, fixtures or
golden vectors, and the arithmetic of this path lives in two un-vendored pip
pins (``mamba_ssm==2.2.2``, ``causal-conv1d==1.4.0``;
/root/reference/requirements.txt:12-13) that are not installable here.  The
oracle restates their published reference functions (``selective_scan_ref``,
``ssd_minimal_discrete``, ``selective_state_update_ref``, ``causal_conv1d_ref``,
``causal_conv1d_update_ref``, ``rms_norm_ref``) and is pinned instead against
(a) an independent in-container restatement, ``transformers`` 5.5
``Mamba2Mixer.torch_forward`` (tests/golden/make_golden.py), (b) its own
recurrent == chunked == single-step identities, and (c) on the GPU box, vllm's
Triton port of the mamba_ssm v2.2.4 kernels.
"""
from .ssm_oracle import *  # noqa: F401,F403
