"""ctypes binding of libomnissm.so (include/omnissm.h).

The structures here mirror the header field-for-field.  There is NO fallback: if the shared
library is missing (or a tensor is not on a CUDA device) the call raises - the product path
never routes through the CPU oracle.
"""
from __future__ import annotations

import ctypes as C
import os
from typing import Optional

import torch

OMNI_MAX_DIMS = 6
OMNI_F32, OMNI_F16, OMNI_BF16, OMNI_I32, OMNI_I64, OMNI_U8 = 0, 1, 2, 3, 4, 5
ACT_NONE, ACT_SILU = 0, 1
SSD_AUTO, SSD_RECURRENT, SSD_CHUNKED_TC = 0, 1, 2

_DTYPES = {
    torch.float32: OMNI_F32, torch.float16: OMNI_F16, torch.bfloat16: OMNI_BF16,
    torch.int32: OMNI_I32, torch.int64: OMNI_I64, torch.uint8: OMNI_U8,
}

_STATUS = {1: "BAD_SHAPE", 2: "BAD_DTYPE", 3: "BAD_STRIDE", 4: "UNSUPPORTED", 5: "CUDA_ERROR"}


class Tensor(C.Structure):
    _fields_ = [("data", C.c_void_p), ("dtype", C.c_int32), ("ndim", C.c_int32),
                ("shape", C.c_int64 * OMNI_MAX_DIMS), ("stride", C.c_int64 * OMNI_MAX_DIMS)]


def _T(*names):
    return [(n, Tensor) for n in names]


class Conv1dFwd(C.Structure):
    _fields_ = _T("x", "weight", "bias", "seq_idx", "initial_states", "out", "final_states") + [("activation", C.c_int32)]


class Conv1dBwd(C.Structure):
    _fields_ = _T("x", "weight", "bias", "dout", "seq_idx", "initial_states", "dx", "dweight", "dbias",
                  "dinitial_states") + [("activation", C.c_int32)]


class Conv1dUpdate(C.Structure):
    _fields_ = _T("x", "conv_state", "weight", "bias", "cache_seqlens", "out") + [("activation", C.c_int32)]


class SsdFwd(C.Structure):
    _fields_ = _T("x", "dt", "A", "B", "C", "D", "z", "dt_bias", "initial_states", "seq_idx", "out", "final_states",
                  "workspace", "chunk_states") + [
        ("chunk_size", C.c_int32), ("dt_softplus", C.c_int32), ("dt_min", C.c_float), ("dt_max", C.c_float),
        ("algo", C.c_int32)]


class SsdBwd(C.Structure):
    _fields_ = _T("x", "dt", "A", "B", "C", "D", "z", "dt_bias", "initial_states", "seq_idx", "out", "dout", "dfinal_states",
                  "dx", "ddt", "dB", "dC", "dz", "dinitial_states", "dA_part", "ddt_bias_part", "dD_part",
                  "workspace", "chunk_states") + [
        ("chunk_size", C.c_int32), ("dt_softplus", C.c_int32), ("dt_min", C.c_float), ("dt_max", C.c_float),
        ("algo", C.c_int32)]


class NormGatedFwd(C.Structure):
    _fields_ = _T("x", "weight", "bias", "z", "out", "rstd", "mean") + [
        ("eps", C.c_float), ("group_size", C.c_int32), ("norm_before_gate", C.c_int32), ("is_rms_norm", C.c_int32)]


class NormGatedBwd(C.Structure):
    _fields_ = _T("x", "weight", "bias", "z", "dout", "rstd", "mean", "dx", "dz", "dw_part", "db_part",
                  "out_recompute") + [
        ("eps", C.c_float), ("group_size", C.c_int32), ("norm_before_gate", C.c_int32), ("is_rms_norm", C.c_int32)]


class AddNormFwd(C.Structure):
    _fields_ = _T("x", "residual", "weight", "bias", "y", "residual_out", "rstd", "mean") + [
        ("eps", C.c_float), ("is_rms_norm", C.c_int32)]


class AddNormBwd(C.Structure):
    _fields_ = _T("xres", "weight", "bias", "dy", "dresidual_in", "rstd", "mean", "dx", "dresidual", "dw_part",
                  "db_part") + [("eps", C.c_float), ("is_rms_norm", C.c_int32)]


class Ssu(C.Structure):
    _fields_ = _T("state", "x", "dt", "A", "B", "C", "D", "z", "dt_bias", "out") + [("dt_softplus", C.c_int32)]


class SelScanFwd(C.Structure):
    _fields_ = _T("u", "delta", "A", "B", "C", "D", "z", "delta_bias", "out", "last_state") + [
        ("delta_softplus", C.c_int32)]


class SelScanBwd(C.Structure):
    _fields_ = _T("u", "delta", "A", "B", "C", "D", "z", "delta_bias", "dout", "du", "ddelta", "dB", "dC", "dz",
                  "dA_part", "dD_part", "ddelta_bias_part", "workspace") + [("delta_softplus", C.c_int32)]


class Gemm(C.Structure):
    _fields_ = _T("a", "b", "a2", "b2", "out")


class SplitConv1dScanFwd(C.Structure):
    _fields_ = _T("zxbcdt", "conv1d_weight", "conv1d_bias", "dt_bias", "A", "D", "initial_states", "seq_idx", "rmsnorm_weight",
                  "outproj_weight", "xbc_conv", "scan_out", "rstd", "y", "out", "final_states", "workspace", "chunk_states") + [
        ("nheads", C.c_int32), ("headdim", C.c_int32), ("ngroups", C.c_int32), ("dstate", C.c_int32), ("chunk_size", C.c_int32),
        ("activation", C.c_int32), ("norm_before_gate", C.c_int32), ("algo", C.c_int32),
        ("dt_min", C.c_float), ("dt_max", C.c_float), ("rmsnorm_eps", C.c_float)]


class DecodeCore(C.Structure):
    _fields_ = _T("zxbcdt", "conv_state", "conv_weight", "conv_bias", "ssm_state", "A", "D", "dt_bias", "norm_weight", "out") + [
        ("eps", C.c_float)]


class SoftmaxCe(C.Structure):
    _fields_ = _T("logits", "labels", "lse", "loss", "scale", "grad") + [("ignore_index", C.c_int64)]


# entry point -> params struct (every symbol include/omnissm.h declares with a params pointer)
ENTRY_POINTS = {
    "omni_causal_conv1d_fwd": Conv1dFwd,
    "omni_causal_conv1d_bwd": Conv1dBwd,
    "omni_causal_conv1d_update": Conv1dUpdate,
    "omni_ssd_chunk_scan_fwd": SsdFwd,
    "omni_ssd_chunk_scan_bwd": SsdBwd,
    "omni_norm_gated_fwd": NormGatedFwd,
    "omni_norm_gated_bwd": NormGatedBwd,
    "omni_add_norm_fwd": AddNormFwd,
    "omni_add_norm_bwd": AddNormBwd,
    "omni_selective_state_update": Ssu,
    "omni_selective_scan_fwd": SelScanFwd,
    "omni_selective_scan_bwd": SelScanBwd,
    "omni_gemm_bf16": Gemm,
    "omni_gemm_f32_decode": Gemm,
    "omni_split_conv1d_scan_fwd": SplitConv1dScanFwd,
    "omni_mamba2_decode_core": DecodeCore,
    "omni_softmax_ce_fwd": SoftmaxCe,
    "omni_softmax_ce_bwd": SoftmaxCe,
}
OTHER_SYMBOLS = ["omni_version", "omni_last_error", "omni_launch_count", "omni_reset_launch_count",
                 "omni_ssd_bwd_workspace_elems", "omni_selective_scan_bwd_workspace_elems", "omni_ssd_bwd_tc_workspace_bytes", "omni_ssd_fwd_workspace_bytes", "omni_selftest", "omni_debug_set_trace", "omni_debug_tmem_bench", "omni_debug_set_mbar_hint", "omni_debug_set_bwd_trace", "omni_debug_set_handoff", "omni_gemm_bf16_supported", "omni_debug_set_gemm_mode", "omni_debug_set_pdl", "omni_ssd_bwd_tc_supported", "omni_ssd_fwd_tc_supported", "omni_ssd_chunk_states_bytes", "omni_ssd_fwd_saves_chunk_states"]

# OMNI_LIB_PATH: A/B experiments only (a second build of the same library, e.g. scripts/ab_build.sh); the product path is
# the in-tree lib/libomnissm.so
LIB_PATH = os.environ.get("OMNI_LIB_PATH") or os.path.join(os.path.dirname(os.path.abspath(__file__)), "lib", "libomnissm.so")
_lib: Optional[C.CDLL] = None


def lib() -> C.CDLL:
    """Load libomnissm.so (once).  Raises if it has not been built - there is no fallback."""
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(LIB_PATH):
        raise RuntimeError(
            f"{LIB_PATH} is missing: build it with `python -m omnimamba_b200.build` (needs nvcc). "
            "omnimamba_b200 has no CPU or PyTorch fallback for its CUDA kernels.")
    l = C.CDLL(LIB_PATH)
    for name, struct in ENTRY_POINTS.items():
        fn = getattr(l, name)
        fn.argtypes = [C.POINTER(struct), C.c_void_p]
        fn.restype = C.c_int
    l.omni_version.restype = C.c_int
    l.omni_last_error.restype = C.c_char_p
    l.omni_launch_count.restype = C.c_int64
    l.omni_reset_launch_count.restype = None
    l.omni_ssd_bwd_workspace_elems.argtypes = [C.c_int64] * 5
    l.omni_ssd_bwd_workspace_elems.restype = C.c_int64
    l.omni_selective_scan_bwd_workspace_elems.argtypes = [C.c_int64] * 4
    l.omni_selective_scan_bwd_workspace_elems.restype = C.c_int64
    l.omni_ssd_bwd_tc_workspace_bytes.argtypes = [C.c_int64] * 6
    l.omni_ssd_bwd_tc_workspace_bytes.restype = C.c_int64
    l.omni_ssd_fwd_workspace_bytes.argtypes = [C.c_int64] * 6
    l.omni_ssd_fwd_workspace_bytes.restype = C.c_int64
    l.omni_selftest.argtypes = [C.c_void_p] * 10 + [C.c_int, C.c_void_p]
    l.omni_selftest.restype = C.c_int
    l.omni_debug_set_trace.argtypes = [C.c_void_p, C.c_int]
    l.omni_debug_set_trace.restype = None
    l.omni_debug_tmem_bench.argtypes = [C.c_void_p, C.c_int, C.c_int, C.c_int, C.c_void_p]
    l.omni_debug_tmem_bench.restype = C.c_int
    l.omni_gemm_bf16_supported.restype = C.c_int
    l.omni_ssd_bwd_tc_supported.argtypes = [C.POINTER(SsdBwd)]
    l.omni_ssd_bwd_tc_supported.restype = C.c_int
    l.omni_ssd_fwd_tc_supported.argtypes = [C.POINTER(SsdFwd)]
    l.omni_ssd_fwd_tc_supported.restype = C.c_int
    l.omni_ssd_chunk_states_bytes.argtypes = [C.c_int64] * 5
    l.omni_ssd_chunk_states_bytes.restype = C.c_int64
    l.omni_ssd_fwd_saves_chunk_states.argtypes = [C.POINTER(SsdFwd)]
    l.omni_ssd_fwd_saves_chunk_states.restype = C.c_int
    l.omni_debug_set_gemm_mode.argtypes = [C.c_int]
    l.omni_debug_set_gemm_mode.restype = None
    l.omni_debug_set_pdl.argtypes = [C.c_int]
    l.omni_debug_set_pdl.restype = None
    l.omni_debug_set_handoff.argtypes = [C.c_uint, C.c_int]
    l.omni_debug_set_handoff.restype = None
    _lib = l
    return l


def tdesc(t: Optional[torch.Tensor]) -> Tensor:
    """Describe a CUDA tensor (or None -> absent) for the C ABI.  (Hot on the host side: at the 72-token prefill of
    inference_t2i.py the kernels take 3-16 us and this wrapper layer used to take 35-70 us per call - scripts/
    bench_host_overhead.py - so shapes and strides are assigned as slices and nothing is looked up twice.)"""
    d = Tensor()
    if t is None:
        return d
    if not t.is_cuda:
        raise RuntimeError("omnimamba_b200 kernels need CUDA tensors (no CPU fallback); got a tensor on " + str(t.device))
    dt = _DTYPES.get(t.dtype)
    if dt is None:
        raise TypeError(f"unsupported dtype {t.dtype}")
    n = t.dim()
    if n > OMNI_MAX_DIMS:
        raise ValueError("too many dimensions")
    # keep "present" semantics for empty tensors: give them a non-null dummy address
    d.data = t.data_ptr() if t.numel() > 0 else 1 << 4
    d.dtype = dt
    d.ndim = n
    if n:
        d.shape[0:n] = t.shape
        d.stride[0:n] = t.stride()
    return d


_FN_CACHE = {}
try:
    _raw_stream = torch._C._cuda_getCurrentRawStream   # (the accessor Triton's launcher uses: no Stream object is built)
except AttributeError:  # pragma: no cover
    _raw_stream = None


def call(name: str, params: C.Structure, device: torch.device) -> None:
    """Invoke an entry point on the current stream of `device`; raise on a non-zero status."""
    fn = _FN_CACHE.get(name)
    if fn is None:
        fn = _FN_CACHE[name] = getattr(lib(), name)
    idx = device.index
    cur = torch.cuda.current_device()
    if (idx is None or idx == cur) and _raw_stream is not None:
        # (the usual case - the tensors live on the current device: no device-guard context manager, ~10 us of Python)
        rc = fn(C.byref(params), C.c_void_p(_raw_stream(cur)))
    else:
        with torch.cuda.device(device):
            stream = torch.cuda.current_stream(device).cuda_stream
            rc = fn(C.byref(params), C.c_void_p(stream))
    if rc != 0:
        l = lib()
        msg = l.omni_last_error().decode("utf-8", "replace")
        kind = _STATUS.get(rc, str(rc))
        if rc in (1, 2, 3):
            raise ValueError(f"{name}: {kind}: {msg}")
        if rc == 4:
            raise NotImplementedError(f"{name}: {msg}")
        raise RuntimeError(f"{name}: {kind}: {msg}")


def launch_count() -> int:
    return int(lib().omni_launch_count())


def reset_launch_count() -> None:
    lib().omni_reset_launch_count()


def ssd_chunk_states_bytes(batch, seqlen, nheads, headdim, dstate) -> int:
    return int(lib().omni_ssd_chunk_states_bytes(batch, seqlen, nheads, headdim, dstate))


def ssd_fwd_workspace_bytes(batch, seqlen, nheads, headdim, ngroups, dstate) -> int:
    return int(lib().omni_ssd_fwd_workspace_bytes(batch, seqlen, nheads, headdim, ngroups, dstate))


def ssd_bwd_tc_workspace_bytes(batch, seqlen, nheads, headdim, ngroups, dstate) -> int:
    return int(lib().omni_ssd_bwd_tc_workspace_bytes(batch, seqlen, nheads, headdim, ngroups, dstate))


def ssd_bwd_workspace_elems(batch, seqlen, nheads, headdim, dstate) -> int:
    return int(lib().omni_ssd_bwd_workspace_elems(batch, seqlen, nheads, headdim, dstate))


def selscan_bwd_workspace_elems(batch, dim, seqlen, dstate) -> int:
    return int(lib().omni_selective_scan_bwd_workspace_elems(batch, dim, seqlen, dstate))


def gemm_supported() -> bool:
    return bool(lib().omni_gemm_bf16_supported())


def _gemm_operand_ok(t) -> bool:
    if t.dim() != 2 or t.dtype != torch.bfloat16 or t.data_ptr() % 16:
        return False
    r, c = t.shape
    if t.stride(1) == 1 and (r == 1 or (t.stride(0) % 8 == 0 and t.stride(0) >= c)):
        return True
    return t.stride(0) == 1 and (c == 1 or (t.stride(1) % 8 == 0 and t.stride(1) >= r))


def gemm_operands_ok(a, b, a2=None, b2=None) -> bool:
    """True when the tcgen05 GEMM takes these operands in place (bf16, one contiguous dim, 16-byte aligned rows)."""
    ok = _gemm_operand_ok(a) and _gemm_operand_ok(b) and a.shape[1] == b.shape[1]
    if a2 is not None:
        ok = ok and _gemm_operand_ok(a2) and _gemm_operand_ok(b2) and (a2.stride(1) == 1) == (a.stride(1) == 1) and \
            (b2.stride(1) == 1) == (b.stride(1) == 1)
    return ok


def gemm(a, b, out_dtype=torch.bfloat16, a2=None, b2=None, out=None):
    """out (M, N) = a (M, K) @ b (N, K)^T [+ a2 @ b2^T] on the tcgen05 GEMM of libomnissm."""
    M, N = a.shape[0], b.shape[0]
    if out is None:
        out = torch.empty(M, N, device=a.device, dtype=out_dtype)
    p = Gemm()
    p.a, p.b, p.a2, p.b2, p.out = tdesc(a), tdesc(b), tdesc(a2), tdesc(b2), tdesc(out)
    call("omni_gemm_bf16", p, a.device)
    return out


def gemm_f32_decode_ok(a, b, a2=None, b2=None) -> bool:
    """True when omni_gemm_f32_decode (3xTF32 weight-streaming kernel) takes these fp32 operands: decode shapes only."""
    def row_major(t):
        r, c = t.shape
        return t.dim() == 2 and t.dtype == torch.float32 and t.data_ptr() % 16 == 0 and (c <= 1 or t.stride(1) == 1) and \
            (r <= 1 or (t.stride(0) % 4 == 0 and t.stride(0) >= c))
    if a.dim() != 2 or b.dim() != 2 or not (row_major(a) and row_major(b)) or a.shape[1] != b.shape[1]:
        return False
    M, K = a.shape
    N = b.shape[0]
    if not (1 <= M <= 128 and N >= 256 and N % 4 == 0 and K >= 128 and K % 4 == 0):
        return False
    if a2 is not None:
        if not (row_major(a2) and row_major(b2)) or a2.shape[1] % 4 or a2.shape[0] != M or b2.shape[0] != N:
            return False
    return gemm_supported()


def gemm_f32_decode(a, b, a2=None, b2=None, out=None):
    """out (M, N) = a (M, K) @ b (N, K)^T [+ a2 @ b2^T], fp32 operands and result, on the 3xTF32 weight-streaming kernel."""
    M, N = a.shape[0], b.shape[0]
    if out is None:
        out = torch.empty(M, N, device=a.device, dtype=torch.float32)
    p = Gemm()
    p.a, p.b, p.a2, p.b2, p.out = tdesc(a), tdesc(b), tdesc(a2), tdesc(b2), tdesc(out)
    call("omni_gemm_f32_decode", p, a.device)
    return out
