"""ctypes binding of libxeq_b200.so (C ABI: include/xeq_b200.h).

There is deliberately no CPU / eager fallback: if the library is missing, or a call is
made with host tensors, the caller gets a loud error."""
from __future__ import annotations

import ctypes
from ctypes import POINTER, c_char_p, c_float, c_int, c_int32, c_size_t, c_void_p
from pathlib import Path

import torch

import os

# XEQ_LIB: load another build of the same C ABI (the test-only SIMT variant, a debug build); default = the product library
LIB_PATH = Path(os.environ.get("XEQ_LIB") or (Path(__file__).resolve().parent / "libxeq_b200.so"))


class XeqDims(ctypes.Structure):
    _fields_ = [("node_dim", c_int32), ("mul0", c_int32), ("mul1", c_int32), ("mul2", c_int32),
                ("num_basis", c_int32), ("cutoff", c_float)]


class XeqGraph(ctypes.Structure):
    _fields_ = [("n_nodes", c_int32), ("n_edges", c_int32), ("n_graphs", c_int32), ("_pad", c_int32),
                ("rowptr", c_void_p), ("col", c_void_p), ("t_rowptr", c_void_p), ("t_row", c_void_p),
                ("t_eid", c_void_p), ("offsets", c_void_p), ("cell", c_void_p), ("node_graph", c_void_p),
                ("tile_ptr", c_void_p), ("t_tile_ptr", c_void_p),
                ("n_tiles", c_int32), ("t_n_tiles", c_int32), ("tile_mode", c_int32), ("max_tile_nodes", c_int32)]


class XeqGemm(ctypes.Structure):
    _fields_ = [("a", c_void_p), ("b", c_void_p), ("bias", c_void_p), ("c", c_void_p),
                ("m", c_int32), ("n", c_int32), ("k", c_int32), ("lda", c_int32), ("ldb", c_int32), ("ldc", c_int32),
                ("a_trans", c_int32), ("b_trans", c_int32), ("alpha", c_float), ("act", c_int32)]


_SIGNATURES = {
    "xeq_version": (c_int, []),
    "xeq_last_error": (c_char_p, []),
    "xeq_num_sms": (c_int, []),
    "xeq_launch_count": (ctypes.c_longlong, []),
    "xeq_radius_graph_workspace_bytes": (c_size_t, [c_int32, c_int32, c_int]),
    "xeq_radius_graph_count": (c_int, [c_void_p, c_int32, c_void_p, c_void_p, c_int32, c_void_p, POINTER(c_int32),
                                       POINTER(c_int32), c_float, c_void_p, c_void_p, c_size_t, c_void_p]),
    "xeq_radius_graph_fill": (c_int, [c_void_p, c_int32, c_void_p, c_void_p, c_int32, c_void_p, POINTER(c_int32),
                                      POINTER(c_int32), c_float, c_void_p, c_void_p, c_void_p, c_void_p, c_void_p,
                                      c_int32, c_void_p, c_void_p, c_size_t, c_void_p]),
    "xeq_csr_from_sorted_coo": (c_int, [c_void_p, c_void_p, c_int32, c_int32, c_void_p, c_void_p, c_void_p, c_void_p]),
    "xeq_csr_transpose_workspace_bytes": (c_size_t, [c_int32, c_int32]),
    "xeq_csr_transpose": (c_int, [c_void_p, c_void_p, c_int32, c_int32, c_void_p, c_void_p, c_void_p, c_void_p,
                                  c_size_t, c_void_p]),
    "xeq_center_tile_edges": (c_int, []),
    "xeq_neighbor_tile_edges": (c_int, []),
    "xeq_csr_tile_count": (c_int, [c_int32, c_int32, c_int32]),
    "xeq_csr_tile_bounds": (c_int, [c_void_p, c_int32, c_int32, c_int32, c_void_p, c_void_p]),
    "xeq_edge_message_fwd_workspace_bytes": (c_size_t, [POINTER(XeqGraph), POINTER(XeqDims)]),
    "xeq_edge_message_fwd": (c_int, [POINTER(XeqGraph), POINTER(XeqDims)] + [c_void_p] * 10 + [c_void_p, c_size_t, c_void_p]),
    "xeq_edge_message_bwd_workspace_bytes": (c_size_t, [POINTER(XeqGraph), POINTER(XeqDims), c_int]),
    "xeq_edge_message_bwd": (c_int, [POINTER(XeqGraph), POINTER(XeqDims)] + [c_void_p] * 14 + [c_void_p, c_size_t, c_void_p]),
    "xeq_edge_message_bwdbwd_workspace_bytes": (c_size_t, [POINTER(XeqGraph), POINTER(XeqDims), c_int]),
    "xeq_edge_message_bwdbwd": (c_int, [POINTER(XeqGraph), POINTER(XeqDims)] + [c_void_p] * 20 + [c_void_p, c_size_t, c_void_p]),
    "xeq_segment_sum": (c_int, [c_void_p, c_void_p, c_int32, c_void_p, c_void_p]),
    "xeq_colsum": (c_int, [c_void_p, c_int32, c_int32, c_int32, c_void_p, c_void_p]),
    "xeq_colsum_weighted": (c_int, [c_void_p, c_void_p, c_int32, c_int32, c_int32, c_void_p, c_void_p]),
    "xeq_rowdot": (c_int, [c_void_p, c_int32, c_void_p, c_void_p, c_int32, c_int32, c_void_p, c_void_p]),
    "xeq_outer": (c_int, [c_void_p, c_void_p, c_int32, c_int32, c_void_p, c_void_p]),
    "xeq_edge_cell_grad_rows": (c_int, [POINTER(XeqGraph), POINTER(XeqDims), c_void_p, c_void_p, c_void_p]),
    "xeq_gemm_workspace_bytes": (c_size_t, [POINTER(XeqGemm), c_int32, c_int32]),
    "xeq_gemm_tf32x3": (c_int, [POINTER(XeqGemm), c_int32, c_int32, c_void_p, c_size_t, c_void_p]),
    "xeq_irreps_norm_workspace_bytes": (c_size_t, [c_int32] * 4),
    "xeq_irreps_norm_fwd": (c_int, [c_void_p] * 3 + [c_int32] * 4 + [c_float, c_void_p, c_void_p]),
    "xeq_irreps_norm_bwd": (c_int, [c_void_p] * 3 + [c_int32, c_void_p] + [c_int32] * 4 + [c_float] + [c_void_p] * 3 + [c_void_p, c_size_t, c_void_p]),
    "xeq_irreps_norm_bwdbwd": (c_int, [c_void_p] * 3 + [c_int32, c_void_p] + [c_int32] * 4 + [c_float] + [c_void_p] * 3 + [c_void_p, c_size_t, c_void_p]),
    "xeq_invariant_dot_fwd": (c_int, [c_void_p] * 2 + [c_int32] * 4 + [c_void_p, c_int32, c_void_p, c_void_p]),
    "xeq_invariant_dot_bwd": (c_int, [c_void_p] * 3 + [c_int32, c_void_p, c_void_p] + [c_int32] * 4 + [c_void_p] * 3),
    "xeq_invariant_dot_bwdbwd": (c_int, [c_void_p] * 3 + [c_int32] + [c_void_p] * 3 + [c_int32] * 4 + [c_void_p] * 5),
    "xeq_gate_residual_fwd": (c_int, [c_void_p] * 5 + [c_int32] * 4 + [c_void_p] * 3),
    "xeq_gate_residual_bwd": (c_int, [c_void_p] * 5 + [c_int32] * 4 + [c_void_p] * 4),
    "xeq_gate_residual_bwdbwd": (c_int, [c_void_p] * 8 + [c_int32] * 4 + [c_void_p] * 6),
    "xeq_silu_fwd": (c_int, [c_void_p, c_size_t, c_void_p, c_void_p]),
    "xeq_silu_bwd": (c_int, [c_void_p, c_void_p, c_size_t, c_void_p, c_void_p]),
    "xeq_silu_bwdbwd": (c_int, [c_void_p, c_void_p, c_void_p, c_size_t, c_void_p, c_void_p, c_void_p]),
    "xeq_layout_convert": (c_int, [c_void_p, c_void_p, c_int32, POINTER(XeqDims), c_int, c_void_p]),
    "xeq_graph_from_coo_bytes": (c_size_t, [c_int32, c_int32, c_int]),
    "xeq_graph_from_coo": (c_int, [c_void_p] * 4 + [c_int32] * 3 + [c_void_p, c_size_t, POINTER(XeqGraph), c_void_p]),
    "xeq_model_weight_count": (c_size_t, [POINTER(XeqDims), c_int32, c_int32, c_int32, c_int32]),
    "xeq_model_create": (c_int, [POINTER(XeqDims), c_int32, c_int32, c_int32, c_int32, c_void_p, c_size_t, POINTER(c_void_p)]),
    "xeq_model_destroy": (None, [c_void_p]),
    "xeq_model_workspace_bytes": (c_size_t, [c_void_p, POINTER(XeqGraph), c_int]),
    "xeq_model_energy_forces": (c_int, [c_void_p, POINTER(XeqGraph)] + [c_void_p] * 6 + [c_void_p, c_size_t, c_void_p]),
    "xeq_model_energy_forces_virial": (c_int, [c_void_p, POINTER(XeqGraph)] + [c_void_p] * 7 + [c_void_p, c_size_t, c_void_p, c_void_p]),
    "xeq_model_energy_forces_mt": (c_int, [c_void_p, POINTER(XeqGraph)] + [c_void_p] * 6 + [c_void_p, c_size_t, c_void_p, c_void_p]),
}

_lib = None


def exported_symbols():
    return sorted(_SIGNATURES)


def get():
    """The loaded library.  Raises if it has not been built (python -m xequinet_b200.build)."""
    global _lib
    if _lib is None:
        if not LIB_PATH.exists():
            raise RuntimeError(
                f"{LIB_PATH} is missing: build it with `python xequinet_b200/build.py` "
                "(xequinet_b200 has no CPU or eager fallback)")
        lib = ctypes.CDLL(str(LIB_PATH))
        for name, (res, args) in _SIGNATURES.items():
            fn = getattr(lib, name)
            fn.restype, fn.argtypes = res, args
        _lib = lib
    return _lib if _profiler is None else _profiler


# ---------------------------------------------------------------------------------------------------------
# per-entry-point device timing (bench.py: step shares by kernel group, tensor-pipe roofline of K3)
# ---------------------------------------------------------------------------------------------------------
_UNTIMED = {"xeq_graph_from_coo_bytes", "xeq_model_weight_count", "xeq_model_create", "xeq_model_destroy", "xeq_csr_tile_count", "xeq_version", "xeq_last_error", "xeq_num_sms", "xeq_launch_count", "xeq_center_tile_edges", "xeq_neighbor_tile_edges"}


class _Profiled:
    """Proxy of the loaded library that brackets every launching entry point with CUDA events on the current stream.
    Only installed between start_profile() / stop_profile(); the product path calls the library directly."""

    def __init__(self, lib):
        self._lib = lib
        self.records = []  # (entry point, start event, end event, flops)

    def __getattr__(self, name):
        fn = getattr(self._lib, name)
        if name in _UNTIMED or name.endswith("_workspace_bytes"):
            return fn

        def timed(*args):
            a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            a.record()
            rc = fn(*args)
            b.record()
            flops = 0.0
            if name == "xeq_gemm_tf32x3":  # (problems, n_problems, ...): 2 m n k per problem
                flops = float(sum(2.0 * args[0][i].m * args[0][i].n * args[0][i].k for i in range(int(args[1]))))
            self.records.append((name, a, b, flops))
            return rc

        return timed


_profiler = None


def start_profile():
    global _profiler
    _profiler = None
    _profiler = _Profiled(get())


def stop_profile():
    """{entry point: (calls, total ms, total flops)}; synchronises the device."""
    global _profiler
    prof, _profiler = _profiler, None
    torch.cuda.synchronize()
    out = {}
    for name, a, b, fl in (prof.records if prof else []):
        c, t, f = out.get(name, (0, 0.0, 0.0))
        out[name] = (c + 1, t + a.elapsed_time(b), f + fl)
    return out


def check(rc: int, what: str):
    if rc != 0:
        msg = get().xeq_last_error().decode()
        raise RuntimeError(f"{what} failed (code {rc}): {msg}")


def ptr(t):
    """Device pointer of a contiguous CUDA tensor (None -> NULL)."""
    if t is None:
        return None
    if not t.is_cuda:
        raise RuntimeError("xequinet_b200 ops need CUDA tensors: there is no CPU fallback")
    if not t.is_contiguous():
        raise RuntimeError("xequinet_b200 ops need contiguous tensors")
    return t.data_ptr()


def ptr_rows(t):
    """Device pointer of a 2-D CUDA tensor with dense rows (unit inner stride; the row stride travels as an `ld` argument)."""
    if t is None:
        return None
    if not t.is_cuda:
        raise RuntimeError("xequinet_b200 ops need CUDA tensors: there is no CPU fallback")
    if t.dim() != 2 or t.stride(1) != 1 or t.stride(0) < t.shape[1]:
        raise RuntimeError("xequinet_b200 ops need row-dense 2-D tensors here")
    return t.data_ptr()


def stream():
    return torch.cuda.current_stream().cuda_stream
