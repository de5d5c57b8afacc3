"""ctypes binding of libp2e.so (include/p2e.h) plus the thin host-side mirror of the reference's
generator / gate surface for the hot path.

There is NO CPU path here: if the CUDA library is missing or no device is visible every call
raises.  numpy arrays are the host buffers of the C-ABI; torch tensors (device memory / streams)
are accepted by the *_dev methods.

Method names follow the reference: `mul_nonnative` -> MulNonnativeGenerator + CheckSumGenerator
(gates/mul_nonnative.rs:249-324, 513-531), `add_nonnative` / `sub_nonnative` /
`add_many_nonnative` / `inv_nonnative` -> the four generators of gadgets/nonnative.rs:599-895,
`div_rem_biguint` -> BigUintDivRemGenerator (gadgets/biguint.rs:483-548), `decompose_secp256k1_scalar`
-> GLVDecompositionGenerator (gadgets/glv.rs:107-171), `eval_mul_nonnative_gate` /
`eval_check_sum_gate` -> Gate::eval_unfiltered_base_batch of the two in-tree gates
(gates/mul_nonnative.rs:101-130, 411-427), `verify_secp256k1_message_witness` -> the witness of
verify_secp256k1_message_circuit (gadgets/ecdsa.rs:30-53).
"""
from __future__ import annotations

import ctypes
import os

import numpy as np

SECP256K1_BASE = 0
SECP256K1_SCALAR = 1

ERR_U32 = 1
ERR_LIMB_RANGE = 2
ERR_NONCANONICAL = 4
ERR_INVERSE_ZERO = 8
ERR_CARRY_RANGE = 16
ERR_DIV_ZERO = 32
ERR_LIMBS_OVERFLOW = 64
ERR_CONNECT = 128
ERR_INTERNAL = 256

UX_RANGE_CHECK, UX_COMPARISON, UX_ADD_MANY, UX_SUBTRACTION, UX_ARITHMETIC = 0, 1, 2, 3, 4
GATE_MUL_NONNATIVE, GATE_CHECK_SUM, GATE_UX = 0, 1, 2

MUL_OUT_R, MUL_OUT_Q, MUL_OUT_CHECK_SUM, MUL_OUT_CARRY, MUL_OUT_ROWS = 0, 9, 18, 35, 51

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.environ.get("P2E_LIB") or os.path.join(_HERE, "libp2e.so")      # P2E_LIB: an experiment build (scripts/build_variant.sh)

_u64p = ctypes.POINTER(ctypes.c_uint64)
_u32p = ctypes.POINTER(ctypes.c_uint32)
_sz = ctypes.c_size_t
_vp = ctypes.c_void_p
_int = ctypes.c_int
_u64 = ctypes.c_uint64

# name -> (restype, argtypes); every symbol include/p2e.h declares
SIGNATURES = {
    "p2e_init": (_int, [_int]),
    "p2e_init_devices": (_int, [_int]),
    "p2e_set_device": (_int, [_int]),
    "p2e_current_device": (_int, []),
    "p2e_shutdown_device": (_int, [_int]),
    "p2e_set_option": (_int, [ctypes.c_char_p, ctypes.c_long]),
    "p2e_get_option": (ctypes.c_long, [ctypes.c_char_p]),
    "p2e_stream_release": (_int, [_vp]),
    "p2e_debug_f8": (_int, [_int, _sz, _vp, _vp, _vp]),
    "p2e_shutdown": (None, []),
    "p2e_last_error": (ctypes.c_char_p, []),
    "p2e_device_count": (_int, []),
    "p2e_synchronize": (_int, []),
    "p2e_malloc": (_int, [ctypes.POINTER(_vp), _sz]),
    "p2e_free": (_int, [_vp]),
    "p2e_host_alloc": (_int, [ctypes.POINTER(_vp), _sz]),
    "p2e_host_free": (_int, [_vp]),
    "p2e_memcpy_h2d": (_int, [_vp, _vp, _sz]),
    "p2e_memcpy_d2h": (_int, [_vp, _vp, _sz]),
    "p2e_mul_nonnative_witness_dev": (_int, [_int, _sz, _vp, _vp, _sz, _vp, _sz, _vp, _vp]),
    "p2e_mul_nonnative_witness": (_int, [_int, _sz, _vp, _vp, _vp, _vp]),
    "p2e_check_sum_witness_dev": (_int, [_sz, _vp, _sz, _vp, _sz, _vp, _vp]),
    "p2e_check_sum_witness": (_int, [_sz, _vp, _vp, _vp]),
    "p2e_add_nonnative_witness_dev": (_int, [_int, _sz, _vp, _int, _vp, _int, _sz, _vp, _sz, _vp, _vp]),
    "p2e_add_nonnative_witness": (_int, [_int, _sz, _vp, _int, _vp, _int, _vp, _vp]),
    "p2e_sub_nonnative_witness_dev": (_int, [_int, _sz, _vp, _int, _vp, _int, _sz, _vp, _sz, _vp, _vp]),
    "p2e_sub_nonnative_witness": (_int, [_int, _sz, _vp, _int, _vp, _int, _vp, _vp]),
    "p2e_add_many_nonnative_witness_dev": (_int, [_int, _sz, _int, _vp, _sz, _vp, _sz, _vp, _vp]),
    "p2e_add_many_nonnative_witness": (_int, [_int, _sz, _int, _vp, _vp, _vp]),
    "p2e_inv_nonnative_witness_dev": (_int, [_int, _sz, _vp, _sz, _vp, _sz, _vp, _vp]),
    "p2e_inv_nonnative_witness": (_int, [_int, _sz, _vp, _vp, _vp]),
    "p2e_glv_decompose_witness_dev": (_int, [_sz, _vp, _sz, _vp, _sz, _vp, _vp]),
    "p2e_glv_decompose_witness": (_int, [_sz, _vp, _vp, _vp]),
    "p2e_biguint_div_rem_witness_dev": (_int, [_sz, _vp, _int, _vp, _int, _sz, _vp, _sz, _vp, _vp]),
    "p2e_biguint_div_rem_witness": (_int, [_sz, _vp, _int, _vp, _int, _vp, _vp]),
    "p2e_eval_mul_nonnative_dev": (_int, [_int, _sz, _vp, _sz, _vp, _sz, _vp]),
    "p2e_eval_mul_nonnative": (_int, [_int, _sz, _vp, _vp]),
    "p2e_eval_check_sum_dev": (_int, [_sz, _vp, _sz, _vp, _sz, _vp]),
    "p2e_eval_check_sum": (_int, [_sz, _vp, _vp]),
    "p2e_tape_slots": (_sz, [_int]),
    "p2e_tape_inputs": (_int, [_int]),
    "p2e_tape_layout": (_sz, [_int, _vp, _vp]),
    "p2e_tape_witness_dev": (_int, [_int, _sz, _vp, _sz, _vp, _sz, _vp, _vp]),
    "p2e_tape_witness": (_int, [_int, _sz, _vp, _vp, _vp]),
    "p2e_tape_compact_layout": (_sz, [_int, _vp, _vp, _vp]),
    "p2e_tape_num_mul": (_int, [_int]),
    "p2e_tape_mul_info": (_int, [_int, _int, _vp, _vp, _vp, _vp]),
    "p2e_tape_mul_factors_host": (_int, [_int, _int, _vp, _vp, _vp, _vp]),
    "p2e_tape_num_range_checks": (_int, [_int]),
    "p2e_tape_range_check_info": (_int, [_int, _int, _int, _vp, _vp, _vp]),
    "p2e_tape_range_check_layout": (_int, [_int, _int, _vp]),
    "p2e_tape_range_check_rows_dev": (_int, [_int, _sz, _vp, _sz, _vp, _vp, _sz, _int, _int, _int, _vp, _sz, _vp, _vp]),
    "p2e_tape_witness_compact": (_int, [_int, _sz, _vp, _vp, _vp, _vp]),
    "p2e_tape_witness_compact_dev": (_int, [_int, _sz, _vp, _sz, _vp, _vp, _sz, _vp, _vp]),
    "p2e_ecdsa_witness_compact_dev": (_int, [_sz, _vp, _vp, _vp, _vp, _vp, _sz, _vp, _vp, _sz, _vp, _vp]),
    "p2e_tape_compact_expand": (_int, [_int, _sz, _vp, _vp, _vp]),
    "p2e_builder_new": (_int, []),
    "p2e_builder_free": (_int, [_int]),
    "p2e_builder_input": (_int, [_int]),
    "p2e_builder_constant_nonnative": (_int, [_int, _vp]),
    "p2e_builder_add_nonnative": (_int, [_int, _int, _int, _int]),
    "p2e_builder_sub_nonnative": (_int, [_int, _int, _int, _int]),
    "p2e_builder_mul_nonnative": (_int, [_int, _int, _int, _int]),
    "p2e_builder_inv_nonnative": (_int, [_int, _int, _int]),
    "p2e_builder_neg_nonnative": (_int, [_int, _int, _int]),
    "p2e_builder_add_many_nonnative": (_int, [_int, _int, _vp, _int]),
    "p2e_builder_nonnative_conditional_neg": (_int, [_int, _int, _int, _int]),
    "p2e_builder_if_nonnative": (_int, [_int, _int, _int, _int, _int]),
    "p2e_builder_connect_nonnative": (_int, [_int, _int, _int]),
    "p2e_builder_affine_point": (_int, [_int, _int, _int]),
    "p2e_builder_constant_affine_point": (_int, [_int, _vp, _vp]),
    "p2e_builder_point_x": (_int, [_int, _int]),
    "p2e_builder_point_y": (_int, [_int, _int]),
    "p2e_builder_curve_assert_valid": (_int, [_int, _int]),
    "p2e_builder_curve_add": (_int, [_int, _int, _int]),
    "p2e_builder_curve_double": (_int, [_int, _int]),
    "p2e_builder_curve_conditional_add": (_int, [_int, _int, _int, _int]),
    "p2e_builder_fixed_base_curve_mul": (_int, [_int, _int]),
    "p2e_builder_curve_msm": (_int, [_int, _int, _int, _int, _int]),
    "p2e_builder_glv_mul": (_int, [_int, _int, _int]),
    "p2e_builder_finish": (_int, [_int]),
    "p2e_ecdsa_tape_slots": (_sz, []),
    "p2e_ecdsa_tape_layout": (_sz, [_vp, _vp]),
    "p2e_ecdsa_witness_dev": (_int, [_sz, _vp, _vp, _vp, _vp, _vp, _sz, _vp, _sz, _vp, _vp]),
    "p2e_ecdsa_witness": (_int, [_sz, _vp, _vp, _vp, _vp, _vp, _vp, _vp]),
    "p2e_glv_mul_tape_slots": (_sz, []),
    "p2e_glv_mul_tape_layout": (_sz, [_vp, _vp]),
    "p2e_glv_mul_witness_dev": (_int, [_sz, _vp, _vp, _vp, _sz, _vp, _sz, _vp, _vp]),
    "p2e_glv_mul_witness": (_int, [_sz, _vp, _vp, _vp, _vp, _vp]),
    "p2e_ux_gate_shape": (_int, [_int, _int, _int, _int, _int, _vp, _vp, _vp]),
    "p2e_eval_ux_dev": (_int, [_int, _int, _int, _int, _int, _sz, _vp, _sz, _vp, _sz, _vp]),
    "p2e_eval_ux": (_int, [_int, _int, _int, _int, _int, _sz, _vp, _vp]),
    "p2e_ux_witness_dev": (_int, [_int, _int, _int, _int, _int, _sz, _vp, _sz, _vp, _sz, _vp, _vp]),
    "p2e_ux_witness": (_int, [_int, _int, _int, _int, _int, _sz, _vp, _vp, _vp]),
    "p2e_genstream_load": (_int, [_vp, _sz]),
    "p2e_genstream_free": (_int, [_int]),
    "p2e_genstream_info": (ctypes.c_long, [_int, _int]),
    "p2e_genstream_slot": (ctypes.c_long, [_int, _int, _u64, _u64]),
    "p2e_genstream_input_slots": (ctypes.c_long, [_int, _vp, _sz]),
    "p2e_genstream_generator_levels": (ctypes.c_long, [_int, _vp, _sz]),
    "p2e_genstream_run_dev": (_int, [_int, _sz, _vp, _int, _vp, _sz, _vp, _sz, _vp, _vp]),
    "p2e_genstream_run": (_int, [_int, _sz, _vp, _int, _vp, _vp, _int, _vp, _vp]),
    "p2e_gate_shape": (_int, [_int, _vp, _vp, _vp]),
    "p2e_eval_filtered_scratch_bytes": (_sz, [_int, _vp, _sz]),
    "p2e_eval_filtered_combined_dev": (_int, [_int, _vp, _int, _int, _int, _int, _sz, _vp, _sz, _vp, _vp, _int, _u64, _int,
                                              _vp, _sz, _vp, _vp]),
    "p2e_eval_filtered_combined": (_int, [_int, _vp, _int, _int, _int, _int, _sz, _vp, _vp, _vp, _int, _u64, _int, _vp]),
    "p2e_ntt_scratch_bytes": (_sz, [_int, _int, _int]),
    "p2e_ntt_dev": (_int, [_int, _int, _int, _vp, _sz, _vp, _sz, _vp, _vp]),
    "p2e_ntt": (_int, [_int, _int, _int, _vp, _vp]),
    "p2e_coset_lde_dev": (_int, [_int, _int, _int, _vp, _sz, _vp, _sz, _vp, _vp]),
    "p2e_coset_lde": (_int, [_int, _int, _int, _vp, _vp]),
    "p2e_set_profiling": (_int, [_int]),
    "p2e_last_kernel_times": (_int, [_vp, _int]),
    "p2e_chain_trace_dump": (_int, [_vp, _vp, _vp, _int]),
    "p2e_blinding_scalar": (_int, [_vp]),
    "p2e_bench_imad": (ctypes.c_double, [_int, _int]),
}


class P2EError(RuntimeError):
    pass


def load_library(path: str = LIB_PATH) -> ctypes.CDLL:
    """dlopen libp2e.so and bind every symbol of include/p2e.h.  Fails loudly when absent."""
    if not os.path.exists(path):
        raise P2EError(
            f"{path} not found: build it with `python -m plonky2_ecdsa_b200.build` "
            "(nvcc, sm_100a).  This package has no CPU fallback.")
    lib = ctypes.CDLL(path)
    for name, (res, args) in SIGNATURES.items():
        fn = getattr(lib, name)          # AttributeError if the symbol is missing
        fn.restype = res
        fn.argtypes = args
    return lib


def blinding_scalar(lib=None) -> int:
    """keccak(0u64) as a secp256k1 scalar (gadgets/curve_msm.rs:33-39), from the library's own Keccak."""
    lib = lib if lib is not None else load_library()
    w = (ctypes.c_uint64 * 4)()
    if lib.p2e_blinding_scalar(ctypes.addressof(w)) != 0:
        raise P2EError("p2e_blinding_scalar failed")
    return sum(int(w[i]) << (64 * i) for i in range(4))


def _ptr(a):
    """numpy array / torch tensor / None -> raw address."""
    if a is None:
        return None
    if isinstance(a, np.ndarray):
        return a.ctypes.data
    return a.data_ptr()          # torch tensor


def _host_u64(a, rows=None) -> np.ndarray:
    a = np.ascontiguousarray(a, dtype=np.uint64)
    if a.ndim != 2:
        raise ValueError("expected a [rows][n] wire-major matrix")
    if rows is not None and a.shape[0] != rows:
        raise ValueError(f"expected {rows} rows, got {a.shape[0]}")
    return a


class CircuitBuilder:
    """Host mirror of the reference's CircuitBuilder extension traits for the nonnative / curve gadgets
    (CircuitBuilderNonNative gadgets/nonnative.rs:53-164, CircuitBuilderCurve gadgets/curve.rs:34-94,
    fixed_base_curve_mul_circuit, curve_msm_circuit, CircuitBuilderGlv): records the gadget calls through
    the p2e_builder_* C-API and `build()` returns a tape id for Engine.tape_witness().  Works without a GPU
    (building a tape is host work); running it does not."""

    def __init__(self, lib=None):
        self.lib = lib if lib is not None else load_library()
        self.h = self.lib.p2e_builder_new()
        if self.h < 0:
            raise P2EError("p2e_builder_new failed")

    def _ok(self, rc):
        if rc < 0:
            raise P2EError((self.lib.p2e_last_error() or b"").decode())
        return rc

    @staticmethod
    def _limbs(v):
        return np.array([(int(v) >> (29 * i)) & 0x1FFFFFFF for i in range(9)], dtype=np.uint32)

    def add_virtual_nonnative_target(self):
        return self._ok(self.lib.p2e_builder_input(self.h))

    def constant_nonnative(self, value: int):
        l = self._limbs(value)
        return self._ok(self.lib.p2e_builder_constant_nonnative(self.h, l.ctypes.data))

    def add_nonnative(self, modulus, a, b):
        return self._ok(self.lib.p2e_builder_add_nonnative(self.h, modulus, a, b))

    def sub_nonnative(self, modulus, a, b):
        return self._ok(self.lib.p2e_builder_sub_nonnative(self.h, modulus, a, b))

    def mul_nonnative(self, modulus, a, b):
        return self._ok(self.lib.p2e_builder_mul_nonnative(self.h, modulus, a, b))

    def inv_nonnative(self, modulus, x):
        return self._ok(self.lib.p2e_builder_inv_nonnative(self.h, modulus, x))

    def neg_nonnative(self, modulus, x):
        return self._ok(self.lib.p2e_builder_neg_nonnative(self.h, modulus, x))

    def add_many_nonnative(self, modulus, xs):
        arr = np.array(list(xs), dtype=np.int32)
        return self._ok(self.lib.p2e_builder_add_many_nonnative(self.h, modulus, arr.ctypes.data, len(arr)))

    def nonnative_conditional_neg(self, modulus, x, b):
        return self._ok(self.lib.p2e_builder_nonnative_conditional_neg(self.h, modulus, x, b))

    def if_nonnative(self, modulus, b, x, y):
        return self._ok(self.lib.p2e_builder_if_nonnative(self.h, modulus, b, x, y))

    def connect_nonnative(self, a, b):
        self._ok(self.lib.p2e_builder_connect_nonnative(self.h, a, b))

    def affine_point(self, x, y):
        return self._ok(self.lib.p2e_builder_affine_point(self.h, x, y))

    def constant_affine_point(self, pt):
        x, y = self._limbs(pt[0]), self._limbs(pt[1])
        return self._ok(self.lib.p2e_builder_constant_affine_point(self.h, x.ctypes.data, y.ctypes.data))

    def point_x(self, p):
        return self._ok(self.lib.p2e_builder_point_x(self.h, p))

    def point_y(self, p):
        return self._ok(self.lib.p2e_builder_point_y(self.h, p))

    def curve_assert_valid(self, p):
        self._ok(self.lib.p2e_builder_curve_assert_valid(self.h, p))

    def curve_add(self, p1, p2):
        return self._ok(self.lib.p2e_builder_curve_add(self.h, p1, p2))

    def curve_double(self, p):
        return self._ok(self.lib.p2e_builder_curve_double(self.h, p))

    def curve_conditional_add(self, p1, p2, b):
        return self._ok(self.lib.p2e_builder_curve_conditional_add(self.h, p1, p2, b))

    def fixed_base_curve_mul_circuit(self, scalar):
        return self._ok(self.lib.p2e_builder_fixed_base_curve_mul(self.h, scalar))

    def curve_msm_circuit(self, p, q, n, m):
        return self._ok(self.lib.p2e_builder_curve_msm(self.h, p, q, n, m))

    def glv_mul(self, p, k):
        return self._ok(self.lib.p2e_builder_glv_mul(self.h, p, k))

    def build(self) -> int:
        """Registers the recorded circuit as a tape; returns its id."""
        tid = self._ok(self.lib.p2e_builder_finish(self.h))
        self.lib.p2e_builder_free(self.h)
        self.h = -1
        return tid


class Engine:
    """`Engine(device)` creates (or re-selects) the library context of that CUDA device and makes it the calling
    thread's current device.  One process may hold several engines (one per GPU); call `select()` (p2e_set_device)
    before using an engine from a thread whose current device is another one."""

    def __init__(self, device: int = 0, lib_path: str = LIB_PATH):
        self.lib = load_library(lib_path)
        rc = self.lib.p2e_init(device)
        if rc != 0:
            raise P2EError(f"p2e_init({device}) failed: {self.last_error()}")
        self.device = device

    def select(self):
        self._check(self.lib.p2e_set_device(self.device))
        return self

    def set_option(self, name: str, value: int):
        self._check(self.lib.p2e_set_option(name.encode(), int(value)))

    def get_option(self, name: str) -> int:
        return int(self.lib.p2e_get_option(name.encode()))

    # ---------------------------------------------------------------- plumbing
    def last_error(self) -> str:
        return (self.lib.p2e_last_error() or b"").decode()

    def _check(self, rc: int):
        if rc != 0:
            raise P2EError(f"p2e call failed ({rc}): {self.last_error()}")

    def synchronize(self):
        self._check(self.lib.p2e_synchronize())

    def close(self):
        self.lib.p2e_shutdown_device(self.device)

    # ---------------------------------------------------------------- witness generation (host buffers)
    def mul_nonnative(self, modulus, x, y):
        """x, y: [9][n] wires -> (out [51][n] rows r|q|check_sum|carry, err [n])."""
        x, y = _host_u64(x, 9), _host_u64(y, 9)
        n = x.shape[1]
        out = np.empty((MUL_OUT_ROWS, n), dtype=np.uint64)
        err = np.empty(n, dtype=np.uint32)
        self._check(self.lib.p2e_mul_nonnative_witness(modulus, n, _ptr(x), _ptr(y), _ptr(out), _ptr(err)))
        return out, err

    def check_sum(self, a):
        a = _host_u64(a, 17)
        n = a.shape[1]
        b = np.empty((16, n), dtype=np.uint64)
        err = np.empty(n, dtype=np.uint32)
        self._check(self.lib.p2e_check_sum_witness(n, _ptr(a), _ptr(b), _ptr(err)))
        return b, err

    def _addsub(self, fn, modulus, a, b):
        a, b = _host_u64(a), _host_u64(b)
        n = a.shape[1] if a.shape[0] else b.shape[1]
        out = np.empty((10, n), dtype=np.uint64)
        err = np.empty(n, dtype=np.uint32)
        self._check(fn(modulus, n, _ptr(a), a.shape[0], _ptr(b), b.shape[0], _ptr(out), _ptr(err)))
        return out, err

    def add_nonnative(self, modulus, a, b):
        """a: [a_limbs][n], b: [b_limbs][n] -> (out [10][n] = sum[9], overflow; err)."""
        return self._addsub(self.lib.p2e_add_nonnative_witness, modulus, a, b)

    def sub_nonnative(self, modulus, a, b):
        return self._addsub(self.lib.p2e_sub_nonnative_witness, modulus, a, b)

    def add_many_nonnative(self, modulus, summands):
        """summands: [k][9][n] -> (out [10][n], err)."""
        s = np.ascontiguousarray(summands, dtype=np.uint64)
        k, nl, n = s.shape
        assert nl == 9
        out = np.empty((10, n), dtype=np.uint64)
        err = np.empty(n, dtype=np.uint32)
        self._check(self.lib.p2e_add_many_nonnative_witness(modulus, n, k, _ptr(s), _ptr(out), _ptr(err)))
        return out, err

    def inv_nonnative(self, modulus, x):
        """x: [9][n] -> (out [18][n] = div[9], inv[9]; err)."""
        x = _host_u64(x, 9)
        n = x.shape[1]
        out = np.empty((18, n), dtype=np.uint64)
        err = np.empty(n, dtype=np.uint32)
        self._check(self.lib.p2e_inv_nonnative_witness(modulus, n, _ptr(x), _ptr(out), _ptr(err)))
        return out, err

    def decompose_secp256k1_scalar(self, k):
        """k: [9][n] -> (out [12][n] = k1[5], k2[5], k1_neg, k2_neg; err)."""
        k = _host_u64(k, 9)
        n = k.shape[1]
        out = np.empty((12, n), dtype=np.uint64)
        err = np.empty(n, dtype=np.uint32)
        self._check(self.lib.p2e_glv_decompose_witness(n, _ptr(k), _ptr(out), _ptr(err)))
        return out, err

    def div_rem_biguint(self, a, b):
        """a: [a_limbs][n], b: [b_limbs][n] -> (div [div_limbs][n], rem [b_limbs][n], err)."""
        a, b = _host_u64(a), _host_u64(b)
        al, bl = a.shape[0], b.shape[0]
        n = b.shape[1]
        dl = 0 if bl > al + 1 else al - bl + 1
        out = np.empty((dl + bl, n), dtype=np.uint64)
        err = np.empty(n, dtype=np.uint32)
        self._check(self.lib.p2e_biguint_div_rem_witness(n, _ptr(a), al, _ptr(b), bl, _ptr(out), _ptr(err)))
        return out[:dl], out[dl:], err

    # ---------------------------------------------------------------- constraint evaluation (host buffers)
    def eval_mul_nonnative_gate(self, modulus, wires):
        """wires: [53][n_rows] -> constraints [17][n_rows]."""
        w = _host_u64(wires, 53)
        n = w.shape[1]
        out = np.empty((17, n), dtype=np.uint64)
        self._check(self.lib.p2e_eval_mul_nonnative(modulus, n, _ptr(w), _ptr(out)))
        return out

    def eval_check_sum_gate(self, wires):
        """wires: [33][n_rows] -> constraints [17][n_rows]."""
        w = _host_u64(wires, 33)
        n = w.shape[1]
        out = np.empty((17, n), dtype=np.uint64)
        self._check(self.lib.p2e_eval_check_sum(n, _ptr(w), _ptr(out)))
        return out

    # ---------------------------------------------------------------- plonky2_ux gates (parity unpinned)
    def ux_gate_shape(self, gate, bits, limb_bits, count=1, num_addends=0):
        """(num_wires, num_constraints, num_inputs) of a plonky2_ux gate row."""
        nw, nc, ni = ctypes.c_int(), ctypes.c_int(), ctypes.c_int()
        self._check(self.lib.p2e_ux_gate_shape(gate, bits, limb_bits, count, num_addends, ctypes.addressof(nw),
                                               ctypes.addressof(nc), ctypes.addressof(ni)))
        return nw.value, nc.value, ni.value

    def ux_gate_witness(self, gate, inputs, bits, limb_bits, count=1, num_addends=0):
        """The gate's generator (SimpleGenerator::run_once) over n instances: inputs [num_inputs][n] ->
        (complete gate rows [num_wires][n], err [n])."""
        nw, _, ni = self.ux_gate_shape(gate, bits, limb_bits, count, num_addends)
        x = _host_u64(inputs, ni)
        n = x.shape[1]
        row = np.empty((nw, n), dtype=np.uint64)
        err = np.empty(n, dtype=np.uint32)
        self._check(self.lib.p2e_ux_witness(gate, bits, limb_bits, count, num_addends, n, _ptr(x), _ptr(row), _ptr(err)))
        return row, err

    def eval_ux_gate(self, gate, wires, bits, limb_bits, count=1, num_addends=0):
        """Gate::eval_unfiltered_base_batch: wires [num_wires][n_rows] -> constraints [num_constraints][n_rows]."""
        nw, nc, _ = self.ux_gate_shape(gate, bits, limb_bits, count, num_addends)
        w = _host_u64(wires, nw)
        n = w.shape[1]
        out = np.empty((nc, n), dtype=np.uint64)
        self._check(self.lib.p2e_eval_ux(gate, bits, limb_bits, count, num_addends, n, _ptr(w), _ptr(out)))
        return out

    def range_check_ux_witness(self, limbs, bits=29, aux_bits=3):
        """range_check_ux_circuit::<BITS>(limbs) (gadgets/nonnative.rs:454-460): one gate row per call."""
        limbs = _host_u64(limbs)
        return self.ux_gate_witness(UX_RANGE_CHECK, limbs, bits, aux_bits, limbs.shape[0])

    def eval_range_check_ux_gate(self, wires, num_input_limbs, bits=29, aux_bits=3):
        return self.eval_ux_gate(UX_RANGE_CHECK, wires, bits, aux_bits, num_input_limbs)

    def comparison_ux_witness(self, first, second, num_bits=29, chunk_bits=2):
        """One ComparisonGate row of list_le_ux_circuit (gadgets/biguint.rs:222-229): result = first <= second."""
        x = np.stack([np.asarray(first, dtype=np.uint64), np.asarray(second, dtype=np.uint64)])
        return self.ux_gate_witness(UX_COMPARISON, x, num_bits, chunk_bits, 1)

    def eval_comparison_ux_gate(self, wires, num_bits=29, chunk_bits=2):
        return self.eval_ux_gate(UX_COMPARISON, wires, num_bits, chunk_bits, 1)

    def add_many_ux_witness(self, inputs, num_addends=3, num_ops=5, bits=29, limb_bits=2):
        """add_many_ux (gadgets/biguint.rs:261): per op the addends then the carry-in."""
        return self.ux_gate_witness(UX_ADD_MANY, inputs, bits, limb_bits, num_ops, num_addends)

    def eval_add_many_ux_gate(self, wires, num_addends=3, num_ops=5, bits=29, limb_bits=2):
        return self.eval_ux_gate(UX_ADD_MANY, wires, bits, limb_bits, num_ops, num_addends)

    # ---------------------------------------------------------------- serialized generator streams
    def load_generator_stream(self, data: bytes):
        """p2e_genstream_load: the level-scheduled executor over a serialized generator list."""
        from .genstream import GeneratorStream
        return GeneratorStream(self.lib, data, P2EError)

    # ---------------------------------------------------------------- selector filtering + alpha-combination
    def gate_shape(self, gate, params=()):
        """(num_wires, num_constraints) of GATE_MUL_NONNATIVE (modulus,), GATE_CHECK_SUM () or GATE_UX
        (ux gate, bits, limb_bits, count, num_addends)."""
        prm = (ctypes.c_int * 5)(*(list(params) + [0] * 5)[:5])
        nw, nc = ctypes.c_int(), ctypes.c_int()
        self._check(self.lib.p2e_gate_shape(gate, ctypes.addressof(prm), ctypes.addressof(nw), ctypes.addressof(nc)))
        return nw.value, nc.value

    def eval_filtered_combined(self, gate, params, wires, alphas, selector=None, gate_index=0, group_range=(0, 1),
                               many_selectors=False, first_power=0, accumulate_into=None):
        """plonky2's Gate::eval_filtered_base_batch followed by this gate's share of
        reduce_with_powers_multi (include/p2e.h): wires [num_wires][n_rows], selector [n_rows] or None ->
        [len(alphas)][n_rows]; `accumulate_into` (same shape) is added to, as
        evaluate_gate_constraints_base_batch sums the gates of a circuit."""
        nw, _ = self.gate_shape(gate, params)
        w = _host_u64(wires, nw)
        n = w.shape[1]
        prm = (ctypes.c_int * 5)(*(list(params) + [0] * 5)[:5])
        al = np.ascontiguousarray(np.asarray(alphas, dtype=np.uint64))
        sel = None if selector is None else np.ascontiguousarray(np.asarray(selector, dtype=np.uint64))
        if sel is not None and sel.shape != (n,):
            raise ValueError("selector must hold one constant per row")
        if accumulate_into is None:
            out = np.empty((len(al), n), dtype=np.uint64)
        else:
            out = np.ascontiguousarray(np.asarray(accumulate_into, dtype=np.uint64)).copy()
            if out.shape != (len(al), n):
                raise ValueError("accumulate_into must be [n_alphas][n_rows]")
        self._check(self.lib.p2e_eval_filtered_combined(gate, ctypes.addressof(prm), gate_index, group_range[0],
                                                        group_range[1], int(many_selectors), n, _ptr(w),
                                                        None if sel is None else _ptr(sel), _ptr(al), len(al),
                                                        first_power, int(accumulate_into is not None), _ptr(out)))
        return out

    # ---------------------------------------------------------------- Goldilocks NTT / coset LDE
    def fft(self, values, inverse=False):
        """plonky2 `fft` / `ifft` of every row: values [n_polys][2^k] -> same shape, natural order."""
        v = _host_u64(values)
        n = v.shape[1]
        log_n = n.bit_length() - 1
        if 1 << log_n != n:
            raise ValueError("row length must be a power of two")
        out = np.empty_like(v)
        self._check(self.lib.p2e_ntt(log_n, 1 if inverse else 0, v.shape[0], _ptr(v), _ptr(out)))
        return out

    def lde_onto_coset(self, values, rate_bits):
        """PolynomialValues::lde_onto_coset(rate_bits) of every row: [n_polys][n] -> [n_polys][n << rate_bits],
        out[p][j] = P_p(7 * omega^j): the wire-major LDE matrix the gate evaluators read."""
        v = _host_u64(values)
        n = v.shape[1]
        log_n = n.bit_length() - 1
        if 1 << log_n != n:
            raise ValueError("row length must be a power of two")
        out = np.empty((v.shape[0], n << rate_bits), dtype=np.uint64)
        self._check(self.lib.p2e_coset_lde(log_n, rate_bits, v.shape[0], _ptr(v), _ptr(out)))
        return out

    # ---------------------------------------------------------------- generic gadget tapes
    (TAPE_ECDSA, TAPE_GLV_MUL, TAPE_CURVE_ADD, TAPE_CURVE_DOUBLE, TAPE_CURVE_COND_ADD, TAPE_FIXED_BASE_MUL,
     TAPE_CURVE_MSM) = range(7)

    def tape_layout(self, tape):
        cnt = self.lib.p2e_tape_layout(tape, None, None)
        kinds = np.empty(cnt, dtype=np.uint32)
        offs = np.empty(cnt, dtype=np.uint32)
        self.lib.p2e_tape_layout(tape, _ptr(kinds), _ptr(offs))
        return kinds, offs, int(self.lib.p2e_tape_slots(tape))

    def tape_witness(self, tape, inputs):
        """inputs: list of [9][n] wire matrices -> (witness [slots][n], err [n])."""
        ins = [_host_u64(a, 9) for a in inputs]
        if len(ins) != self.lib.p2e_tape_inputs(tape):
            raise ValueError("wrong number of tape inputs")
        n = ins[0].shape[1]
        slots = int(self.lib.p2e_tape_slots(tape))
        out = np.empty((slots, n), dtype=np.uint64)
        err = np.empty(n, dtype=np.uint32)
        arr = (ctypes.c_void_p * len(ins))(*[a.ctypes.data for a in ins])
        self._check(self.lib.p2e_tape_witness(tape, n, ctypes.cast(arr, ctypes.c_void_p), _ptr(out), _ptr(err)))
        return out, err

    # ---- the plonky2_ux range-check rows of mul_nonnative (nonnative.rs:453-460)
    def tape_mul_info(self, tape):
        """-> list of (x_limbs, y_limbs, out_slot, modulus) for every mul_nonnative instance of the tape, gadget order."""
        out = []
        xl, yl, md = ctypes.c_int(0), ctypes.c_int(0), ctypes.c_int(0)
        sl = ctypes.c_uint32(0)
        for j in range(self.lib.p2e_tape_num_mul(tape)):
            self._check(self.lib.p2e_tape_mul_info(tape, j, ctypes.addressof(xl), ctypes.addressof(yl), ctypes.addressof(sl), ctypes.addressof(md)))
            out.append((xl.value, yl.value, sl.value, md.value))
        return out

    def tape_range_checks(self, tape, limb_bits=3):
        """Every range_check_ux_circuit call of the tape, gadget order -> (generator index [c], inputs [c], bits [c],
        wire offsets [c + 1]) as numpy arrays (p2e_tape_range_check_info / _layout)."""
        nc = self.lib.p2e_tape_num_range_checks(tape)
        gen = np.empty(nc, dtype=np.int32)
        cnt = np.empty(nc, dtype=np.int32)
        bits = np.empty(nc, dtype=np.int32)
        off = np.empty(nc + 1, dtype=np.uint64)
        self._check(self.lib.p2e_tape_range_check_info(tape, 0, nc, _ptr(gen), _ptr(cnt), _ptr(bits)))
        self._check(self.lib.p2e_tape_range_check_layout(tape, limb_bits, _ptr(off)))
        return gen, cnt, bits, off

    def tape_range_check_rows(self, tape, inputs, witness, c0, c1, limb_bits=3, compact=False):
        """The gate rows of range_check_ux_circuit calls [c0, c1) for every signature, from a finished witness (plain
        [slots][n] uint64, or with compact=True the narrow [rows][n] uint32 matrix of the compact format).  Host arrays in,
        (rows [wires of the calls][n] uint64, err [n]) out -- through device memory (torch)."""
        import torch
        ins = [torch.from_numpy(_host_u64(a, 9).view(np.int64)).cuda() for a in inputs]
        n = int(witness.shape[1])
        wd = torch.from_numpy(np.ascontiguousarray(witness).view(np.int32 if compact else np.int64)).cuda()
        off = self.tape_range_checks(tape, limb_bits)[3]
        rows = torch.empty((max(int(off[c1] - off[c0]), 1), n), dtype=torch.int64, device="cuda")
        err = torch.empty(n, dtype=torch.int32, device="cuda")
        arr = (ctypes.c_void_p * max(len(ins), 1))(*[t.data_ptr() for t in ins])
        self._check(self.lib.p2e_tape_range_check_rows_dev(tape, n, ctypes.cast(arr, ctypes.c_void_p), n, None if compact else wd.data_ptr(),
                                                           wd.data_ptr() if compact else None, n, c0, c1, limb_bits, rows.data_ptr(), n,
                                                           err.data_ptr(), None))
        self.synchronize()
        return rows.cpu().numpy().view(np.uint64)[: int(off[c1] - off[c0])], err.cpu().numpy().view(np.uint32)

    def tape_compact_layout(self, tape):
        """-> (map [slots] uint32, narrow_rows, wide_rows).  map[slot] = (1 << 31 | wide row) for a full field
        element, (1 << 30 | narrow row) for a 34-bit CheckSum carry (low word there; its two top bits are
        bits 2j, 2j+1 of the narrow row that follows the tuple's 16 low words), else the narrow row."""
        slots = int(self.lib.p2e_tape_slots(tape))
        m = np.empty(slots, dtype=np.uint32)
        nn, nw = ctypes.c_uint32(0), ctypes.c_uint32(0)
        self.lib.p2e_tape_compact_layout(tape, _ptr(m), ctypes.addressof(nn), ctypes.addressof(nw))
        return m, int(nn.value), int(nw.value)

    def tape_witness_compact(self, tape, inputs, derive_check_sum=False):
        """Like tape_witness but in the compact host format: (narrow [ns][n] uint32, wide [ws][n] uint64, err).
        derive_check_sum: the derived form -- no `wide` matrix is shipped (returned as None), the check_sum
        slots are rebuilt from the carries by expand_compact (include/p2e.h)."""
        ins = [_host_u64(a, 9) for a in inputs]
        n = ins[0].shape[1]
        _, nn, nw = self.tape_compact_layout(tape)
        narrow = np.empty((nn, n), dtype=np.uint32)
        wide = None if derive_check_sum else np.empty((nw, n), dtype=np.uint64)
        err = np.empty(n, dtype=np.uint32)
        arr = (ctypes.c_void_p * len(ins))(*[a.ctypes.data for a in ins])
        self._check(self.lib.p2e_tape_witness_compact(tape, n, ctypes.cast(arr, ctypes.c_void_p), _ptr(narrow),
                                                      None if wide is None else _ptr(wide), _ptr(err)))
        return narrow, wide, err

    @staticmethod
    def expand_compact(cmap, narrow, wide=None):
        """Compact -> plain [slots][n] uint64 (host-side helper for tests / the Rust shim's unpack).
        wide=None: the derived form -- check_sum slots are rebuilt from the carries (include/p2e.h)."""
        out = np.empty((cmap.shape[0], narrow.shape[1]), dtype=np.uint64)
        is_wide = (cmap >> np.uint32(31)).astype(bool)
        is_carry = ((cmap >> np.uint32(30)) & np.uint32(1)).astype(bool) & ~is_wide
        plain = ~is_wide & ~is_carry
        out[plain] = narrow[cmap[plain]]
        # CheckSum carries: 16 low words, then one word holding the 16 two-bit tops of the tuple
        rows = (cmap[is_carry] & np.uint32(0x1FFFFFFF)).astype(np.int64)
        j = np.arange(rows.shape[0]) % 16
        tops = (narrow[rows - j + 16].astype(np.uint64) >> (2 * j[:, None]).astype(np.uint64)) & np.uint64(3)
        out[is_carry] = narrow[rows].astype(np.uint64) | (tops << np.uint64(32))
        if wide is not None:
            out[is_wide] = wide[cmap[is_wide] & np.uint32(0x7FFFFFFF)]
        else:
            # check_sum_j = c_j * 2^29 - c_{j-1} with c_j = b_j - 2^33 (c_{-1} = c_16 = 0), as a Goldilocks element;
            # the carries of a MUL tuple's CheckSum tuple sit 17 slots after the matching check_sum slot
            ws = np.nonzero(is_wide)[0]
            jj = (cmap[ws] & np.uint32(0x7FFFFFFF)).astype(np.int64) % 17
            off = np.int64(1) << np.int64(33)
            cur = np.where((jj < 16)[:, None], out[np.minimum(ws + 17, cmap.shape[0] - 1)].astype(np.int64) - off, 0)
            prev = np.where((jj > 0)[:, None], out[ws + 16].astype(np.int64) - off, 0)
            v = cur * (np.int64(1) << np.int64(29)) - prev                     # |v| < 2^62
            gl_p = np.uint64(0xFFFFFFFF00000001)
            out[ws] = np.where(v < 0, (v.astype(np.uint64) + gl_p), v.astype(np.uint64))
        return out

    def curve_add(self, x1, y1, x2, y2):
        """curve_add(p1, p2) witness (gadgets/curve.rs:202-223): 231 slots per instance."""
        return self.tape_witness(self.TAPE_CURVE_ADD, [x1, y1, x2, y2])

    def curve_double(self, x, y):
        """curve_double(p) witness (gadgets/curve.rs:160-185): 282 slots per instance."""
        return self.tape_witness(self.TAPE_CURVE_DOUBLE, [x, y])

    def curve_conditional_add(self, x1, y1, x2, y2, b):
        """curve_conditional_add(p1, p2, b) witness (gadgets/curve.rs:225-243); b: [n] of 0/1."""
        bm = np.zeros((9, len(b)), dtype=np.uint64)
        bm[0] = np.asarray(b, dtype=np.uint64)
        return self.tape_witness(self.TAPE_CURVE_COND_ADD, [x1, y1, x2, y2, bm])

    def fixed_base_curve_mul(self, scalar):
        """fixed_base_curve_mul_circuit(G, scalar) witness (gadgets/curve_fixed_base.rs:18-66)."""
        return self.tape_witness(self.TAPE_FIXED_BASE_MUL, [scalar])

    def curve_msm(self, px, py, qx, qy, n, m):
        """curve_msm_circuit(p, q, n, m) witness = n*p + m*q (gadgets/curve_msm.rs:21-79), 9-limb scalars."""
        return self.tape_witness(self.TAPE_CURVE_MSM, [px, py, qx, qy, n, m])

    # ---------------------------------------------------------------- ECDSA circuit witness
    def ecdsa_tape_layout(self):
        cnt = self.lib.p2e_ecdsa_tape_layout(None, None)
        kinds = np.empty(cnt, dtype=np.uint32)
        offs = np.empty(cnt, dtype=np.uint32)
        self.lib.p2e_ecdsa_tape_layout(_ptr(kinds), _ptr(offs))
        return kinds, offs, int(self.lib.p2e_ecdsa_tape_slots())

    def verify_secp256k1_message_witness(self, msg, r, s, pkx, pky):
        """msg, r, s, pkx, pky: [9][n] -> (witness [slots][n] in tape order, err [n])."""
        ins = [_host_u64(a, 9) for a in (msg, r, s, pkx, pky)]
        n = ins[0].shape[1]
        slots = int(self.lib.p2e_ecdsa_tape_slots())
        out = np.empty((slots, n), dtype=np.uint64)
        err = np.empty(n, dtype=np.uint32)
        self._check(self.lib.p2e_ecdsa_witness(n, *[_ptr(a) for a in ins], _ptr(out), _ptr(err)))
        return out, err

    def glv_mul_tape_layout(self):
        cnt = self.lib.p2e_glv_mul_tape_layout(None, None)
        kinds = np.empty(cnt, dtype=np.uint32)
        offs = np.empty(cnt, dtype=np.uint32)
        self.lib.p2e_glv_mul_tape_layout(_ptr(kinds), _ptr(offs))
        return kinds, offs, int(self.lib.p2e_glv_mul_tape_slots())

    def glv_mul_witness(self, k, px, py):
        ins = [_host_u64(a, 9) for a in (k, px, py)]
        n = ins[0].shape[1]
        slots = int(self.lib.p2e_glv_mul_tape_slots())
        out = np.empty((slots, n), dtype=np.uint64)
        err = np.empty(n, dtype=np.uint32)
        self._check(self.lib.p2e_glv_mul_witness(n, *[_ptr(a) for a in ins], _ptr(out), _ptr(err)))
        return out, err

    # ---------------------------------------------------------------- measurement
    CHAIN_KERNELS = ["k_prep", "k_chain_replay", "k_chain_replay(join)", "k_chain_invert", "k_chain_affine",
                     "k_chain_expand", "k_chain_checks"]

    def set_profiling(self, on: bool):
        self._check(self.lib.p2e_set_profiling(1 if on else 0))

    def last_kernel_times(self):
        buf = np.zeros(64, dtype=np.float32)
        k = self.lib.p2e_last_kernel_times(_ptr(buf), 64)
        if k < 0:
            raise P2EError(f"p2e_last_kernel_times failed: {self.last_error()}")
        return [float(v) for v in buf[:k]]

    def bench_imad(self, wide: bool, iters: int = 4096) -> float:
        v = self.lib.p2e_bench_imad(1 if wide else 0, iters)
        if v < 0:
            raise P2EError(f"p2e_bench_imad failed: {self.last_error()}")
        return v
