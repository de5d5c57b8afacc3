"""B200-native witness-generation / constraint-evaluation engine for the hot path of
Weobe/plonky2-ecdsa (29-bit-limb nonnative secp256k1 generators, MulNonnative/CheckSum gates).

The compute path is CUDA only (plonky2_ecdsa_b200/libp2e.so, built by plonky2_ecdsa_b200.build).
"""
from .engine import (  # noqa: F401
    Engine, CircuitBuilder, P2EError, load_library, LIB_PATH, SIGNATURES,
    SECP256K1_BASE, SECP256K1_SCALAR,
    ERR_U32, ERR_LIMB_RANGE, ERR_NONCANONICAL, ERR_INVERSE_ZERO, ERR_CARRY_RANGE,
    ERR_DIV_ZERO, ERR_LIMBS_OVERFLOW, ERR_CONNECT, ERR_INTERNAL,
    UX_RANGE_CHECK, UX_COMPARISON, UX_ADD_MANY, UX_SUBTRACTION, UX_ARITHMETIC,
    GATE_MUL_NONNATIVE, GATE_CHECK_SUM, GATE_UX,
    MUL_OUT_R, MUL_OUT_Q, MUL_OUT_CHECK_SUM, MUL_OUT_CARRY, MUL_OUT_ROWS,
)
from .genstream import GeneratorStream, GeneratorStreamWriter, TargetCircuitBuilder, wire, virtual  # noqa: F401,E402
