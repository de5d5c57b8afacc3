// p2e_tape.h -- the "tape": a host-side C++ mirror of the reference's gadget call graph
// (src/gadgets/{nonnative,curve,curve_fixed_base,curve_msm,glv,ecdsa}.rs).  Calling the mirrored
// gadget methods records, in gadget call order,
//   * every in-tree generator instance with the slot offset of its output tuple (the witness layout),
//   * "prep ops": nonnative generators outside curve arithmetic, run by a per-signature interpreter,
//   * "point ops": curve_add / curve_double / curve_conditional_add instances, run by the
//     projective chain-replay + batched-inversion + expansion kernels (p2e_chain.cu),
//   * "checks": connect_nonnative constraints the witness must satisfy (ERR_CONNECT otherwise).
// The tape is identical for every signature, is built once on the host and lives in device memory.
#pragma once
#include <stdint.h>
#include <vector>

namespace p2e {

enum GenKind : uint32_t { G_ADD = 0, G_ADDMANY = 1, G_SUB = 2, G_INV = 3, G_MUL = 4, G_CHECKSUM = 5, G_GLV = 6 };

enum Space : uint8_t { SP_NONE = 0, SP_OUT = 1, SP_IN = 2, SP_CONST = 3, SP_CPOINT = 4 };   // SP_CPOINT: const_points pool (MulRec operands only)

// A nonnative / BigUint target: `nl` limbs starting at `off` in `space`, optionally multiplied by
// a bool target (mul_nonnative_by_bool, nonnative.rs:278-295), itself optionally negated (builder.not).
struct Ref {
  uint8_t space;
  uint8_t nl;
  uint8_t bspace;   // SP_NONE: no bool factor
  uint8_t binv;
  uint32_t off;
  uint32_t boff;
};

enum PrepOpcode : uint8_t { OP_ADD = 0, OP_SUB = 1, OP_MUL = 2, OP_INV = 3, OP_GLV = 4, OP_ADDMANY = 5 };
struct PrepOp {
  uint8_t opcode;
  uint8_t mod;
  uint8_t pad;      // OP_ADDMANY: number of summands, stored at extra_refs[a.boff ...]
  uint8_t round;    // dependency round (see Tape::n_rounds)
  uint32_t slot;    // first output slot
  Ref a, b;
};

struct Check {      // connect_nonnative(a, b): limb-wise equality, missing limbs must be zero
  Ref a, b;
};

enum PtOpType : uint8_t { PT_ADD = 0, PT_DBL = 1, PT_CADD_FIXED = 2, PT_CADD_MSM = 3, PT_CADD_BOOL = 4 };
enum PtFlags : uint8_t { PF_P1_CUR = 1, PF_P2_CUR = 2, PF_SET_CUR = 4, PF_P2_AFFINE = 8 };
constexpr int PF_CHAIN_SHIFT = 4;               // flags bits 4-5: independent chain id (0 fixed-base, 1 msm, 2 tail)
constexpr int N_CHAINS = 3;
constexpr int32_t PID_CONST_BASE = 1 << 20;     // pid >= this: affine constant #(pid - base)
struct PtOp {
  uint8_t type;
  uint8_t flags;
  uint16_t aux;     // window index (fixed base), 2-bit digit index (msm) or input index of the BoolTarget
  int32_t p1, p2;   // point ids (ignored when the *_CUR flag is set)
  uint32_t slot;    // first output slot of this curve op's generators
  uint32_t round;   // dependency round
};

// One mul_nonnative instance (nonnative.rs:390-464) in gadget order: its factors and the first slot of its MulNonnative
// tuple (r[9] q[9] check_sum[17], followed by the CheckSum tuple b[16]).  What the seven range_check_ux_circuit calls of
// nonnative.rs:453-460 read.
struct MulRec {
  Ref x, y;
  uint32_t out;
  uint32_t mod;
};

struct InputPoint {  // per-signature affine point whose coordinates are nonnative targets
  Ref x, y;
  uint32_t round;    // round from which the coordinates exist (loaded by that round's replay kernels)
};

struct Tape {
  // layout
  std::vector<uint32_t> gen_kind, gen_off;
  uint32_t n_slots = 0;
  uint32_t n_inputs = 0;            // input nonnative values (9 wires each)
  // Rounds: a nonnative generator that consumes a coordinate of a curve-op result can only run after
  // the chain kernels produced it.  Round r = prep ops of round r, then the curve ops of round r;
  // values from curve ops of round r are usable by prep ops / points of round r + 1.  The built-in
  // tapes (and most circuits) have a single round.
  uint32_t n_rounds = 1;
  // programs
  std::vector<PrepOp> prep;
  std::vector<PtOp> ptops;
  std::vector<Check> checks;
  std::vector<MulRec> muls;             // every mul_nonnative instance, gadget order
  std::vector<Ref> extra_refs;          // operand lists of OP_ADDMANY
  std::vector<InputPoint> in_points;    // pid = ptops.size() + index
  std::vector<uint32_t> const_limbs;    // SP_CONST pool (u32 limbs)
  std::vector<uint32_t> const_points;   // affine constants: 18 limbs each (x[9], y[9])
  int32_t pre_pid[16];                  // msm table entries (curve_msm.rs:41-62)
  Ref fixed_scalar;                     // u1 (9 limbs) or space NONE
  Ref msm_k1, msm_k2;                   // 5 limbs each or space NONE
  uint32_t fixed_round = 0, msm_round = 0;   // rounds of the fixed-base / msm loops (their scalars exist from then on)
  uint32_t fixed_table_base = 0;        // const point index of window 0 entry 0
  int32_t chain_begin[N_CHAINS] = {0, 0, 0}, chain_end[N_CHAINS] = {0, 0, 0};   // contiguous op ranges
};

enum TapeId : int {
  TAPE_ECDSA = 0,        // verify_secp256k1_message_circuit (gadgets/ecdsa.rs:30-53); inputs msg, r, s, pk.x, pk.y
  TAPE_GLV_MUL = 1,      // glv_mul(p, k) (gadgets/glv.rs:87-104); inputs k, p.x, p.y
  TAPE_CURVE_ADD = 2,    // curve_add(p1, p2) (gadgets/curve.rs:202-223); inputs x1, y1, x2, y2
  TAPE_CURVE_DOUBLE = 3, // curve_double(p) (gadgets/curve.rs:160-185); inputs x, y
  TAPE_CURVE_COND_ADD = 4, // curve_conditional_add(p1, p2, b) (gadgets/curve.rs:225-243); inputs x1, y1, x2, y2, b
  TAPE_FIXED_BASE_MUL = 5, // fixed_base_curve_mul_circuit(G, scalar) (gadgets/curve_fixed_base.rs:18-66); input scalar
  TAPE_CURVE_MSM = 6,    // curve_msm_circuit(p, q, n, m) (gadgets/curve_msm.rs:21-79); inputs p.x, p.y, q.x, q.y, n, m
  N_TAPES = 7
};
constexpr int MAX_TAPE_INPUTS = 16;  // input wire matrices per tape
constexpr int TAPE_USER_BASE = 16;   // ids of tapes made with the builder C-API (p2e_builder_finish)
const Tape& get_tape(int id);
bool tape_exists(int id);
inline const Tape& ecdsa_tape() { return get_tape(TAPE_ECDSA); }
inline const Tape& glv_mul_tape() { return get_tape(TAPE_GLV_MUL); }

}  // namespace p2e
