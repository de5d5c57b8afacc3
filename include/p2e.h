/* p2e.h -- C-ABI of the B200 witness / constraint engine for Weobe/plonky2-ecdsa's hot path.
 *
 * Every entry point replaces, in BATCHED form (n independent instances per call), one piece of
 * the reference's SimpleGenerator::run_once / Gate::eval_unfiltered surface.  The reference
 * interface each one stands in for is cited as file:line of the reference repository.
 *
 * Conventions
 *   - All field elements on the boundary are canonical Goldilocks u64 (what a plonky2 wire
 *     holds).  A nonnative value is 9 little-endian 29-bit limbs, one per wire.
 *   - Matrices are WIRE-MAJOR ("structure of arrays"): element (wire w, instance i) lives at
 *     ptr[w * stride + i], stride >= n, exactly plonky2's EvaluationVarsBaseBatch layout.
 *   - `*_dev` entry points take DEVICE pointers and a cudaStream_t (as void*; NULL = the
 *     library's stream) and only enqueue work.  The un-suffixed entry points take HOST
 *     pointers, copy in, run, copy out and synchronise (stride == n).
 *   - Return value: 0 on success, negative on CUDA / argument errors (p2e_last_error()).
 *   - Per-instance contract violations (where the reference panics) never abort: they set
 *     bits in err[i] (P2E_ERR_*), the analogue of the reference's panic for that instance.
 *   - There is no CPU fallback: without a CUDA device every compute entry point fails.
 */
#ifndef P2E_H
#define P2E_H
#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define P2E_SECP256K1_BASE 0   /* plonky2::field::secp256k1_base::Secp256K1Base   */
#define P2E_SECP256K1_SCALAR 1 /* plonky2::field::secp256k1_scalar::Secp256K1Scalar */

#define P2E_ERR_U32 1u             /* "Value doesn't fit in u32"      gates/mul_nonnative.rs:262,271 */
#define P2E_ERR_LIMB_RANGE 2u      /* assert limb < 2^29              gates/mul_nonnative.rs:275-276 */
#define P2E_ERR_NONCANONICAL 4u    /* from_noncanonical_biguint >= 2^256 (plonky2 field)             */
#define P2E_ERR_INVERSE_ZERO 8u    /* FF::inverse() of zero           gadgets/nonnative.rs:863       */
#define P2E_ERR_CARRY_RANGE 16u    /* assert b_i + 2^33 < 2^34        gates/mul_nonnative.rs:527     */
#define P2E_ERR_DIV_ZERO 32u       /* BigUint div_rem by zero         gadgets/biguint.rs:515         */
#define P2E_ERR_LIMBS_OVERFLOW 64u /* assert num_limbs >= limbs.len() gadgets/biguint.rs:473         */
#define P2E_ERR_CONNECT 128u       /* a connect_* / assert_* of the gadget cannot be satisfied       */
#define P2E_ERR_INTERNAL 256u      /* device-side self check failed                                  */

/* rows of the output matrix of p2e_mul_nonnative_witness (gate wire order, mul_nonnative.rs:41-59,384-390) */
#define P2E_MUL_OUT_R 0          /* 9 rows : MulNonnativeGate wires 18..26            */
#define P2E_MUL_OUT_Q 9          /* 9 rows : MulNonnativeGate wires 27..35            */
#define P2E_MUL_OUT_CHECK_SUM 18 /* 17 rows: MulNonnativeGate wires 36..52 (= CheckSumGate a) */
#define P2E_MUL_OUT_CARRY 35     /* 16 rows: CheckSumGate wires 17..32 (b_i + 2^33)   */
#define P2E_MUL_OUT_ROWS 51

/* ---- lifetime, devices, threads --------------------------------------------------------
 * The reference is ONE Rust process; a drop-in host must be able to reach every GPU of the box from it.
 *   p2e_init(device)         creates the context of CUDA device `device` (stream, scratch, device copies of the
 *                            tapes / tables) and makes it the CALLING THREAD's current device.  May be called once
 *                            per device; calling it again only re-selects the device.
 *   p2e_init_devices(n)      = p2e_init(0 .. n-1); the calling thread keeps (or gets) device 0.
 *   p2e_set_device(device)   selects the calling thread's current device (like cudaSetDevice).  Every other entry
 *                            point works on the current device of the thread that calls it; a thread that never
 *                            chose one uses the first device that was initialised.
 * Threading contract:
 *   - `*_dev` entry points are re-entrant per (device, stream): different host threads may enqueue on different
 *     streams / devices concurrently.  They never block on the GPU.
 *   - Host-pointer entry points (no `_dev`) share the context's stream and scratch buffers: the library
 *     serialises them PER DEVICE with a mutex, so concurrent callers (plonky2 evaluates gates from rayon
 *     workers) are safe but run one after another on one device; use one thread per device, or the `_dev`
 *     forms with your own streams, for concurrency.
 *   - p2e_last_error() is per thread. */
int p2e_init(int device);
int p2e_init_devices(int n_devices);
int p2e_set_device(int device);
int p2e_current_device(void);           /* -1: none */
int p2e_shutdown_device(int device);    /* releases one device's context */
void p2e_shutdown(void);                /* releases every context */
const char* p2e_last_error(void);
int p2e_device_count(void);
int p2e_synchronize(void);              /* the current device's library stream */
/* Tunables (process-wide; defaults in parentheses): "chain_chunk" (65536) signatures per scratch chunk of the chain
 * kernels; "host_chunks" (4) D2H / compute pipeline depth of p2e_tape_witness_compact; "genstream_fused" (1) and
 * "genstream_graph" (1) executor forms of p2e_genstream_run*; "genstream_walk" (1: streams of 32 levels or more run as ONE launch of the level
 * walker -- a CTA owns 32 circuits and walks every level itself; 2: always; 0: never), "genstream_walk_shape" (0: by batch size; else 100 * warps + circuits per CTA); "ntt_log_tile" (12); "chain_trace" (0) diagnostic timeline; "chain_phases" (63) diagnostic bit mask of the chain kernels to launch
 * (anything else flags every row P2E_ERR_INTERNAL); "chain_replay_lanes" (0: two lanes per (signature, chain) below 4096 signatures per call, one lane from there; 1 / 2: always) lanes of
 * the Jacobian replay kernel; "chain_graph" (1) replay a repeated chain call (same tape, sizes, buffers, stream) as one CUDA graph launch; "chain_split" (0; 1: on -- for hosts that keep fewer than ~24 caller streams in flight) run the
 * per-signature kernels of a chain call on an internal high-priority stream and the full-GPU kernels on a low-priority one.  Nothing is read from the
 * environment. */
int p2e_set_option(const char* name, long value);
long p2e_get_option(const char* name);
/* The chain entry points keep one scratch arena per caller stream (up to 64 per device; an idle arena is recycled
 * when a 65th stream shows up).  A host that is done with a stream can return its arena right away. */
int p2e_stream_release(void* stream);

/* device / pinned-host memory for callers that are not torch (Rust/C hosts) */
int p2e_malloc(void** dptr, size_t bytes);
int p2e_free(void* dptr);
int p2e_host_alloc(void** hptr, size_t bytes); /* pinned */
int p2e_host_free(void* hptr);
int p2e_memcpy_h2d(void* dst, const void* src, size_t bytes);
int p2e_memcpy_d2h(void* dst, const void* src, size_t bytes);

/* ---- witness generation --------------------------------------------------------------- */

/* MulNonnativeGenerator::run_once + CheckSumGenerator::run_once
 * (gates/mul_nonnative.rs:249-324, 513-531).  x, y: [9][stride] wires.  out: [51][stride]
 * (rows P2E_MUL_OUT_*).  err: [n]. */
int p2e_mul_nonnative_witness_dev(int modulus, size_t n, const uint64_t* x, const uint64_t* y,
                                  size_t in_stride, uint64_t* out, size_t out_stride,
                                  uint32_t* err, void* stream);
int p2e_mul_nonnative_witness(int modulus, size_t n, const uint64_t* x, const uint64_t* y,
                              uint64_t* out, uint32_t* err);

/* CheckSumGenerator::run_once alone (gates/mul_nonnative.rs:513-531): a [17][stride] -> b [16][stride] */
int p2e_check_sum_witness_dev(size_t n, const uint64_t* a, size_t in_stride, uint64_t* b,
                              size_t out_stride, uint32_t* err, void* stream);
int p2e_check_sum_witness(size_t n, const uint64_t* a, uint64_t* b, uint32_t* err);

/* NonNativeAdditionGenerator::run_once (gadgets/nonnative.rs:626-645).
 * a: [a_limbs][stride], b: [b_limbs][stride] (limb counts 0..9, as the BigUintTargets have);
 * out: [10][stride] = sum[9], overflow. */
int p2e_add_nonnative_witness_dev(int modulus, size_t n, const uint64_t* a, int a_limbs,
                                  const uint64_t* b, int b_limbs, size_t in_stride, uint64_t* out,
                                  size_t out_stride, uint32_t* err, void* stream);
int p2e_add_nonnative_witness(int modulus, size_t n, const uint64_t* a, int a_limbs,
                              const uint64_t* b, int b_limbs, uint64_t* out, uint32_t* err);

/* NonNativeSubtractionGenerator::run_once (gadgets/nonnative.rs:792-810): out = diff[9], overflow */
int p2e_sub_nonnative_witness_dev(int modulus, size_t n, const uint64_t* a, int a_limbs,
                                  const uint64_t* b, int b_limbs, size_t in_stride, uint64_t* out,
                                  size_t out_stride, uint32_t* err, void* stream);
int p2e_sub_nonnative_witness(int modulus, size_t n, const uint64_t* a, int a_limbs,
                              const uint64_t* b, int b_limbs, uint64_t* out, uint32_t* err);

/* NonNativeMultipleAddsGenerator::run_once (gadgets/nonnative.rs:696-728).
 * summands: [k][9][stride] (k summands of 9 limbs each); out = sum[9], overflow (u32). */
int p2e_add_many_nonnative_witness_dev(int modulus, size_t n, int k, const uint64_t* summands,
                                       size_t in_stride, uint64_t* out, size_t out_stride,
                                       uint32_t* err, void* stream);
int p2e_add_many_nonnative_witness(int modulus, size_t n, int k, const uint64_t* summands,
                                   uint64_t* out, uint32_t* err);

/* NonNativeInverseGenerator::run_once (gadgets/nonnative.rs:857-872).
 * x: [9][stride]; out: [18][stride] = div[9], inv[9] (the order of the generator's set_* calls). */
int p2e_inv_nonnative_witness_dev(int modulus, size_t n, const uint64_t* x, size_t in_stride,
                                  uint64_t* out, size_t out_stride, uint32_t* err, void* stream);
int p2e_inv_nonnative_witness(int modulus, size_t n, const uint64_t* x, uint64_t* out, uint32_t* err);

/* GLVDecompositionGenerator::run_once (gadgets/glv.rs:128-142 -> curve/glv.rs:39-77).
 * k: [9][stride]; out: [12][stride] = k1[5], k2[5], k1_neg, k2_neg. */
int p2e_glv_decompose_witness_dev(size_t n, const uint64_t* k, size_t in_stride, uint64_t* out,
                                  size_t out_stride, uint32_t* err, void* stream);
int p2e_glv_decompose_witness(size_t n, const uint64_t* k, uint64_t* out, uint32_t* err);

/* BigUintDivRemGenerator::run_once (gadgets/biguint.rs:508-518), BITS = 29.
 * a: [a_limbs][stride], b: [b_limbs][stride], a_limbs <= 18, 1 <= b_limbs <= 18;
 * out: [div_limbs + b_limbs][stride] = div, rem with div_limbs = (b_limbs > a_limbs + 1) ? 0 :
 * a_limbs - b_limbs + 1 (biguint.rs:391-399). */
int p2e_biguint_div_rem_witness_dev(size_t n, const uint64_t* a, int a_limbs, const uint64_t* b,
                                    int b_limbs, size_t in_stride, uint64_t* out,
                                    size_t out_stride, uint32_t* err, void* stream);
int p2e_biguint_div_rem_witness(size_t n, const uint64_t* a, int a_limbs, const uint64_t* b,
                                int b_limbs, uint64_t* out, uint32_t* err);

/* ---- constraint evaluation ------------------------------------------------------------ */

/* MulNonnativeGate::eval_unfiltered over a batch of base-field points, i.e. plonky2's
 * eval_unfiltered_base_batch for this gate (gates/mul_nonnative.rs:101-130).
 * wires: [53][stride] (x 0..8, y 9..17, r 18..26, q 27..35, check_sum 36..52); out: [17][stride]. */
int p2e_eval_mul_nonnative_dev(int modulus, size_t n_rows, const uint64_t* wires, size_t stride,
                               uint64_t* out, size_t out_stride, void* stream);
int p2e_eval_mul_nonnative(int modulus, size_t n_rows, const uint64_t* wires, uint64_t* out);

/* CheckSumGate::eval_unfiltered (gates/mul_nonnative.rs:411-427).
 * wires: [33][stride] (a 0..16, b 17..32); out: [17][stride]. */
int p2e_eval_check_sum_dev(size_t n_rows, const uint64_t* wires, size_t stride, uint64_t* out,
                           size_t out_stride, void* stream);
int p2e_eval_check_sum(size_t n_rows, const uint64_t* wires, uint64_t* out);

/* ---- gadget-chain witnesses ("tapes") ---------------------------------------------------- */
/* A tape is the ordered list of in-tree generator instances a gadget creates.  Generic entry
 * points over the built-in tapes; every input is a [9][stride] wire matrix (a BoolTarget input
 * uses wire 0 of its matrix), `inputs` is an array of p2e_tape_inputs(tape) pointers. */
#define P2E_TAPE_ECDSA 0          /* verify_secp256k1_message_circuit  gadgets/ecdsa.rs:30-53 : msg, r, s, pk.x, pk.y */
#define P2E_TAPE_GLV_MUL 1        /* glv_mul(p, k)                     gadgets/glv.rs:87-104  : k, p.x, p.y */
#define P2E_TAPE_CURVE_ADD 2      /* curve_add(p1, p2)                 gadgets/curve.rs:202-223: x1, y1, x2, y2 */
#define P2E_TAPE_CURVE_DOUBLE 3   /* curve_double(p)                   gadgets/curve.rs:160-185: x, y */
#define P2E_TAPE_CURVE_COND_ADD 4 /* curve_conditional_add(p1, p2, b)  gadgets/curve.rs:225-243: x1, y1, x2, y2, b */
#define P2E_TAPE_FIXED_BASE_MUL 5 /* fixed_base_curve_mul_circuit(G, s) gadgets/curve_fixed_base.rs:18-66: s */
#define P2E_TAPE_CURVE_MSM 6      /* curve_msm_circuit(p, q, n, m)     gadgets/curve_msm.rs:21-79: p.x, p.y, q.x, q.y, n, m */
size_t p2e_tape_slots(int tape);
int p2e_tape_inputs(int tape);
size_t p2e_tape_layout(int tape, uint32_t* kinds, uint32_t* offsets);
/* The plonky2_ux rows UNDER the in-tree generators: range checks.  Every range_check_ux_circuit call of the crate sits in
 * gadgets/nonnative.rs and each is ONE gate row: add_many_nonnative (:330-331: the sum's limbs; the overflow as a single
 * input), sub_nonnative (:373: the difference's limbs), mul_nonnative (:454-460: x, y, r, q, r again -- one 29-bit input
 * per limb the target has -- then b[0..8] and b[8..16] of its CheckSum row as 34-bit inputs).  The call list and its
 * inputs are in-tree (pinned); the gate-row layout is the restated plonky2_ux one (inputs first, then
 * ceil(bits / limb_bits) aux limbs of limb_bits bits per input: DESIGN.md section 9, PARITY UNPINNED -- limb_bits is an
 * argument).
 *   p2e_tape_num_mul / p2e_tape_mul_info: the mul_nonnative instances of a tape in gadget order (ECDSA: 1 087): limbs of
 *       the two factors (a constant's leading zero limbs are dropped, biguint.rs:165-175), first slot of the MulNonnative
 *       tuple, modulus.
 *   p2e_tape_num_range_checks: range_check_ux_circuit calls of the tape, gadget order (ECDSA: 9 168 per signature).
 *   p2e_tape_range_check_info: for calls [first, first + count): index of the generator that makes the call (into
 *       p2e_tape_layout), number of inputs, bits per input.
 *   p2e_tape_range_check_layout: wire_offsets[c] = first wire of call c's row when the rows of all calls are laid end to
 *       end ([num + 1] entries; inputs (1 + ceil(bits / limb_bits)) wires per call).
 *   p2e_tape_range_check_rows_dev: rows[(wire_offsets[c] - wire_offsets[c0] + w) * rs + i] for calls [c0, c1) of every
 *       signature i, read from the finished device witness -- the plain u64 matrix (witness) OR the compact narrow matrix
 *       (narrow), stride ws -- and the tape's inputs.  err[i]: P2E_ERR_LIMB_RANGE when a value does not fit its aux limbs
 *       (the row is unsatisfiable).  At most 65 535 calls per launch. */
int p2e_tape_num_mul(int tape);
int p2e_tape_mul_info(int tape, int j, int* x_limbs, int* y_limbs, uint32_t* out_slot, int* modulus);
/* host-side resolution of instance j's two factors for ONE signature (loads only; inputs: [n_inputs] x 9 wires, witness:
 * [slots] plain values of that signature) -> 9 limbs each, zero-padded: what the rows of calls x and y hold */
int p2e_tape_mul_factors_host(int tape, int j, const uint64_t* const* inputs, const uint64_t* witness, uint64_t* x9, uint64_t* y9);
int p2e_tape_num_range_checks(int tape);
int p2e_tape_range_check_info(int tape, int first, int count, int32_t* generator, int32_t* inputs, int32_t* bits);
int p2e_tape_range_check_layout(int tape, int limb_bits, uint64_t* wire_offsets);
int p2e_tape_range_check_rows_dev(int tape, size_t n, const uint64_t* const* inputs, size_t is, const uint64_t* witness,
                                  const uint32_t* narrow, size_t ws, int c0, int c1, int limb_bits, uint64_t* rows, size_t rs,
                                  uint32_t* err, void* stream);
int p2e_tape_witness_dev(int tape, size_t n, const uint64_t* const* inputs, size_t in_stride,
                         uint64_t* out, size_t out_stride, uint32_t* err, void* stream);
int p2e_tape_witness(int tape, size_t n, const uint64_t* const* inputs, uint64_t* out, uint32_t* err);

/* Builder: the reference's CircuitBuilder extension traits (gadgets/nonnative.rs:53-164,
 * gadgets/curve.rs:34-94, curve_fixed_base.rs, curve_msm.rs, glv.rs) as a handle-based C-API, so a
 * host shim can mirror ANY composition of these gadgets and run its witness as a tape.  Value / point
 * handles are small non-negative ints; every call returns a handle (or 0) on success and a negative
 * number on error.  p2e_builder_finish() registers the tape and returns its id (>= 16) for
 * p2e_tape_*().  Inputs are [9][n] wire matrices in p2e_builder_input() order.
 * Generators that consume coordinates of curve-op results (and points made of such values) are
 * scheduled in a later dependency round automatically.
 * Limits of this version: one fixed_base_curve_mul and one curve_msm / glv_mul per tape; BoolTargets
 * must be inputs. */
int p2e_builder_new(void);
int p2e_builder_free(int builder);
int p2e_builder_input(int builder);                                            /* add_virtual_nonnative_target */
int p2e_builder_constant_nonnative(int builder, const uint32_t* limbs9);       /* nonnative.rs:205-208 */
int p2e_builder_add_nonnative(int builder, int modulus, int x, int y);         /* nonnative.rs:245-276 */
int p2e_builder_sub_nonnative(int builder, int modulus, int x, int y);         /* nonnative.rs:356-388 */
int p2e_builder_mul_nonnative(int builder, int modulus, int x, int y);         /* nonnative.rs:390-464 */
int p2e_builder_inv_nonnative(int builder, int modulus, int x);                /* nonnative.rs:502-536 */
int p2e_builder_neg_nonnative(int builder, int modulus, int x);                /* nonnative.rs:491-500 */
int p2e_builder_add_many_nonnative(int builder, int modulus, const int* xs, int k); /* nonnative.rs:310-353 */
int p2e_builder_nonnative_conditional_neg(int builder, int modulus, int x, int bool_input); /* nonnative.rs:584-596 */
int p2e_builder_if_nonnative(int builder, int modulus, int bool_input, int x, int y);       /* nonnative.rs:297-308 */
int p2e_builder_connect_nonnative(int builder, int x, int y);                  /* nonnative.rs:215-221 */
int p2e_builder_affine_point(int builder, int x, int y);                       /* AffinePointTarget { x, y } */
int p2e_builder_constant_affine_point(int builder, const uint32_t* x9, const uint32_t* y9); /* curve.rs:99-105 */
int p2e_builder_point_x(int builder, int point);
int p2e_builder_point_y(int builder, int point);
int p2e_builder_curve_assert_valid(int builder, int point);                    /* curve.rs:123-135 */
int p2e_builder_curve_add(int builder, int p1, int p2);                        /* curve.rs:202-223 */
int p2e_builder_curve_double(int builder, int p);                              /* curve.rs:160-185 */
int p2e_builder_curve_conditional_add(int builder, int p1, int p2, int bool_input); /* curve.rs:225-243 */
int p2e_builder_fixed_base_curve_mul(int builder, int scalar);                 /* curve_fixed_base.rs:18-66 (base = G) */
int p2e_builder_curve_msm(int builder, int p, int q, int n, int m);            /* curve_msm.rs:21-79 */
int p2e_builder_glv_mul(int builder, int p, int k);                            /* glv.rs:87-104 */
int p2e_builder_finish(int builder);

/* Compact host format (cuts the PCIe bytes of the host-pointer path by ~38 %): slots that always fit
 * 32 bits (limbs of ADD / ADDMANY / SUB / INV / GLV tuples, overflow flags, r and q of MUL tuples) are
 * delivered as uint32_t in `narrow` [narrow_rows][n]; check_sum slots (full field elements) as uint64_t
 * in `wide` [wide_rows][n]; the 16 carries of a CheckSum tuple (b_j + 2^33 < 2^34,
 * gates/mul_nonnative.rs:527) as 16 narrow rows of low words followed by ONE narrow row whose bits
 * 2j, 2j+1 are the two top bits of carry j.  p2e_tape_compact_layout() fills, for every slot of the
 * plain layout, (1u << 31 | wide row), (1u << 30 | narrow row of the low word; bit 29 marks the first
 * carry of a tuple) or the narrow row; it returns the slot count and writes the two row counts.
 * Same values, same order, different storage width (rows flagged in err[] may hold carries >= 2^34,
 * which this format truncates -- the reference panics on those).
 * Derived form (`wide` == NULL, another 36 % fewer bytes): the check_sum slots are not shipped at all.
 * Column j of a MUL tuple is determined by the carries of the CheckSum tuple that follows it in the tape,
 *   check_sum_j = (b_j - 2^33) * 2^29 - (b_{j-1} - 2^33)   (j = 0: no second term; j = 16: -(b_15 - 2^33))
 * as a field element -- the CheckSumGate's own constraints (gates/mul_nonnative.rs:411-427) set to zero --
 * so the host rebuilds it with two integer operations per slot while it fills the wires.  The device
 * kernels write the compact rows directly (no plain witness is materialised) from the same un-carried columns
 * the carries come from; a row that is not of this form (only rows whose generator already failed) gets
 * P2E_ERR_CARRY_RANGE in err[i]. */
size_t p2e_tape_compact_layout(int tape, uint32_t* map, uint32_t* narrow_slots, uint32_t* wide_slots);
/* device-resident compact witness: narrow [narrow_slots][stride] uint32, wide [wide_slots][stride] uint64 or NULL */
int p2e_tape_witness_compact_dev(int tape, size_t n, const uint64_t* const* inputs, size_t in_stride,
                                 uint32_t* narrow, uint64_t* wide, size_t out_stride, uint32_t* err, void* stream);
int p2e_tape_witness_compact(int tape, size_t n, const uint64_t* const* inputs, uint32_t* narrow,
                             uint64_t* wide, uint32_t* err);
/* Host-side decoder (pure format conversion, no device): plain out[slot][i] from the compact matrices; wide == NULL
 * decodes the derived form (check_sum slots rebuilt from the carries). */
int p2e_tape_compact_expand(int tape, size_t n, const uint32_t* narrow, const uint64_t* wide, uint64_t* out);

/* ---- ECDSA verification circuit witness (tape) ----------------------------------------- */

/* Number of u64 slots one verify_secp256k1_message_circuit witness occupies in tape order
 * (gadgets/ecdsa.rs:30-53 and callees): inputs excluded. */
size_t p2e_ecdsa_tape_slots(void);
/* Fills kinds[i] / offsets[i] for generator instance i of the tape; returns the instance count
 * (call with NULLs to query).  kinds: 0 ADD 1 ADDMANY 2 SUB 3 INV 4 MUL 5 CHECKSUM 6 GLV. */
size_t p2e_ecdsa_tape_layout(uint32_t* kinds, uint32_t* offsets);

/* Witness of n independent verify_secp256k1_message_circuit instances.
 * msg, r, s: [9][stride] scalar-field limbs; pkx, pky: [9][stride] base-field limbs.
 * out: [p2e_ecdsa_tape_slots()][out_stride] every in-tree generator's output tuple in gadget
 * call order.  err: [n]. */
int p2e_ecdsa_witness_dev(size_t n, const uint64_t* msg, const uint64_t* r, const uint64_t* s,
                          const uint64_t* pkx, const uint64_t* pky, size_t in_stride,
                          uint64_t* out, size_t out_stride, uint32_t* err, void* stream);
int p2e_ecdsa_witness(size_t n, const uint64_t* msg, const uint64_t* r, const uint64_t* s,
                      const uint64_t* pkx, const uint64_t* pky, uint64_t* out, uint32_t* err);
/* the same witness in the compact device format (p2e_tape_witness_compact_dev on the ECDSA tape; wide may be NULL) */
int p2e_ecdsa_witness_compact_dev(size_t n, const uint64_t* msg, const uint64_t* r, const uint64_t* s,
                                  const uint64_t* pkx, const uint64_t* pky, size_t in_stride, uint32_t* narrow,
                                  uint64_t* wide, size_t out_stride, uint32_t* err, void* stream);

/* glv_mul(p, k) sub-circuit alone (gadgets/glv.rs:87-104): GLV decomposition + MSM chain. */
size_t p2e_glv_mul_tape_slots(void);
size_t p2e_glv_mul_tape_layout(uint32_t* kinds, uint32_t* offsets);
int p2e_glv_mul_witness_dev(size_t n, const uint64_t* k, const uint64_t* px, const uint64_t* py,
                            size_t in_stride, uint64_t* out, size_t out_stride, uint32_t* err,
                            void* stream);
int p2e_glv_mul_witness(size_t n, const uint64_t* k, const uint64_t* px, const uint64_t* py,
                        uint64_t* out, uint32_t* err);

/* ---- plonky2_ux gates (range check, comparison, add-many, subtraction, arithmetic) ------------
 * PARITY UNPINNED: plonky2_ux is an un-vendored, un-pinned git dependency of the reference
 * (Cargo.toml:19); these entry points implement the spec restated in oracle/ux_spec.py (upstream
 * plonky2_u32's gates generalised to `bits`-wide limbs) with every geometry parameter an argument.
 * They stand in for Gate::eval_unfiltered_base_batch / SimpleGenerator::run_once of the gates behind
 *   range_check_ux_circuit  gadgets/nonnative.rs:330-331,373,454-460   (P2E_UX_RANGE_CHECK)
 *   list_le_ux_circuit      gadgets/biguint.rs:222-229                 (P2E_UX_COMPARISON)
 *   add_many_ux             gadgets/biguint.rs:261                     (P2E_UX_ADD_MANY)
 *   sub_ux                  gadgets/biguint.rs:284                     (P2E_UX_SUBTRACTION)
 *   mul_ux                  gadgets/biguint.rs:305,335                 (P2E_UX_ARITHMETIC)
 * Geometry: `bits` = width of a UXTarget (29 / 34 in the reference; comparison: num_bits);
 * `limb_bits` = width of the aux limbs / chunks (1..4); `count` = input limbs per row (range check),
 * ops per row (add-many, subtraction, arithmetic) or 1 (comparison); `num_addends` (add-many only).
 * Wire maps and constraint order: oracle/ux_spec.py.
 *   p2e_ux_gate_shape : num_wires / num_constraints of the gate row, num_inputs of its generator.
 *   p2e_eval_ux       : wires [num_wires][stride] -> out [num_constraints][stride].
 *   p2e_ux_witness    : inputs [num_inputs][stride] -> complete gate rows [num_wires][stride], err [n].
 *     inputs per gate: range check: the limbs; comparison: first, second; add-many: per op the
 *     addends then the carry-in; subtraction: per op x, y, borrow-in; arithmetic: per op m0, m1, addend.
 *     err[i] has P2E_ERR_LIMB_RANGE when a value does not fit its limbs (unsatisfiable row). */
#define P2E_UX_RANGE_CHECK 0
#define P2E_UX_COMPARISON 1
#define P2E_UX_ADD_MANY 2
#define P2E_UX_SUBTRACTION 3
#define P2E_UX_ARITHMETIC 4
int p2e_ux_gate_shape(int gate, int bits, int limb_bits, int count, int num_addends, int* num_wires,
                      int* num_constraints, int* num_inputs);
int p2e_eval_ux_dev(int gate, int bits, int limb_bits, int count, int num_addends, size_t n,
                    const uint64_t* wires, size_t in_stride, uint64_t* out, size_t out_stride, void* stream);
int p2e_eval_ux(int gate, int bits, int limb_bits, int count, int num_addends, size_t n,
                const uint64_t* wires, uint64_t* out);
int p2e_ux_witness_dev(int gate, int bits, int limb_bits, int count, int num_addends, size_t n,
                       const uint64_t* inputs, size_t in_stride, uint64_t* row, size_t out_stride,
                       uint32_t* err, void* stream);
int p2e_ux_witness(int gate, int bits, int limb_bits, int count, int num_addends, size_t n,
                   const uint64_t* inputs, uint64_t* row, uint32_t* err);

/* ---- serialized generator streams + level-scheduled executor -------------
 * A stream is the concatenation of records  u32 tag | SimpleGenerator::serialize() bytes  -- what plonky2's
 * WitnessGeneratorSerializer::write_generator emits for a registry whose indices are the P2E_GEN_* tags
 * below (the FF type parameter of the nonnative generators is not part of serialize(), so it is part of
 * the tag).  Bodies are the reference's own (file:line per tag); primitives are plonky2's
 * util::serialization (out of tree, restated): usize and field elements 8 bytes LE, u32 4 bytes LE, bool
 * 1 byte, Target = bool is_wire then usize row, usize column (wire) or usize index (virtual),
 * target_vec = usize length then targets.
 * Three record kinds are not in-tree generators: P2E_GEN_COPY (two Targets of one copy-constraint
 * partition -- CircuitBuilder::connect), and two plonky2 generators restated so that the bool / select
 * glue of the gadgets can run in the same pass: P2E_GEN_CONSTANT = ConstantGenerator {usize row, usize
 * constant_index, usize wire_index, field constant} and P2E_GEN_ARITHMETIC = ArithmeticBaseGenerator
 * {usize row, field const_0, field const_1, usize i}: wire 4i+3 = const_0 * wire 4i * wire 4i+1 + const_1 *
 * wire 4i+2.
 * p2e_genstream_load parses a stream, joins copy partitions into SLOTS (rows of the value table), orders the
 * generators into dependency levels (dependencies() = the targets a record reads) and returns a handle.
 * p2e_genstream_run* executes every generator once, level by level, for n independent circuits:
 * inputs = values of the slots no generator produces (p2e_genstream_input_slots), table [slots][stride].
 * A slot written twice must receive equal values (plonky2: "partition set twice with different values"),
 * otherwise err[i] gets P2E_ERR_CONNECT; generator panics set the usual P2E_ERR_* bits. */
#define P2E_GEN_MUL_NONNATIVE_BASE 0              /* MulNonnativeGenerator<F,D,Secp256K1Base>   gates/mul_nonnative.rs:326-330 */
#define P2E_GEN_MUL_NONNATIVE_SCALAR 1            /* MulNonnativeGenerator<F,D,Secp256K1Scalar>                               */
#define P2E_GEN_CHECK_SUM 2                       /* CheckSumGenerator                          gates/mul_nonnative.rs:533-536 */
#define P2E_GEN_NONNATIVE_ADDITION_BASE 3         /* NonNativeAdditionGenerator                 gadgets/nonnative.rs:647-652   */
#define P2E_GEN_NONNATIVE_ADDITION_SCALAR 4
#define P2E_GEN_NONNATIVE_MULTIPLE_ADDS_BASE 5    /* NonNativeMultipleAddsGenerator             gadgets/nonnative.rs:730-740   */
#define P2E_GEN_NONNATIVE_MULTIPLE_ADDS_SCALAR 6
#define P2E_GEN_NONNATIVE_SUBTRACTION_BASE 7      /* NonNativeSubtractionGenerator              gadgets/nonnative.rs:812-817   */
#define P2E_GEN_NONNATIVE_SUBTRACTION_SCALAR 8
#define P2E_GEN_NONNATIVE_INVERSE_BASE 9          /* NonNativeInverseGenerator                  gadgets/nonnative.rs:874-878   */
#define P2E_GEN_NONNATIVE_INVERSE_SCALAR 10
#define P2E_GEN_BIGUINT_DIV_REM 11                /* BigUintDivRemGenerator                     gadgets/biguint.rs:520-526     */
#define P2E_GEN_GLV_DECOMPOSITION 12              /* GLVDecompositionGenerator                  gadgets/glv.rs:144-150         */
#define P2E_GEN_COPY 13
#define P2E_GEN_CONSTANT 14
#define P2E_GEN_ARITHMETIC 15
/* keccak(0u64) as a secp256k1 scalar -- the blinding constant of curve_msm_circuit / fixed_base_curve_mul_circuit
 * (gadgets/curve_msm.rs:33-39), 4 little-endian 64-bit words; host-only. */
int p2e_blinding_scalar(uint64_t* words4);

/* plonky2 glue generators of the gadgets' scalar handling (restated, parity unpinned against the dependency):
 *   P2E_GEN_BASE_SPLIT     BaseSplitGenerator<2> { usize row, usize num_limbs }: split_le_base::<2> (gadgets/split_nonnative.rs:29-37)
 *   P2E_GEN_RANDOM_ACCESS  RandomAccessGenerator { usize row, usize bits, usize num_copies, usize num_extra_constants,
 *                          usize copy }: random_access (gadgets/curve_windowed_mul.rs:97-104)
 *   P2E_GEN_EQUALITY       EqualityGenerator { Target x, Target y, Target equal, Target inv }: is_equal
 *                          (gadgets/curve_msm.rs:69, gadgets/curve_fixed_base.rs:57) */
#define P2E_GEN_BASE_SPLIT 16
#define P2E_GEN_RANDOM_ACCESS 17
#define P2E_GEN_EQUALITY 18
/* { Target }: a target the host sets before generation (plonky2 PartialWitness / public input).  Needed when a gadget
 * connects a generator output to an input (connect_nonnative(r, point.x), gadgets/ecdsa.rs:48-52): the writer is then checked
 * against the input (P2E_ERR_CONNECT) instead of producing it. */
#define P2E_GEN_INPUT 19
int p2e_genstream_load(const uint8_t* bytes, size_t len);          /* handle >= 0, or < 0 (p2e_last_error) */
int p2e_genstream_free(int handle);
long p2e_genstream_info(int handle, int what); /* 0 slots, 1 generators, 2 levels, 3 launch groups, 4 input slots, 5 targets */
long p2e_genstream_slot(int handle, int is_wire, uint64_t row_or_index, uint64_t column);
long p2e_genstream_input_slots(int handle, uint32_t* out, size_t cap);
long p2e_genstream_generator_levels(int handle, int32_t* out, size_t cap);   /* level of generator k, in stream order */
int p2e_genstream_run_dev(int handle, size_t n, const uint32_t* in_slots /* host */, int n_in,
                          const uint64_t* in_values /* device [n_in][in_stride] */, size_t in_stride,
                          uint64_t* table /* device [slots][stride] */, size_t stride, uint32_t* err, void* stream);
int p2e_genstream_run(int handle, size_t n, const uint32_t* in_slots, int n_in, const uint64_t* in_values,
                      const uint32_t* out_slots, int n_out, uint64_t* out_values, uint32_t* err);

/* ---- selector filtering + alpha-combination around the gate evaluators ----------
 * The step after Gate::eval_unfiltered_base_batch in plonky2's quotient stage (out of tree, restated in
 * oracle/vanishing_spec.py): Gate::eval_filtered_base_batch multiplies a row's constraints by
 *   filter(s) = prod_{j in [group_start, group_end), j != gate_index} (j - s) * (many_selectors ? (2^32-1 - s) : 1)
 * (s = the row's selector constant; many_selectors = the circuit has more than one selector polynomial),
 * evaluate_gate_constraints_base_batch adds the gates' vectors position by position, and
 * eval_vanishing_poly_base_batch folds all terms with reduce_with_powers_multi (term t times alpha^t).
 * One call does this for one gate type over n LDE rows:
 *   out[a][row] (+)= filter(row) * sum_c constraint_c(row) * alphas[a]^(first_power + c),  a < n_alphas <= 4
 * (`accumulate` != 0 adds to what `out` holds, so the gates of a circuit are summed call by call;
 * first_power = number of vanishing terms that precede the gate constraints; selector == NULL: filter = 1).
 * gate / params (int[5]): P2E_GATE_MUL_NONNATIVE {modulus}, P2E_GATE_CHECK_SUM {}, P2E_GATE_UX {ux gate,
 * bits, limb_bits, count, num_addends}.  The two in-tree gates are fused (wires are read once, only
 * n_alphas columns are written), and so is the plonky2_ux range-check gate (scratch: only its table of challenge
 * powers); the other P2E_GATE_UX gates go through p2e_eval_ux_dev into scratch and a generic fold kernel.
 * scratch: p2e_eval_filtered_scratch_bytes() bytes of device memory (NULL is fine for the two in-tree gates). */
#define P2E_GATE_MUL_NONNATIVE 0
#define P2E_GATE_CHECK_SUM 1
#define P2E_GATE_UX 2
int p2e_gate_shape(int gate, const int* params, int* num_wires, int* num_constraints);
size_t p2e_eval_filtered_scratch_bytes(int gate, const int* params, size_t n);
int p2e_eval_filtered_combined_dev(int gate, const int* params, int gate_index, int group_start, int group_end,
                                   int many_selectors, size_t n, const uint64_t* wires, size_t in_stride,
                                   const uint64_t* selector, const uint64_t* alphas, int n_alphas,
                                   uint64_t first_power, int accumulate, uint64_t* out, size_t out_stride,
                                   uint64_t* scratch, void* stream);
int p2e_eval_filtered_combined(int gate, const int* params, int gate_index, int group_start, int group_end,
                               int many_selectors, size_t n, const uint64_t* wires, const uint64_t* selector,
                               const uint64_t* alphas, int n_alphas, uint64_t first_power, int accumulate,
                               uint64_t* out);

/* ---- Goldilocks NTT / coset LDE (the step that feeds the gate evaluators) ---------
 * plonky2 conventions (field crate, out of tree, restated): omega_n = POWER_OF_TWO_GENERATOR^(2^(32 - log_n)),
 * POWER_OF_TWO_GENERATOR = 7^((p-1)/2^32) = 1753635133440165772; natural order in and out,
 * X[k] = sum_i x[i] omega_n^(i k); the inverse divides by n.  p2e_coset_lde is
 * PolynomialValues::lde_onto_coset: ifft, zero-pad by 2^rate_bits, coset fft with shift g = 7:
 * out[j] = P(g * omega_(n << rate_bits)^j), wire-major [poly][lde_row] -- the matrix p2e_eval_* reads.
 * Matrices are [n_polys][stride]; log_n 2..24; scratch: p2e_ntt_scratch_bytes() bytes of device memory
 * (rate_bits = 0 for p2e_ntt_dev).  in and out may not alias. */
size_t p2e_ntt_scratch_bytes(int log_n, int rate_bits, int n_polys);
int p2e_ntt_dev(int log_n, int inverse, int n_polys, const uint64_t* in, size_t in_stride, uint64_t* out,
                size_t out_stride, uint64_t* scratch, void* stream);
int p2e_ntt(int log_n, int inverse, int n_polys, const uint64_t* in, uint64_t* out);
int p2e_coset_lde_dev(int log_n, int rate_bits, int n_polys, const uint64_t* values, size_t in_stride,
                      uint64_t* out, size_t out_stride, uint64_t* scratch, void* stream);
int p2e_coset_lde(int log_n, int rate_bits, int n_polys, const uint64_t* values, uint64_t* out);

/* ---- measurement helpers ---------------------------------------------------------------- */
/* Per-kernel timing of the chain entry points: when enabled, CUDA events are recorded on the
 * launching stream between the kernels of the NEXT chain call; p2e_last_kernel_times() then
 * synchronises and returns their durations in ms, in launch order (k_prep, k_chain_replay,
 * k_chain_replay(join), k_chain_invert, k_chain_affine, k_chain_expand, k_chain_checks).
 * Returns the number of kernels written (<= cap), or a negative error. */
int p2e_set_profiling(int on);
int p2e_last_kernel_times(float* ms, int cap);
/* Timeline of pipelined chain calls: with p2e_set_option("chain_trace", 1) every chain call records a timing event
 * after each of its kernels on the internal stream that ran it.  p2e_chain_trace_dump synchronises the device and
 * returns (call index, kernel id, ms since the first event) triples, then clears the trace.  Kernel ids: 0 call start,
 * 1 k_prep done, 2 k_chain_replay done, 3 join done, 4 k_chain_invert_* done, 5 full-GPU phase may start,
 * 6 k_chain_affine done, 7 k_chain_expand done. */
int p2e_chain_trace_dump(int* calls, int* kernels, float* ms, int cap);
/* Dependent-free IMAD / IMAD.WIDE throughput microbenchmark on the current device; returns
 * giga-ops per second (one op = one 32-bit lane multiply-add), or a negative error. */
double p2e_bench_imad(int wide, int iters);
/* Test hook: runs one operation of the saturated 8 x 32-bit field (csrc/p2e_f8.cuh, PTX carry-chain bodies) on the
 * device for n operand pairs; a, b, out: [n][9] uint32, 8 words used.  op: 0 mul, 1 sqr, 2 add, 3 sub, 4..7 times
 * 2 / 3 / 4 / 8, 8 canonical form, 11 inverse, 12 / 13 remainder / quotient of a * b by p.  tests/test_gpu_f8.py compares with the portable bodies. */
int p2e_debug_f8(int op, size_t n, const uint32_t* a, const uint32_t* b, uint32_t* out);

#ifdef __cplusplus
}
#endif
#endif /* P2E_H */
