/* p2gpu.h — C ABI of libp2gpu.so, the B200 (sm_100a) backend for the hot path of plonky2's
 * CircuitData::prove as driven by the 0xPARC/plonky2-aes gadget crates.
 *
 * What it replaces.  The reference enters the path through one call, `data.prove(pw)`
 * (/root/reference/aes-gcm/examples/aes_gcm_128.rs:52, aes-gcm/src/circuit_gcm.rs:781,
 * aes-gcm/src/circuit_aes.rs:404-725, feistel/src/circuit.rs:151, ecgfp5/src/circuit.rs:97,
 * poseidon-cipher/src/circuit.rs:188), which lands in the un-vendored dependency
 * plonky2 @ 109d517 (/root/reference/Cargo.toml:12).  A maintainer patches that crate
 * ([patch] on Cargo.toml:12) so that plonk::prover::prove_with_partition_witness calls
 * p2g_prove(); INTEGRATION.md shows the Rust `extern "C"` block and the patch.
 *
 * Conventions: every function returns 0 on success or a negative P2G_E_* code; field elements
 * are canonical little-endian u64 (< p = 2^64 - 2^32 + 1); the caller owns every host buffer;
 * the library owns all device memory behind opaque handles; one ctx per device, used from one
 * host thread at a time.  Nothing here falls back to the CPU: without a CUDA device
 * p2g_ctx_create fails with P2G_E_CUDA.
 */
#ifndef P2GPU_H
#define P2GPU_H
#include <stdint.h>
#include <stddef.h>
#include "p2witness.h"   /* p2w_program_desc: the generator program p2g_wprog_load takes */

#ifdef __cplusplus
extern "C" {
#endif

#define P2G_OK 0
#define P2G_E_CUDA (-1)    /* CUDA runtime error; text via p2g_last_error */
#define P2G_E_BADARG (-2)
#define P2G_E_UNSAT (-3)   /* the FRI final polynomial has non-zero high coefficients (an internal consistency
                              check).  NOTE: a wire matrix that violates a gate or copy constraint is NOT detected
                              here: with quotient_degree_factor = 2^rate_bits the quotient is interpolated exactly
                              on the 8n points, so p2g_prove returns P2G_OK and a proof the verifier rejects --
                              the behaviour of upstream release builds, where the check is a debug assertion.  The
                              reference's `prove(..).is_err()` on a bad witness
                              (/root/reference/aes-gcm/src/circuit_aes.rs:403-405) comes from witness generation
                              (conflicting partition values), which stays on the host: p2w_generate* returns
                              P2W_E_CONFLICT (include/p2witness.h). */
#define P2G_E_POW (-4)
#define P2G_E_NOMEM (-5)

typedef struct p2g_ctx p2g_ctx;
typedef struct p2g_batch p2g_batch;     /* device-resident PolynomialBatch (fri/oracle.rs) */
typedef struct p2g_circuit p2g_circuit; /* device-resident prover data of one CircuitData */

int32_t p2g_version(void);
int32_t p2g_ctx_create(int32_t device, p2g_ctx** out);
void p2g_ctx_destroy(p2g_ctx* ctx);
const char* p2g_last_error(p2g_ctx* ctx);
int32_t p2g_ctx_sync(p2g_ctx* ctx);
/* the CUDA stream (cudaStream_t) every call on this ctx is ordered on — for event timing */
void* p2g_ctx_stream(p2g_ctx* ctx);

/* ---- PolynomialBatch::from_values / from_coeffs  (fri/oracle.rs) ------------------------------
 * cols: ncols polynomials of n = 2^log_n words each, column-major ([ncols][n]).
 * `*_host` read host memory (H2D inside the call); `*_dev` read device memory already in HBM.
 * cap_out (host, 4 * 2^cap_height words) receives the Merkle cap; may be NULL. */
int32_t p2g_commit_from_values(p2g_ctx* ctx, const uint64_t* cols_host, uint32_t ncols, uint32_t log_n,
                               uint32_t rate_bits, uint32_t cap_height, p2g_batch** out, uint64_t* cap_out);
int32_t p2g_commit_from_coeffs(p2g_ctx* ctx, const uint64_t* cols_host, uint32_t ncols, uint32_t log_n,
                               uint32_t rate_bits, uint32_t cap_height, p2g_batch** out, uint64_t* cap_out);
int32_t p2g_commit_from_values_dev(p2g_ctx* ctx, const uint64_t* cols_dev, uint32_t ncols, uint32_t log_n,
                                   uint32_t rate_bits, uint32_t cap_height, p2g_batch** out, uint64_t* cap_out);
int32_t p2g_commit_from_coeffs_dev(p2g_ctx* ctx, const uint64_t* cols_dev, uint32_t ncols, uint32_t log_n,
                                   uint32_t rate_bits, uint32_t cap_height, p2g_batch** out, uint64_t* cap_out);
/* Multi-GPU split of ONE commitment (coset sharding): this rank extends and hashes only the leaf blocks
 * [blk_first, blk_first + blk_count) of the 2^rate_bits cosets (block b = coset bitrev(b), a contiguous run of n
 * Merkle leaves).  cap_part_out receives this shard's 2^cap_height * blk_count / 2^rate_bits cap entries; the
 * ranks all-gather them (NCCL) in block order to obtain MerkleTree::new(...).cap.  blk_count: power of two. */
int32_t p2g_commit_blocks_from_values_dev(p2g_ctx* ctx, const uint64_t* cols_dev, uint32_t ncols, uint32_t log_n,
                                          uint32_t rate_bits, uint32_t cap_height, uint32_t blk_first, uint32_t blk_count,
                                          p2g_batch** out, uint64_t* cap_part_out);
int32_t p2g_batch_free(p2g_ctx* ctx, p2g_batch* b);
/* read-back (parity tests, query phase). */
int32_t p2g_batch_get_coeffs(p2g_ctx* ctx, const p2g_batch* b, uint64_t* out /*[ncols][n]*/);
/* LDE values, column-major, bit-reversed index order: out[c*N + j] = f_c(7 * w_N^bitrev(j)) */
int32_t p2g_batch_get_lde(p2g_ctx* ctx, const p2g_batch* b, uint64_t* out /*[ncols][N]*/);
/* digests of tree level `level` (0 = leaf digests); level == path_len returns the cap */
int32_t p2g_batch_get_level(p2g_ctx* ctx, const p2g_batch* b, uint32_t level, uint64_t* out);
/* MerkleTree::get(leaf) + MerkleTree::prove(leaf): row (ncols words) and siblings (path_len*4) */
int32_t p2g_batch_open_leaf(p2g_ctx* ctx, const p2g_batch* b, uint64_t leaf_index, uint64_t* row_out,
                            uint64_t* siblings_out);

/* ---- stand-alone MerkleTree::new over row-major leaves (hash/merkle_tree.rs) ------------------ */
int32_t p2g_merkle_cap(p2g_ctx* ctx, const uint64_t* leaves_host, uint32_t log_leaves, uint32_t leaf_len,
                       uint32_t cap_height, uint64_t* cap_out, uint64_t* leaf_digests_out /* may be NULL */);
/* hash_n_to_m_no_pad on `count` independent inputs of `len` words each (row-major) -> 4 words each */
int32_t p2g_hash_no_pad_many(p2g_ctx* ctx, const uint64_t* in_host, uint32_t count, uint32_t len, uint64_t* out_host);

/* ---- circuit description: what CircuitBuilder::build leaves in ProverOnlyCircuitData +
 *      CommonCircuitData (plonk/circuit_data.rs) and the quotient kernel needs ------------------ */
enum { P2G_GATE_NOOP = 0, P2G_GATE_CONSTANT = 1, P2G_GATE_PUBLIC_INPUT = 2, P2G_GATE_ARITHMETIC = 3,
       P2G_GATE_LOOKUP = 4, P2G_GATE_LOOKUP_TABLE = 5, P2G_GATE_POSEIDON = 6 };
typedef struct {
    int32_t kind;
    int32_t selector_index;
    int32_t group_start, group_end;
    int32_t num_constraints;
    int32_t param0;
} p2g_gate;
typedef struct {
    int32_t degree_bits;
    int32_t num_wires, num_routed_wires, num_constants;
    int32_t num_challenges, quotient_degree_factor;
    int32_t rate_bits, cap_height, pow_bits, num_query_rounds;
    int32_t num_reduction_arity_bits; int32_t reduction_arity_bits[16];
    int32_t num_selectors, num_lookup_selectors;
    int32_t num_gates; const p2g_gate* gates;
    int32_t num_gate_constraints;
    int32_t num_partial_products;
    int32_t num_luts;
    const int32_t* lut_lens;
    const uint16_t* lut_data;      /* (inp,out) pairs, all LUTs concatenated */
    const int32_t* lookup_rows;    /* [num_luts][3]: last_lu_gate, last_lut_gate, first_lut_gate */
    int32_t num_public_inputs;
    const uint64_t* k_is;          /* [num_routed_wires] */
    const uint64_t* constants_sigmas; /* [(selectors+lookup selectors+constants+routed)][n] values, host */
    uint64_t circuit_digest[4];
} p2g_circuit_desc;

int32_t p2g_circuit_load(p2g_ctx* ctx, const p2g_circuit_desc* desc, p2g_circuit** out,
                         uint64_t* constants_sigmas_cap_out /* may be NULL */);
int32_t p2g_circuit_free(p2g_ctx* ctx, p2g_circuit* c);
/* number of u64 words of a serialised proof for this circuit */
size_t p2g_proof_words(const p2g_circuit* c);

/* ---- ProofWithPublicInputs::to_bytes / from_bytes (plonky2 util/serialization/mod.rs) -----------------
 * Converts between the flat proof words p2g_prove returns (layout: DESIGN.md section 5) and upstream's byte
 * format, so the Rust side finishes with `ProofWithPublicInputs::from_bytes(bytes, common_data)`:
 * little-endian u64 field elements in upstream's write order (caps; openings with the two lookup vectors
 * after plonk_zs_next; FRI caps; per query the initial-tree rows and steps, every Merkle path preceded by its
 * length as one u8; final polynomial; pow witness; the public-input count as u64, then the public inputs).
 * Host-only: needs no context and no device.  from_bytes rejects non-canonical elements, wrong path lengths
 * and trailing bytes with P2G_E_BADARG. */
size_t p2g_proof_bytes_len(const p2g_circuit_desc* desc);
int32_t p2g_proof_to_bytes(const p2g_circuit_desc* desc, const uint64_t* words, size_t nwords, uint8_t* out, size_t cap_bytes,
                           size_t* len_out);
int32_t p2g_proof_from_bytes(const p2g_circuit_desc* desc, const uint8_t* bytes, size_t len, uint64_t* words_out,
                             size_t cap_words, size_t* words_len);

/* ---- the whole hot path: prove_with_partition_witness (plonk/prover.rs) ----------------------
 * wires: [num_wires][n] host, column-major full witness; public_inputs: num_public_inputs words.
 * proof_out: flat u64 proof (layout in DESIGN.md, identical to the oracle's). */
int32_t p2g_prove(p2g_ctx* ctx, const p2g_circuit* c, const uint64_t* wires_host, const uint64_t* public_inputs,
                  uint64_t* proof_out, size_t proof_cap_words, size_t* proof_words_out);
/* ---- stages of the hot path on their own --------------------------------------------
 * p2g_quotient = compute_quotient_polys + the commitment of the quotient chunks (plonk/prover.rs): wires / zs are whole
 * batches of this circuit (p2g_commit_from_values*: 135 wire columns; Z, partial products and lookup polynomials in the
 * prover's column order), challenges as plonky2's get_n_challenges returned them: betas / gammas / alphas
 * [num_challenges], deltas [num_challenges][4] = (a, b, alpha, delta) per challenge (NULL without lookups).
 * p2g_open = OpeningSet::new: f(zeta) of every polynomial of the given batches, in batch and column order, as (c0, c1). */
int32_t p2g_quotient(p2g_ctx* ctx, const p2g_circuit* c, const p2g_batch* wires, const p2g_batch* zs, const uint64_t* public_inputs,
                     const uint64_t* betas, const uint64_t* gammas, const uint64_t* deltas, const uint64_t* alphas,
                     p2g_batch** quotient_out, uint64_t* cap_out);
int32_t p2g_open(p2g_ctx* ctx, const p2g_batch* const* batches, uint32_t n_batches, const uint64_t zeta[2], uint64_t* openings_out);
/* p2g_fri_prove = PolynomialBatch::prove_openings (fri/oracle.rs) + fri_proof (fri/prover.rs) on their own: batch
 * combination of all committed polynomials at zeta and g * zeta, LDE, commit phase (caps, folding challenges), final
 * polynomial, proof of work (lowest nonce) and the query rounds.  wires / zs / quotient are whole batches of this circuit
 * (the preprocessed batch is the circuit's own), zeta the point the caller drew after the quotient cap, challenger_io the
 * state of the caller's Challenger AFTER it observed the openings, 30 words: [0..12) sponge state, [12..20) input buffer,
 * [20..28) output buffer, [28] input length, [29] output length; on return it holds the state after the proof-of-work
 * response and the query indices were drawn.  fri_out receives the FriProof part of the flat proof layout (DESIGN.md
 * section 5: commit-phase caps, query rounds, final polynomial, pow witness), p2g_fri_proof_words(c) words. */
size_t p2g_fri_proof_words(const p2g_circuit* c);
int32_t p2g_fri_prove(p2g_ctx* ctx, const p2g_circuit* c, const p2g_batch* wires, const p2g_batch* zs, const p2g_batch* quotient,
                      const uint64_t zeta[2], uint64_t* challenger_io, uint64_t* fri_out, size_t fri_cap_words, size_t* fri_words_out);
/* Batch of independent proofs (BASELINE config 5): proof i is proved on context i mod n_ctx, one host thread per
 * context inside the call; contexts may sit on one GPU (several proofs in flight) or on several.  circuits[t] must
 * have been loaded on ctxs[t].  status_out[i] receives each proof's return code; returns the first failure. */
int32_t p2g_prove_batch(p2g_ctx* const* ctxs, const p2g_circuit* const* circuits, uint32_t n_ctx,
                        const uint64_t* const* wires_host, const uint64_t* const* public_inputs /* may be NULL */, uint32_t n_proofs,
                        uint64_t* const* proofs_out, size_t proof_cap_words, int32_t* status_out);
/* same with the witness already in HBM */
int32_t p2g_prove_dev(p2g_ctx* ctx, const p2g_circuit* c, const uint64_t* wires_dev, const uint64_t* public_inputs,
                      uint64_t* proof_out, size_t proof_cap_words, size_t* proof_words_out);

/* ---- ONE proof split over `world` GPUs by coset (SURVEY.md section 8(e) split 2; BASELINE north_star: "column-sharded
 * LDE / Merkle leaves within one large proof, with an NCCL all-gather over NVLink of leaf digests") -------------------
 * One process per GPU.  Rank r extends, hashes and evaluates only the 2^rate_bits / world cosets it owns (contiguous
 * Merkle leaf blocks); the quotient at a point reads rows i and i + 8, which lie in the same coset, so the only data
 * exchanged are: the cap entries of every commitment (wires, Z, quotient, each FRI layer), the per-coset interpolants
 * of the two quotient columns (16 N / world bytes), the last FRI layer and the query records.  Each exchange is ONE
 * all-gather the caller provides: `exchange(user, stage, bytes)` must gather `bytes` bytes from every rank's `send_dev`
 * into every rank's `recv_dev` ([world][bytes], rank order) and return 0 once recv_dev is complete (e.g.
 * ncclAllGather + stream synchronize; tests on one GPU run the ranks as host threads and copy between their buffers).
 * send_dev: p2g_shard_buffer_bytes(c, world) bytes, recv_dev: world times that, both DEVICE memory owned by the caller.
 * Every rank passes the same wires and returns the same proof, bit-identical to p2g_prove's. */
typedef int32_t (*p2g_exchange_fn)(void* user, int32_t stage, uint64_t bytes_per_rank);
size_t p2g_shard_buffer_bytes(const p2g_circuit* c, uint32_t world);
int32_t p2g_prove_sharded(p2g_ctx* ctx, const p2g_circuit* c, const uint64_t* wires_host, const uint64_t* public_inputs,
                          uint32_t rank, uint32_t world, uint64_t* send_dev, uint64_t* recv_dev, size_t buf_bytes,
                          p2g_exchange_fn exchange, void* user, uint64_t* proof_out, size_t proof_cap_words, size_t* proof_words_out);

/* stage read-backs for parity tests (valid after a p2g_prove on this ctx) */
typedef struct {
    uint64_t betas[4], gammas[4], deltas[16], alphas[4];
    uint64_t zeta[2], fri_alpha[2], fri_betas[32];
    uint64_t pow_witness;
    uint64_t query_indices[64];
} p2g_transcript;
/* Device-side PartitionWitness::full_witness (plonky2 iop/witness.rs) in front of the prover.
 * The host keeps one value per copy-constraint partition ("slot"); `wire_map[col*n + row]` is the slot
 * of that wire cell (or -1 for an empty cell) and (fixed_pos, fixed_val) are the constant cells.
 * p2g_prove_slots uploads the slot values (a few MB instead of the num_wires * n matrix), gathers the
 * wire matrix on the device and runs the same prover as p2g_prove_dev; the proof is identical to
 * p2g_prove on the host-filled matrix. */
typedef struct p2g_wmap p2g_wmap;
int32_t p2g_wmap_load(p2g_ctx* ctx, const p2g_circuit* c, const int32_t* wire_map /*[num_wires][n]*/, uint32_t num_slots,
                      const int64_t* fixed_pos, const uint64_t* fixed_val, uint32_t num_fixed, p2g_wmap** out);
int32_t p2g_wmap_free(p2g_ctx* ctx, p2g_wmap* m);
int32_t p2g_prove_slots(p2g_ctx* ctx, const p2g_circuit* c, const p2g_wmap* m, const uint64_t* slots_host /*[num_slots]*/,
                        const uint64_t* public_inputs, uint64_t* proof_out, size_t proof_cap_words, size_t* proof_words_out);
/* the gathered wire matrix of the last p2g_prove_slots call is not kept; this fills one for tests */
int32_t p2g_wmap_fill(p2g_ctx* ctx, const p2g_circuit* c, const p2g_wmap* m, const uint64_t* slots_host, uint64_t* wires_out_host);
/* Witness generation on the device: generate_partial_witness + set_lookup_wires
 * (plonky2 iop/generator.rs, plonk/prover.rs) for the ArithmeticGate / LookupGate / equality / ConstantGate /
 * PoseidonGate generators, from the same program description libp2witness.so interprets on the host.  The program
 * is level-scheduled at load; one warp evaluates one witness.  `input_slots`: the partitions the caller sets
 * (PartialWitness::set_target), fixed per program; every call then passes their values in the same order.
 * p2g_prove_inputs = generators + full_witness + prove, with a few hundred input values as the only H2D traffic;
 * the proof equals p2g_prove on the host-generated wire matrix.  Generator failures return the P2W_E_* code of
 * include/p2witness.h (P2W_E_CONFLICT: a preset partition disagrees with the generated value -- how the reference's
 * prove() rejects a wrong ciphertext, /root/reference/aes-gcm/src/circuit_aes.rs:403-405; P2W_E_LOOKUP). */
typedef struct p2g_wprog p2g_wprog;
int32_t p2g_wprog_load(p2g_ctx* ctx, const p2w_program_desc* prog, const int32_t* input_slots, uint32_t num_inputs, p2g_wprog** out);
int32_t p2g_wprog_free(p2g_ctx* ctx, p2g_wprog* p);
uint32_t p2g_wprog_ext_slots(const p2g_wprog* p);   /* = p2w_ext_slots of the same program */
uint32_t p2g_wprog_levels(const p2g_wprog* p);      /* dependency depth of the program */
/* `count` witnesses: input_vals [count][num_inputs] (host) -> extended slot vectors [count][ext_slots] (host) */
int32_t p2g_wprog_generate(p2g_ctx* ctx, const p2g_wprog* p, const uint64_t* input_vals, uint32_t count, uint64_t* ext_out);
/* batch form, witnesses stay in HBM: ext_dev [count][ext_slots] and flags_dev [count] are caller-owned DEVICE buffers;
 * asynchronous (ordered on the context's stream).  flags: 0 ok, bit 0 non-canonical input, bit 1 lookup miss, bit 2
 * conflict.  p2g_prove_slots_dev takes one such vector (device pointer) where p2g_prove_slots takes a host one. */
int32_t p2g_wprog_generate_dev(p2g_ctx* ctx, const p2g_wprog* p, const uint64_t* input_vals_host, uint32_t count,
                               uint64_t* ext_dev, int32_t* flags_dev);
int32_t p2g_prove_slots_dev(p2g_ctx* ctx, const p2g_circuit* c, const p2g_wmap* m, const uint64_t* slots_dev,
                            const uint64_t* public_inputs, uint64_t* proof_out, size_t proof_cap_words, size_t* proof_words_out);
int32_t p2g_prove_inputs(p2g_ctx* ctx, const p2g_circuit* c, const p2g_wmap* m, const p2g_wprog* prog,
                         const uint64_t* input_vals_host /*[num_inputs]*/, const uint64_t* public_inputs, uint64_t* proof_out,
                         size_t proof_cap_words, size_t* proof_words_out);
int32_t p2g_last_transcript(p2g_ctx* ctx, p2g_transcript* out);
int32_t p2g_last_zs_values(p2g_ctx* ctx, uint64_t* out /*[num_zs_cols][n]*/);
int32_t p2g_last_quotient_chunks(p2g_ctx* ctx, uint64_t* out /*[num_challenges*qdf][n]*/);

/* per-stage device times of the last p2g_prove, milliseconds (CUDA events on the ctx stream) */
typedef struct {
    float h2d, wires_commit, zs_build, zs_commit, quotient, quotient_commit, openings, fri_combine,
          fri_commit, pow, queries, total;
} p2g_timings;
int32_t p2g_last_timings(p2g_ctx* ctx, p2g_timings* out);
/* bit 0: per-stage CUDA-event timing; bit 1: keep stage dumps for the read-backs above */
int32_t p2g_set_timing(p2g_ctx* ctx, int32_t enabled);
/* device ms of the last commit's three kernels groups: inverse NTT, coset LDE, Merkle tree */
int32_t p2g_last_commit_timings(p2g_ctx* ctx, float out[3]);
/* Debug aid (contexts created with P2G_CANARY=1 in the environment): every device block the library allocates gets a
 * guard band that is verified when the block is released; returns the number of blocks checked and of overwritten
 * bands.  P2G_E_BADARG when the context was not created in canary mode. */
int32_t p2g_debug_canary(p2g_ctx* ctx, uint64_t* blocks_checked, uint64_t* failures);
/* kernels launched by the library since it was loaded (all contexts) */
uint64_t p2g_launch_count(void);

/* ---- individual hot-path stages, exposed for parity tests and kernel benchmarks -------------- */
/* fri_proof_of_work (fri/prover.rs): lowest nonce whose response has >= pow_bits leading zeros.
 * state: the 12-word duplex state with buffered inputs already overwritten; pos: slot of the nonce. */
int32_t p2g_pow_grind(p2g_ctx* ctx, const uint64_t state[12], uint32_t pos, uint32_t pow_bits, uint64_t* nonce_out);
/* one FRI commit-phase fold (arity 2^arity_bits) of N ext values in bit-reversed order on the
 * coset shift*<w_N>: out[k] = P'(x_k^arity) for the reference's coefficient-domain fold. */
int32_t p2g_fri_fold(p2g_ctx* ctx, const uint64_t* values_host /*[N][2]*/, uint32_t log_n_values, uint32_t arity_bits,
                     uint64_t shift, const uint64_t beta[2], uint64_t* out_host /*[N>>arity][2]*/);
/* chained Poseidon permutations without memory traffic: the INT-pipe peak used as roofline
 * denominator for the Merkle kernels.  Returns permutations per second. */
int32_t p2g_poseidon_peak(p2g_ctx* ctx, uint32_t iters, double* perms_per_sec);
/* Device field arithmetic exposed for parity tests (plonky2 field/src/goldilocks_field.rs Add / Sub /
 * Mul and to_canonical_u64): for i < n,
 *   out[0][i] = a+b, out[1][i] = a-b, out[2][i] = a*b   (a, b canonical),
 *   out[3][i] = canonical form of the arbitrary u64 la[i],
 *   out[4][i] = canonical form of la[i]*lb[i]            (la, lb arbitrary u64 residues),
 *   out[5][i] = canonical form of a * 2^(i mod 96)       (the shift-multiply of the last NTT pass). */
int32_t p2g_field_ops(p2g_ctx* ctx, const uint64_t* a, const uint64_t* b, const uint64_t* la, const uint64_t* lb,
                      size_t n, uint64_t* out /*[6][n]*/);

#ifdef __cplusplus
}
#endif
#endif
