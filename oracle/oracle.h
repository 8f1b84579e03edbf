/* TEST INFRASTRUCTURE — CPU oracle (see gl64.h header).  PARITY UNPINNED against the Rust
 * dependency; pinned only by the recalled upstream Poseidon vectors (tests/test_oracle_kat.py)
 * and by algebraic self-checks + the restated verifier accepting every proof.
 *
 * Restates, by upstream path (0xPolygonZero/plonky2, the crate the reference pins at
 * /root/reference/Cargo.toml:12 and enters through `data.prove(pw)`, e.g.
 * /root/reference/aes-gcm/src/circuit_gcm.rs:781):
 *   field/src/fft.rs, field/src/polynomial/mod.rs            -> orc_fft, orc_ifft, orc_coset_fft, orc_coset_ifft
 *   plonky2/src/hash/{poseidon,poseidon_goldilocks,hashing}  -> orc_poseidon, orc_hash_*
 *   plonky2/src/hash/{merkle_tree,merkle_proofs}.rs          -> orc_merkle_*
 *   plonky2/src/fri/oracle.rs (PolynomialBatch)              -> orc_batch_*
 *   plonky2/src/iop/challenger.rs                            -> orc_challenger_*
 *   plonky2/src/plonk/{prover,vanishing_poly,verifier}.rs,
 *   plonky2/src/fri/{prover,verifier}.rs                     -> orc_prove / orc_verify (prover.c)
 */
#ifndef ORACLE_H
#define ORACLE_H
#include "gl64.h"

#ifdef __cplusplus
extern "C" {
#endif

/* ---- Poseidon / hashing ---- */
void orc_poseidon(gl_t s[12]);        /* sparse partial rounds (what the prover uses) */
void orc_poseidon_naive(gl_t s[12]);  /* textbook rounds; tests check both agree */
void orc_hash_n_to_m_no_pad(const gl_t* in, size_t n, gl_t* out, size_t m);
void orc_hash_no_pad(const gl_t* in, size_t n, gl_t out[4]);
void orc_hash_or_noop(const gl_t* in, size_t n, gl_t out[4]);
void orc_two_to_one(const gl_t l[4], const gl_t r[4], gl_t out[4]);

/* ---- FFT (natural order in and out, as upstream) ---- */
void orc_fft(gl_t* a, int log_n);
void orc_ifft(gl_t* a, int log_n);
void orc_coset_fft(gl_t* a, int log_n, gl_t shift);
void orc_coset_ifft(gl_t* a, int log_n, gl_t shift);
size_t orc_reverse_bits(size_t x, int bits);

/* ---- Merkle tree over row-major leaves ---- */
typedef struct {
    size_t num_leaves;   /* power of two */
    size_t leaf_len;
    int cap_height;
    const gl_t* leaves;  /* borrowed: [num_leaves][leaf_len] */
    gl_t* digests;       /* levels 0..L-1 concatenated: level k has num_leaves>>k digests of 4; L = log2(num_leaves)-cap_height */
    gl_t* cap;           /* [2^cap_height][4] */
} orc_merkle;
orc_merkle* orc_merkle_new(const gl_t* leaves, size_t num_leaves, size_t leaf_len, int cap_height);
void orc_merkle_free(orc_merkle* t);
size_t orc_merkle_path_len(const orc_merkle* t);
/* siblings out: [path_len][4] */
void orc_merkle_prove(const orc_merkle* t, size_t leaf_index, gl_t* siblings);
int orc_merkle_verify(const gl_t* leaf, size_t leaf_len, size_t leaf_index, const gl_t* cap, int cap_height,
                      const gl_t* siblings, size_t path_len);
const gl_t* orc_merkle_level(const orc_merkle* t, int level); /* level == L returns cap */

/* ---- PolynomialBatch ---- */
typedef struct {
    int ncols, log_n, rate_bits, cap_height;
    gl_t* coeffs;   /* [ncols][n] */
    gl_t* leaves;   /* [N][ncols], row j = evaluations at 7*w_N^{bitrev(j)} */
    orc_merkle* tree;
} orc_batch;
orc_batch* orc_batch_from_values(const gl_t* cols, int ncols, int log_n, int rate_bits, int cap_height);
orc_batch* orc_batch_from_coeffs(const gl_t* cols, int ncols, int log_n, int rate_bits, int cap_height);
void orc_batch_free(orc_batch* b);

/* ---- Challenger ---- */
typedef struct {
    gl_t state[12];
    gl_t in_buf[8]; int in_len;
    gl_t out_buf[8]; int out_len;
} orc_challenger;
void orc_challenger_init(orc_challenger* c);
void orc_challenger_observe(orc_challenger* c, gl_t x);
void orc_challenger_observe_many(orc_challenger* c, const gl_t* x, size_t n);
gl_t orc_challenger_get(orc_challenger* c);
ext_t orc_challenger_get_ext(orc_challenger* c);

/* ---- Circuit description & proof (prover.c) ---- */
enum { ORC_GATE_NOOP = 0, ORC_GATE_CONSTANT = 1, ORC_GATE_PUBLIC_INPUT = 2, ORC_GATE_ARITHMETIC = 3,
       ORC_GATE_LOOKUP = 4, ORC_GATE_LOOKUP_TABLE = 5, ORC_GATE_POSEIDON = 6 };

typedef struct {
    int32_t kind;
    int32_t selector_index;       /* which selector polynomial */
    int32_t group_start, group_end; /* gate index range sharing that selector */
    int32_t num_constraints;
    int32_t param0;               /* ARITHMETIC: num_ops; CONSTANT: num_consts */
} orc_gate;

typedef struct {
    int32_t degree_bits;
    int32_t num_wires, num_routed_wires, num_constants; /* num_constants = gate constants only */
    int32_t num_challenges, quotient_degree_factor;
    int32_t rate_bits, cap_height, pow_bits, num_query_rounds;
    int32_t num_reduction_arity_bits; int32_t reduction_arity_bits[16];
    int32_t num_selectors, num_lookup_selectors;
    int32_t num_gates; const orc_gate* gates;
    int32_t num_gate_constraints;
    int32_t num_partial_products;
    int32_t num_luts;
    const int32_t* lut_lens;      /* [num_luts] */
    const uint16_t* lut_data;     /* concatenated (inp,out) pairs */
    const int32_t* lookup_rows;   /* [num_luts][3] = last_lu_gate, last_lut_gate, first_lut_gate */
    int32_t num_public_inputs;
    const gl_t* k_is;             /* [num_routed_wires] */
    const gl_t* constants_sigmas; /* [(num_selectors+num_lookup_selectors+num_constants+num_routed_wires)][n] values */
    gl_t circuit_digest[4];
} orc_circuit;

typedef struct orc_prover_data orc_prover_data;  /* circuit + preprocessed commitment */
orc_prover_data* orc_circuit_load(const orc_circuit* c);
void orc_circuit_free(orc_prover_data* pd);
const gl_t* orc_circuit_cap(const orc_prover_data* pd);

/* Proof is serialised into a flat u64 buffer (layout documented in prover.c / DESIGN.md).
 * Returns number of u64 written, or negative error (-3 = unsatisfied witness). */
long orc_prove(const orc_prover_data* pd, const gl_t* wires /*[num_wires][n]*/, const gl_t* public_inputs,
               gl_t* proof_out, size_t proof_cap);
/* 0 = accept, negative = reject code */
int orc_verify(const orc_prover_data* pd, const gl_t* proof, size_t proof_len);
size_t orc_proof_len(const orc_circuit* c);

/* stage dumps for GPU parity tests */
typedef struct {
    gl_t betas[4], gammas[4], deltas[16], alphas[4];
    ext_t zeta, fri_alpha; ext_t fri_betas[16];
    gl_t pow_witness;
    uint64_t query_indices[64];
} orc_transcript;
long orc_prove_debug(const orc_prover_data* pd, const gl_t* wires, const gl_t* public_inputs,
                     gl_t* proof_out, size_t proof_cap, orc_transcript* tr,
                     gl_t* zs_pp_lookup_values /* optional [num_zs_cols][n] */,
                     gl_t* quotient_chunk_coeffs /* optional [nch*qdf][n] */);
int orc_num_threads(void);
void orc_set_num_threads(int n);   /* launchers such as torchrun export OMP_NUM_THREADS=1 */

#ifdef __cplusplus
}
#endif
#endif
