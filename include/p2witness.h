/* p2witness.h — host-side witness generation helper (libp2witness.so, plain C++, no CUDA).
 *
 * Not on the GPU hot path: north_star keeps "circuit building and witness generation on the
 * host".  In the reference this work is done by plonky2's generators
 * (iop/generator.rs::generate_partial_witness + prover.rs::set_lookup_wires, entered from
 * `data.prove(pw)`, /root/reference/aes-gcm/src/circuit_gcm.rs:781); the Python circuit
 * builder of this repo records the same generators as a straight-line program and this
 * library evaluates it, producing the full wire matrix that p2g_prove consumes.
 */
#ifndef P2WITNESS_H
#define P2WITNESS_H
#include <stdint.h>
#include <stddef.h>
#ifdef __cplusplus
extern "C" {
#endif

enum { P2W_OP_ARITH = 0,   /* s0 = c0*s1*s2 + c1*s3                         (ArithmeticGate generator) */
       P2W_OP_LOOKUP = 1,  /* s0 = lut[s4 as lut index][value(s1)]          (LookupGenerator)          */
       P2W_OP_EQ = 2,      /* s0 = (s2 == s3), s1 = (s2 - s3)^-1 or 0       (EqualityGenerator)        */
       P2W_OP_CONST = 3,   /* s0 = c0                                        (ConstantGate generator)   */
       P2W_OP_POSEIDON = 4 /* s0 = index into poseidon_rows: PoseidonGate generator (gates/poseidon.rs) */ };

#define P2W_E_CONFLICT (-10)  /* a partition was set twice with different values */
#define P2W_E_LOOKUP (-11)    /* looked-up value is not a key of the table */
#define P2W_E_UNSET (-12)     /* a generator input was never set */
#define P2W_E_BADARG (-2)

typedef struct {
    uint32_t num_slots;
    uint32_t num_ops;
    const int32_t* ops;          /* [num_ops][6]: kind, s0, s1, s2, s3, s4 */
    const uint64_t* op_consts;   /* [num_ops][2] */
    uint32_t num_luts;
    const int32_t* lut_lens;     /* [num_luts] */
    const uint16_t* lut_data;    /* (inp,out) pairs concatenated */
    uint32_t num_wires, log_n;
    const int32_t* wire_slot;    /* [num_wires][n] column-major: slot id or -1 */
    uint32_t num_fixed;
    const int64_t* fixed_pos;    /* col*n + row */
    const uint64_t* fixed_val;
    /* set_lookup_wires: per LUT, the slots looked up (multiplicity counting), the number of
     * padded LookupGate slots (counted on entry 0) and the wire position of every entry's
     * multiplicity cell */
    const int32_t* lookup_counts;   /* [num_luts] */
    const int32_t* lookup_slots;    /* concatenated input slots */
    const int32_t* lookup_padding;  /* [num_luts] */
    const int64_t* mult_pos;        /* concatenated, one per LUT entry */
    /* PoseidonGate rows: [num_poseidon][25] = row, 12 input slots, 12 output slots */
    uint32_t num_poseidon;
    const int32_t* poseidon_rows;
} p2w_program_desc;

typedef struct p2w_program p2w_program;
int32_t p2w_program_create(const p2w_program_desc* d, p2w_program** out);
void p2w_program_destroy(p2w_program* p);
/* inputs: (slot, value) pairs set by the caller (PartialWitness::set_target).
 * wires_out: [num_wires][n] column-major, fully overwritten. */
int32_t p2w_generate(const p2w_program* p, const int32_t* input_slots, const uint64_t* input_vals,
                     uint32_t num_inputs, uint64_t* wires_out);
/* Two-step form for the device-side wire fill (p2g_prove_slots, include/p2gpu.h).  The generators and
 * set_lookup_wires fill an "extended slot vector" -- the analogue of PartitionWitness::values
 * (iop/witness.rs): program slots, then the multiplicity of every LUT entry, then the 111 internal
 * wires of every PoseidonGate row -- and `wire_map` ([num_wires][n], index into that vector or -1)
 * plus the list of constant cells say where each value goes, which is what
 * PartitionWitness::full_witness does with representative_map. */
uint32_t p2w_ext_slots(const p2w_program* p);
int32_t p2w_wire_map(const p2w_program* p, int32_t* map_out /*[num_wires][n]*/);
int32_t p2w_fixed_cells(const p2w_program* p, uint32_t* count, const int64_t** pos, const uint64_t** val);
int32_t p2w_generate_slots(const p2w_program* p, const int32_t* input_slots, const uint64_t* input_vals,
                           uint32_t num_inputs, uint64_t* ext_out /*[p2w_ext_slots]*/);
int32_t p2w_generate_slots_many(const p2w_program* p, const int32_t* input_slots, const uint64_t* input_vals,
                                uint32_t num_inputs, uint32_t count, uint64_t* ext_out);
/* threads of the *_many entry points; 0 = OpenMP default (OMP_NUM_THREADS / all cores) */
void p2w_set_num_threads(int32_t n);
int32_t p2w_num_threads(void);
/* p2w_generate for `count` independent witnesses laid out back to back (OpenMP over witnesses) */
int32_t p2w_generate_many(const p2w_program* p, const int32_t* input_slots, const uint64_t* input_vals,
                          uint32_t num_inputs, uint32_t count, uint64_t* wires_out);
#ifdef __cplusplus
}
#endif
#endif
