//! `#[repr(C)]` mirrors of the structs and constants of include/p2gpu.h and include/p2witness.h
//! (field order and widths as in the headers; tests/test_abi.py checks the field lists).
#![allow(non_camel_case_types)]

macro_rules! opaque { ($($n:ident),*) => { $(#[repr(C)] pub struct $n { _p: [u8; 0] })* } }
opaque!(p2g_ctx, p2g_batch, p2g_circuit, p2g_wmap, p2g_wprog);

/// all-gather callback of `p2g_prove_sharded`: gather `bytes_per_rank` bytes of every rank's send buffer into every
/// rank's receive buffer, return 0 when the receive buffer is complete
pub type p2g_exchange_fn = Option<unsafe extern "C" fn(user: *mut std::os::raw::c_void, stage: i32, bytes_per_rank: u64) -> i32>;

pub const P2G_OK: i32 = 0;
pub const P2G_E_CUDA: i32 = -1;
pub const P2G_E_BADARG: i32 = -2;
pub const P2G_E_UNSAT: i32 = -3;
pub const P2G_E_POW: i32 = -4;
pub const P2G_E_NOMEM: i32 = -5;
pub const P2W_E_CONFLICT: i32 = -10;
pub const P2W_E_LOOKUP: i32 = -11;
pub const P2W_E_UNSET: i32 = -12;

pub const P2G_GATE_NOOP: i32 = 0;
pub const P2G_GATE_CONSTANT: i32 = 1;
pub const P2G_GATE_PUBLIC_INPUT: i32 = 2;
pub const P2G_GATE_ARITHMETIC: i32 = 3;
pub const P2G_GATE_LOOKUP: i32 = 4;
pub const P2G_GATE_LOOKUP_TABLE: i32 = 5;
pub const P2G_GATE_POSEIDON: i32 = 6;

pub const P2W_OP_ARITH: i32 = 0;
pub const P2W_OP_LOOKUP: i32 = 1;
pub const P2W_OP_EQ: i32 = 2;
pub const P2W_OP_CONST: i32 = 3;
pub const P2W_OP_POSEIDON: i32 = 4;

#[repr(C)] #[derive(Clone, Copy, Debug, Default)]
pub struct p2g_gate {
    pub kind: i32, pub selector_index: i32, pub group_start: i32, pub group_end: i32,
    pub num_constraints: i32, pub param0: i32,
}

#[repr(C)]
pub struct p2g_circuit_desc {
    pub degree_bits: i32,
    pub num_wires: i32, pub num_routed_wires: i32, pub num_constants: i32,
    pub num_challenges: i32, pub quotient_degree_factor: i32,
    pub rate_bits: i32, pub cap_height: i32, pub pow_bits: i32, pub num_query_rounds: i32,
    pub num_reduction_arity_bits: i32, pub reduction_arity_bits: [i32; 16],
    pub num_selectors: i32, pub num_lookup_selectors: i32,
    pub num_gates: i32, pub gates: *const p2g_gate,
    pub num_gate_constraints: i32,
    pub num_partial_products: i32,
    pub num_luts: i32,
    pub lut_lens: *const i32, pub lut_data: *const u16, pub lookup_rows: *const i32,
    pub num_public_inputs: i32,
    pub k_is: *const u64, pub constants_sigmas: *const u64,
    pub circuit_digest: [u64; 4],
}

#[repr(C)] #[derive(Default, Clone, Copy, Debug)]
pub struct p2g_timings {
    pub h2d: f32, pub wires_commit: f32, pub zs_build: f32, pub zs_commit: f32, pub quotient: f32,
    pub quotient_commit: f32, pub openings: f32, pub fri_combine: f32, pub fri_commit: f32, pub pow: f32,
    pub queries: f32, pub total: f32,
}

#[repr(C)] #[derive(Clone, Copy)]
pub struct p2g_transcript {
    pub betas: [u64; 4], pub gammas: [u64; 4], pub deltas: [u64; 16], pub alphas: [u64; 4],
    pub zeta: [u64; 2], pub fri_alpha: [u64; 2], pub fri_betas: [u64; 32],
    pub pow_witness: u64,
    pub query_indices: [u64; 64],
}

#[repr(C)]
pub struct p2w_program_desc {
    pub num_slots: u32,
    pub num_ops: u32,
    pub ops: *const i32, pub op_consts: *const u64,
    pub num_luts: u32,
    pub lut_lens: *const i32, pub lut_data: *const u16,
    pub num_wires: u32, pub log_n: u32,
    pub wire_slot: *const i32,
    pub num_fixed: u32,
    pub fixed_pos: *const i64, pub fixed_val: *const u64,
    pub lookup_counts: *const i32, pub lookup_slots: *const i32, pub lookup_padding: *const i32,
    pub mult_pos: *const i64,
    pub num_poseidon: u32,
    pub poseidon_rows: *const i32,
}
