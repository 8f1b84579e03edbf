//! Raw bindings of include/p2gpu.h — one item per C declaration, same order.
//! The safe wrapper (`GpuProver`) below is what the patched
//! `plonky2::plonk::prover::prove_with_partition_witness` calls (rust/plonky2-patch/prover.rs).
#![allow(non_camel_case_types)]
use std::ffi::CStr;
use std::os::raw::{c_char, c_void};

#[repr(C)] pub struct p2g_ctx { _p: [u8; 0] }
#[repr(C)] pub struct p2g_batch { _p: [u8; 0] }
#[repr(C)] pub struct p2g_wmap { _p: [u8; 0] }
#[repr(C)] pub struct p2g_circuit { _p: [u8; 0] }

pub const P2G_OK: i32 = 0;
pub const P2G_E_CUDA: i32 = -1;
pub const P2G_E_BADARG: i32 = -2;
pub const P2G_E_UNSAT: i32 = -3;
pub const P2G_E_POW: i32 = -4;

pub const P2G_GATE_NOOP: i32 = 0;
pub const P2G_GATE_CONSTANT: i32 = 1;
pub const P2G_GATE_PUBLIC_INPUT: i32 = 2;
pub const P2G_GATE_ARITHMETIC: i32 = 3;
pub const P2G_GATE_LOOKUP: i32 = 4;
pub const P2G_GATE_LOOKUP_TABLE: i32 = 5;
pub const P2G_GATE_POSEIDON: i32 = 6;

#[repr(C)] #[derive(Clone, Copy, Debug)]
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

extern "C" {
    pub fn p2g_version() -> i32;
    pub fn p2g_ctx_create(device: i32, out: *mut *mut p2g_ctx) -> i32;
    pub fn p2g_ctx_destroy(ctx: *mut p2g_ctx);
    pub fn p2g_last_error(ctx: *mut p2g_ctx) -> *const c_char;
    pub fn p2g_ctx_sync(ctx: *mut p2g_ctx) -> i32;
    pub fn p2g_ctx_stream(ctx: *mut p2g_ctx) -> *mut c_void;
    pub fn p2g_commit_from_values(ctx: *mut p2g_ctx, cols: *const u64, ncols: u32, log_n: u32, rate_bits: u32,
                                  cap_height: u32, out: *mut *mut p2g_batch, cap_out: *mut u64) -> i32;
    pub fn p2g_commit_from_coeffs(ctx: *mut p2g_ctx, cols: *const u64, ncols: u32, log_n: u32, rate_bits: u32,
                                  cap_height: u32, out: *mut *mut p2g_batch, cap_out: *mut u64) -> i32;
    pub fn p2g_batch_free(ctx: *mut p2g_ctx, b: *mut p2g_batch) -> i32;
    pub fn p2g_batch_open_leaf(ctx: *mut p2g_ctx, b: *const p2g_batch, leaf: u64, row_out: *mut u64, siblings_out: *mut u64) -> i32;
    pub fn p2g_circuit_load(ctx: *mut p2g_ctx, desc: *const p2g_circuit_desc, out: *mut *mut p2g_circuit, cap_out: *mut u64) -> i32;
    pub fn p2g_circuit_free(ctx: *mut p2g_ctx, c: *mut p2g_circuit) -> i32;
    pub fn p2g_proof_words(c: *const p2g_circuit) -> usize;
    pub fn p2g_prove(ctx: *mut p2g_ctx, c: *const p2g_circuit, wires: *const u64, public_inputs: *const u64,
                     proof_out: *mut u64, cap_words: usize, words_out: *mut usize) -> i32;
    pub fn p2g_prove_dev(ctx: *mut p2g_ctx, c: *const p2g_circuit, wires_dev: *const u64, public_inputs: *const u64,
                         proof_out: *mut u64, cap_words: usize, words_out: *mut usize) -> i32;
    // device-side PartitionWitness::full_witness: wire_map = representative_map flattened to [col*n + row]
    pub fn p2g_wmap_load(ctx: *mut p2g_ctx, c: *const p2g_circuit, wire_map: *const i32, num_slots: u32,
                         fixed_pos: *const i64, fixed_val: *const u64, num_fixed: u32, out: *mut *mut p2g_wmap) -> i32;
    pub fn p2g_wmap_free(ctx: *mut p2g_ctx, m: *mut p2g_wmap) -> i32;
    pub fn p2g_prove_slots(ctx: *mut p2g_ctx, c: *const p2g_circuit, m: *const p2g_wmap, slots: *const u64,
                           public_inputs: *const u64, proof_out: *mut u64, cap_words: usize, words_out: *mut usize) -> i32;
    pub fn p2g_set_timing(ctx: *mut p2g_ctx, enabled: i32) -> i32;
    pub fn p2g_last_timings(ctx: *mut p2g_ctx, out: *mut p2g_timings) -> i32;
}

/// One GPU context + one loaded circuit.  `!Sync`: a context is used by one host thread at a time
/// (spawn one `GpuProver` per in-flight proof, exactly like `host/sharding.py::BatchProver`).
pub struct GpuProver { ctx: *mut p2g_ctx, circuit: *mut p2g_circuit }
unsafe impl Send for GpuProver {}

#[derive(Debug)]
pub enum GpuError { Unsatisfied, Backend(i32, String) }

impl GpuProver {
    /// # Safety: every pointer inside `desc` must be valid for the duration of the call.
    pub unsafe fn load(device: i32, desc: &p2g_circuit_desc, expected_cap: &[u64]) -> Result<Self, GpuError> {
        let mut ctx = std::ptr::null_mut();
        let rc = p2g_ctx_create(device, &mut ctx);
        if rc != P2G_OK { return Err(GpuError::Backend(rc, "no CUDA device".into())); }
        let mut circuit = std::ptr::null_mut();
        let mut cap = vec![0u64; expected_cap.len()];
        let rc = p2g_circuit_load(ctx, desc, &mut circuit, cap.as_mut_ptr());
        if rc != P2G_OK { let e = Self::err(ctx, rc); p2g_ctx_destroy(ctx); return Err(e); }
        // the GPU commitment of (constants, sigmas) must equal VerifierOnlyCircuitData::constants_sigmas_cap
        assert_eq!(cap, expected_cap, "constants_sigmas cap mismatch between CPU build() and GPU load");
        Ok(Self { ctx, circuit })
    }
    unsafe fn err(ctx: *mut p2g_ctx, rc: i32) -> GpuError {
        if rc == P2G_E_UNSAT { return GpuError::Unsatisfied; }
        GpuError::Backend(rc, CStr::from_ptr(p2g_last_error(ctx)).to_string_lossy().into_owned())
    }
    /// `wires`: num_wires columns of `degree` canonical u64, column-major.  Returns the flat proof words.
    pub fn prove(&mut self, wires: &[u64], public_inputs: &[u64]) -> Result<Vec<u64>, GpuError> {
        unsafe {
            let cap = p2g_proof_words(self.circuit);
            let mut out = vec![0u64; cap];
            let mut n = 0usize;
            let rc = p2g_prove(self.ctx, self.circuit, wires.as_ptr(), public_inputs.as_ptr(), out.as_mut_ptr(), cap, &mut n);
            if rc != P2G_OK { return Err(Self::err(self.ctx, rc)); }
            out.truncate(n);
            Ok(out)
        }
    }
}
impl Drop for GpuProver {
    fn drop(&mut self) { unsafe { p2g_circuit_free(self.ctx, self.circuit); p2g_ctx_destroy(self.ctx); } }
}
