//! FFI crate for libp2gpu.so — the B200 backend of plonky2's `prove()` hot path (include/p2gpu.h).
//!
//! `ffi` holds one `extern "C"` item per C declaration, GENERATED from the header
//! (tools/gen_rust_bindings.py; tests/test_abi.py fails when the two drift); `types` the `#[repr(C)]`
//! structs; below, the safe layer the patched `plonky2::plonk::prover` uses
//! (rust/plonky2-patch/src/gpu_prover.rs).
//!
//! NOT compiled in the build image (no cargo / rustc there).  Link: `P2GPU_LIB_DIR=<repo>/plonky2_aes_b200`.
#![allow(non_camel_case_types)]
pub mod ffi;
pub mod types;

use std::ffi::CStr;
use std::ptr;

pub use ffi::*;
pub use types::*;

#[derive(Debug)]
pub enum GpuError {
    /// witness generation found a preset partition that disagrees with the generated value
    /// (`prove(..).is_err()` of aes-gcm/src/circuit_aes.rs:403-405)
    WitnessConflict,
    Backend(i32, String),
}
impl std::fmt::Display for GpuError {
    fn fmt(&self, f: &mut std::fmt::Formatter<'_>) -> std::fmt::Result {
        match self {
            GpuError::WitnessConflict => write!(f, "partition set twice with different values"),
            GpuError::Backend(c, m) => write!(f, "p2gpu error {c}: {m}"),
        }
    }
}
impl std::error::Error for GpuError {}

/// One context = one CUDA stream + memory pool on one device.  `!Sync`: used by one host thread at a time;
/// keep one per in-flight proof (several proofs in flight hide each other's Fiat-Shamir round trips).
pub struct GpuContext { raw: *mut p2g_ctx, pub device: i32 }
unsafe impl Send for GpuContext {}
impl GpuContext {
    pub fn new(device: i32) -> Result<Self, GpuError> {
        let mut raw = ptr::null_mut();
        let rc = unsafe { p2g_ctx_create(device, &mut raw) };
        if rc != P2G_OK { return Err(GpuError::Backend(rc, "p2g_ctx_create: no usable CUDA device (there is no CPU fallback)".into())); }
        Ok(Self { raw, device })
    }
    fn err(&self, rc: i32) -> GpuError {
        if rc == P2W_E_CONFLICT { return GpuError::WitnessConflict; }
        let msg = unsafe { CStr::from_ptr(p2g_last_error(self.raw)) }.to_string_lossy().into_owned();
        GpuError::Backend(rc, msg)
    }
    pub fn raw(&self) -> *mut p2g_ctx { self.raw }
}
impl Drop for GpuContext { fn drop(&mut self) { unsafe { p2g_ctx_destroy(self.raw) } } }

/// Owned host copy of everything `p2g_circuit_desc` points to (the descriptor borrows from it).
#[derive(Clone, Default)]
pub struct CircuitDescOwned {
    pub degree_bits: i32, pub num_wires: i32, pub num_routed_wires: i32, pub num_constants: i32,
    pub num_challenges: i32, pub quotient_degree_factor: i32, pub rate_bits: i32, pub cap_height: i32,
    pub pow_bits: i32, pub num_query_rounds: i32, pub reduction_arity_bits: Vec<i32>,
    pub num_selectors: i32, pub num_lookup_selectors: i32, pub gates: Vec<p2g_gate>,
    pub num_gate_constraints: i32, pub num_partial_products: i32,
    pub lut_lens: Vec<i32>, pub lut_data: Vec<u16>, pub lookup_rows: Vec<i32>,
    pub num_public_inputs: i32, pub k_is: Vec<u64>,
    /// [(selectors + lookup selectors + constants + routed wires)][n] VALUES on the subgroup, column-major
    pub constants_sigmas: Vec<u64>,
    pub circuit_digest: [u64; 4],
}
impl CircuitDescOwned {
    pub fn as_desc(&self) -> p2g_circuit_desc {
        let mut rab = [0i32; 16];
        rab[..self.reduction_arity_bits.len()].copy_from_slice(&self.reduction_arity_bits);
        p2g_circuit_desc {
            degree_bits: self.degree_bits, num_wires: self.num_wires, num_routed_wires: self.num_routed_wires,
            num_constants: self.num_constants, num_challenges: self.num_challenges,
            quotient_degree_factor: self.quotient_degree_factor, rate_bits: self.rate_bits, cap_height: self.cap_height,
            pow_bits: self.pow_bits, num_query_rounds: self.num_query_rounds,
            num_reduction_arity_bits: self.reduction_arity_bits.len() as i32, reduction_arity_bits: rab,
            num_selectors: self.num_selectors, num_lookup_selectors: self.num_lookup_selectors,
            num_gates: self.gates.len() as i32, gates: self.gates.as_ptr(),
            num_gate_constraints: self.num_gate_constraints, num_partial_products: self.num_partial_products,
            num_luts: self.lut_lens.len() as i32, lut_lens: self.lut_lens.as_ptr(), lut_data: self.lut_data.as_ptr(),
            lookup_rows: self.lookup_rows.as_ptr(), num_public_inputs: self.num_public_inputs,
            k_is: self.k_is.as_ptr(), constants_sigmas: self.constants_sigmas.as_ptr(), circuit_digest: self.circuit_digest,
        }
    }
    /// bytes of `ProofWithPublicInputs::to_bytes` for the flat proof words (host-only code of the library)
    pub fn proof_words_to_bytes(&self, words: &[u64]) -> Result<Vec<u8>, GpuError> {
        let d = self.as_desc();
        unsafe {
            let cap = p2g_proof_bytes_len(&d);
            let mut out = vec![0u8; cap];
            let mut len = 0usize;
            let rc = p2g_proof_to_bytes(&d, words.as_ptr(), words.len(), out.as_mut_ptr(), cap, &mut len);
            if rc != P2G_OK { return Err(GpuError::Backend(rc, "p2g_proof_to_bytes".into())); }
            out.truncate(len);
            Ok(out)
        }
    }
}

/// The prover data of one `CircuitData` resident on one device (preprocessed commitment, domain tables).
pub struct GpuCircuit { raw: *mut p2g_circuit, ctx: *mut p2g_ctx, pub desc: CircuitDescOwned }
unsafe impl Send for GpuCircuit {}
impl GpuCircuit {
    /// `expected_cap`: `VerifierOnlyCircuitData::constants_sigmas_cap` flattened; the GPU commitment must equal it.
    pub fn load(ctx: &GpuContext, desc: CircuitDescOwned, expected_cap: &[u64]) -> Result<Self, GpuError> {
        let d = desc.as_desc();
        let mut raw = ptr::null_mut();
        let mut cap = vec![0u64; expected_cap.len()];
        let rc = unsafe { p2g_circuit_load(ctx.raw, &d, &mut raw, cap.as_mut_ptr()) };
        if rc != P2G_OK { return Err(ctx.err(rc)); }
        if cap != expected_cap {
            unsafe { p2g_circuit_free(ctx.raw, raw) };
            return Err(GpuError::Backend(P2G_E_BADARG, "constants_sigmas cap differs between build() and the device".into()));
        }
        Ok(Self { raw, ctx: ctx.raw, desc })
    }
    /// `wires`: num_wires columns of n canonical u64, column-major (`MatrixWitness::wire_values`).
    /// Returns the flat proof words.
    pub fn prove(&self, ctx: &GpuContext, wires: &[u64], public_inputs: &[u64]) -> Result<Vec<u64>, GpuError> {
        assert!(ctx.raw == self.ctx, "a circuit is bound to the context it was loaded on");
        unsafe {
            let cap = p2g_proof_words(self.raw);
            let mut out = vec![0u64; cap];
            let mut n = 0usize;
            let pi = if public_inputs.is_empty() { ptr::null() } else { public_inputs.as_ptr() };
            let rc = p2g_prove(ctx.raw, self.raw, wires.as_ptr(), pi, out.as_mut_ptr(), cap, &mut n);
            if rc != P2G_OK { return Err(ctx.err(rc)); }
            out.truncate(n);
            Ok(out)
        }
    }
    /// Proof as upstream bytes: `ProofWithPublicInputs::from_bytes(bytes, common_data)` finishes the job.
    pub fn prove_bytes(&self, ctx: &GpuContext, wires: &[u64], public_inputs: &[u64]) -> Result<Vec<u8>, GpuError> {
        let words = self.prove(ctx, wires, public_inputs)?;
        self.desc.proof_words_to_bytes(&words)
    }
    /// The FRI phase on its own (`PolynomialBatch::prove_openings` + `fri_proof`): `wires` / `zs` / `quotient` are raw
    /// batch handles of this circuit (`p2g_commit_from_values*`, `p2g_quotient`), `zeta` the evaluation point and
    /// `challenger` the caller's `Challenger` state after it observed the openings, 30 words
    /// (12 sponge, 8 input buffer, 8 output buffer, input length, output length); it is updated in place.
    /// Returns the FriProof part of the flat proof words.
    pub fn fri_prove(&self, ctx: &GpuContext, wires: *const p2g_batch, zs: *const p2g_batch, quotient: *const p2g_batch,
                     zeta: [u64; 2], challenger: &mut [u64; 30]) -> Result<Vec<u64>, GpuError> {
        assert!(ctx.raw == self.ctx, "a circuit is bound to the context it was loaded on");
        unsafe {
            let cap = p2g_fri_proof_words(self.raw);
            let mut out = vec![0u64; cap];
            let mut n = 0usize;
            let rc = p2g_fri_prove(ctx.raw, self.raw, wires, zs, quotient, zeta.as_ptr(), challenger.as_mut_ptr(),
                                   out.as_mut_ptr(), cap, &mut n);
            if rc != P2G_OK { return Err(ctx.err(rc)); }
            out.truncate(n);
            Ok(out)
        }
    }
    pub fn raw(&self) -> *mut p2g_circuit { self.raw }
}
impl Drop for GpuCircuit { fn drop(&mut self) { unsafe { p2g_circuit_free(self.ctx, self.raw); } } }
