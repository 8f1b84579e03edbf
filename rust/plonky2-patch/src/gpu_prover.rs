//! Patch for the plonky2 fork pinned at /root/reference/Cargo.toml:12 (0xPARC/plonky2 @ 109d517): the body of
//! `plonk::prover::prove_with_partition_witness` between `full_witness()` and the returned proof is replaced
//! by one call into libp2gpu.so.  Add this file as `plonky2/src/plonk/gpu_prover.rs`, `pub mod gpu_prover;`
//! in `plonk/mod.rs`, `p2gpu-sys = { path = ".../rust/p2gpu-sys" }` in plonky2's Cargo.toml, and make
//! `prove_with_partition_witness` start with `if let Some(r) = gpu_prover::try_prove(..) { return r; }`
//!.  The gadget crates, `CircuitBuilder`, `prove`, `verify` keep their signatures:
//! `data.prove(pw)` (aes-gcm/examples/aes_gcm_128.rs:52) lands here unchanged.
//!
//! NOT compiled in the build image (no cargo / rustc).  Written against the upstream API at the pinned rev.
use std::collections::HashMap;
use std::sync::{Mutex, OnceLock};

use anyhow::{anyhow, Result};
use p2gpu_sys::{CircuitDescOwned, GpuCircuit, GpuContext, GpuError, p2g_gate};

use crate::field::extension::Extendable;
use crate::field::types::PrimeField64;
use crate::hash::hash_types::RichField;
use crate::iop::witness::{PartitionWitness, WitnessWrite};
use crate::plonk::circuit_data::{CommonCircuitData, ProverOnlyCircuitData};
use crate::plonk::config::GenericConfig;
use crate::plonk::proof::ProofWithPublicInputs;
use crate::plonk::prover::set_lookup_wires;

/// include/p2gpu.h gate kinds; `None` = a gate the backend does not evaluate (the caller falls through to
/// the stock CPU prover, e.g. for pod2's NNF / ecGFp5 gates)
fn gate_kind(id: &str) -> Option<(i32, i32)> {
    let param = |key: &str| -> i32 {
        id.split(key).nth(1).map(|s| s.chars().take_while(|c| c.is_ascii_digit()).collect::<String>())
            .and_then(|s| s.parse().ok()).unwrap_or(0)
    };
    if id.starts_with("NoopGate") { Some((p2gpu_sys::P2G_GATE_NOOP, 0)) }
    else if id.starts_with("ConstantGate") { Some((p2gpu_sys::P2G_GATE_CONSTANT, param("num_consts: "))) }
    else if id.starts_with("PublicInputGate") { Some((p2gpu_sys::P2G_GATE_PUBLIC_INPUT, 0)) }
    else if id.starts_with("ArithmeticGate") { Some((p2gpu_sys::P2G_GATE_ARITHMETIC, param("num_ops: "))) }
    else if id.starts_with("LookupGate") { Some((p2gpu_sys::P2G_GATE_LOOKUP, 0)) }
    else if id.starts_with("LookupTableGate") { Some((p2gpu_sys::P2G_GATE_LOOKUP_TABLE, 0)) }
    else if id.starts_with("PoseidonGate") { Some((p2gpu_sys::P2G_GATE_POSEIDON, 0)) }
    else { None }
}

/// Fills the descriptor from what `CircuitBuilder::build` left in the circuit data.  `None` when the circuit
/// is outside what the backend supports (zero knowledge, other presets, unknown gates).
pub fn describe<F, C, const D: usize>(
    prover_data: &ProverOnlyCircuitData<F, C, D>,
    common: &CommonCircuitData<F, D>,
) -> Option<CircuitDescOwned>
where F: RichField + Extendable<D> + PrimeField64, C: GenericConfig<D, F = F> {
    let cfg = &common.config;
    if cfg.zero_knowledge || D != 2 || cfg.fri_config.rate_bits != 3 || common.quotient_degree_factor != 8 { return None; }
    let sel = &common.selectors_info;
    let mut gates = Vec::with_capacity(common.gates.len());
    for (i, g) in common.gates.iter().enumerate() {
        let (kind, param0) = gate_kind(&g.0.id())?;
        let s = sel.selector_indices[i];
        let grp = &sel.groups[s];
        gates.push(p2g_gate { kind, selector_index: s as i32, group_start: grp.start as i32, group_end: grp.end as i32,
                              num_constraints: g.0.num_constraints() as i32, param0 });
    }
    // the preprocessed polynomials are kept as coefficients; the backend takes their values on the subgroup
    let n = common.degree();
    let mut constants_sigmas = Vec::with_capacity(prover_data.constants_sigmas_commitment.polynomials.len() * n);
    for p in &prover_data.constants_sigmas_commitment.polynomials {
        constants_sigmas.extend(p.clone().fft().values.iter().map(|x| x.to_canonical_u64()));
    }
    Some(CircuitDescOwned {
        degree_bits: common.degree_bits() as i32, num_wires: cfg.num_wires as i32, num_routed_wires: cfg.num_routed_wires as i32,
        num_constants: cfg.num_constants as i32, num_challenges: cfg.num_challenges as i32,
        quotient_degree_factor: common.quotient_degree_factor as i32, rate_bits: cfg.fri_config.rate_bits as i32,
        cap_height: cfg.fri_config.cap_height as i32, pow_bits: cfg.fri_config.proof_of_work_bits as i32,
        num_query_rounds: cfg.fri_config.num_query_rounds as i32,
        reduction_arity_bits: common.fri_params.reduction_arity_bits.iter().map(|&x| x as i32).collect(),
        num_selectors: sel.num_selectors() as i32, num_lookup_selectors: common.num_lookup_selectors as i32, gates,
        num_gate_constraints: common.num_gate_constraints as i32, num_partial_products: common.num_partial_products as i32,
        lut_lens: common.luts.iter().map(|l| l.len() as i32).collect(),
        lut_data: common.luts.iter().flat_map(|l| l.iter().flat_map(|&(a, b)| [a, b])).collect(),
        lookup_rows: prover_data.lookup_rows.iter()
            .flat_map(|r| [r.last_lu_gate as i32, r.last_lut_gate as i32, r.first_lut_gate as i32]).collect(),
        num_public_inputs: common.num_public_inputs as i32,
        k_is: common.k_is.iter().map(|x| x.to_canonical_u64()).collect(),
        constants_sigmas,
        circuit_digest: prover_data.circuit_digest.elements.map(|x| x.to_canonical_u64()),
    })
}

/// One context + loaded circuit per (device, circuit digest); `CircuitData` is immutable after build and reused
/// across proofs (aes-gcm/src/circuit_aes.rs:394-410), so the preprocessed commitment is uploaded once.
struct Loaded { ctx: GpuContext, circuit: GpuCircuit }
static CACHE: OnceLock<Mutex<HashMap<[u64; 4], Option<&'static Mutex<Loaded>>>>> = OnceLock::new();

fn loaded<F, C, const D: usize>(
    prover_data: &ProverOnlyCircuitData<F, C, D>,
    common: &CommonCircuitData<F, D>,
) -> Result<Option<&'static Mutex<Loaded>>>
where F: RichField + Extendable<D> + PrimeField64, C: GenericConfig<D, F = F> {
    let key = prover_data.circuit_digest.elements.map(|x| x.to_canonical_u64());
    let mut cache = CACHE.get_or_init(|| Mutex::new(HashMap::new())).lock().unwrap();
    if let Some(e) = cache.get(&key) { return Ok(*e); }
    let entry = match describe(prover_data, common) {
        None => None,
        Some(desc) => {
            let device = std::env::var("P2GPU_DEVICE").ok().and_then(|s| s.parse().ok()).unwrap_or(0);
            let ctx = GpuContext::new(device).map_err(|e| anyhow!("{e}"))?;
            // the cap the CPU build() committed to: the device commitment must reproduce it
            let expected: Vec<u64> = prover_data.constants_sigmas_commitment.merkle_tree.cap.0.iter()
                .flat_map(|h| h.elements.map(|x| x.to_canonical_u64())).collect();
            let circuit = GpuCircuit::load(&ctx, desc, &expected).map_err(|e| anyhow!("{e}"))?;
            Some(&*Box::leak(Box::new(Mutex::new(Loaded { ctx, circuit }))))
        }
    };
    cache.insert(key, entry);
    Ok(entry)
}

/// The GPU path of `prove_with_partition_witness`.  `None`: circuit not supported by the backend, the caller
/// continues with the stock CPU prover.  `Some(Err)`: a backend failure.
pub fn try_prove<F, C, const D: usize>(
    prover_data: &ProverOnlyCircuitData<F, C, D>,
    common: &CommonCircuitData<F, D>,
    partition_witness: &mut PartitionWitness<F>,
) -> Option<Result<ProofWithPublicInputs<F, C, D>>>
where F: RichField + Extendable<D> + PrimeField64, C: GenericConfig<D, F = F> {
    let gpu = match loaded(prover_data, common) {
        Ok(Some(g)) => g,
        Ok(None) => return None,
        Err(e) => return Some(Err(e)),
    };
    Some((|| {
        // host side, unchanged upstream code: lookup multiplicities, public inputs, the wire matrix
        if !common.luts.is_empty() { set_lookup_wires(prover_data, common, partition_witness)?; }
        let public_inputs: Vec<u64> = partition_witness.get_targets(&prover_data.public_inputs)
            .iter().map(|x| x.to_canonical_u64()).collect();
        let witness = partition_witness.clone().full_witness();
        let n = common.degree();
        let mut wires = Vec::with_capacity(common.config.num_wires * n);       // [num_wires][n], column-major
        for col in 0..common.config.num_wires {
            for row in 0..n { wires.push(witness.get_wire(row, col).to_canonical_u64()); }
        }
        // ---- the hot path: LDE + Merkle, Z / lookups, quotient, openings, FRI, PoW, queries ----
        let g = gpu.lock().unwrap();
        let bytes = g.circuit.prove_bytes(&g.ctx, &wires, &public_inputs).map_err(|e| match e {
            GpuError::WitnessConflict => anyhow!("Partition was set twice with different values"),
            e => anyhow!("{e}"),
        })?;
        ProofWithPublicInputs::<F, C, D>::from_bytes(bytes, common)
    })())
}
