//! Reference fixture dumper — the pinning path for the CPU oracle and the CUDA prover of this repo.
//!
//! NOT compiled in the build image (no cargo / rustc, the plonky2 fork is not vendored).  A maintainer
//! with the reference's toolchain (nightly-2026-02-15, network access for the git dependencies) runs it
//! ONCE and drops the two files it writes into `tests/golden/`; `tests/test_reference_fixture.py` then
//! feeds the dumped circuit description and wire matrix to the oracle and to `p2g_prove` and compares the
//! proof BYTES with the ones the real `plonky2 @ 109d517` produced.  No code change is needed.
//!
//!   cp tools/dump_reference_fixture.rs /path/to/plonky2-aes/aes-gcm/examples/dump_fixture.rs
//!   cd /path/to/plonky2-aes
//!   RAYON_NUM_THREADS=1 cargo run --release -p plonky2-aes --example dump_fixture -- /tmp/fixtures
//!   cp /tmp/fixtures/reference_*.p2gfix  <this repo>/tests/golden/
//!
//! RAYON_NUM_THREADS=1 matters: `fri_proof_of_work` searches the nonce with a parallel `find_any`; with
//! one worker the range is scanned from 0 upwards, so the witness is the LOWEST valid nonce, which is
//! what this repo's provers return (BASELINE.json north_star).
//!
//! Circuits (the reference's own fixed test inputs):
//!   reference_c1_aes128_block.p2gfix   `test_encrypt_block_test_vector_op::<4,4,10>`, FIPS-197 App. B
//!                                      (aes-gcm/src/circuit_aes.rs:619-726)
//!   reference_c2_aes_gcm_256.p2gfix    `AesGcmTarget::<4,4,10,256,true>`, key=[42;16], nonce=[111;12],
//!                                      pt=[42;256] (aes-gcm/src/circuit_gcm.rs:737-783), BASELINE config 2
//!
//! File format (little endian; read by tests/fixture_format.py):
//!   magic "P2GFIX1\0", u32 section count, then per section:
//!   u32 name length, name bytes, u32 element size in bytes (1, 2, 4 or 8), u64 element count, raw data.
//! Sections: config (i32 x 15), reduction_arity_bits (i32), gates (i32 x 6 per gate: kind,
//! selector_index, group_start, group_end, num_constraints, param0), gate_ids (u8, '\n'-joined),
//! lut_lens (i32), lut_data (u16 pairs), lookup_rows (i32 x 3 per LUT), k_is (u64),
//! constants_sigmas (u64, [columns][n] VALUES on the subgroup, column-major), circuit_digest (u64 x 4),
//! constants_sigmas_cap (u64), wires (u64, [num_wires][n] column-major), public_inputs (u64),
//! proof_bytes (u8, `ProofWithPublicInputs::to_bytes`).
#![allow(incomplete_features)]
#![feature(generic_const_exprs)]

use std::fs::File;
use std::io::Write;

use anyhow::Result;
use plonky2::field::goldilocks_field::GoldilocksField as F;
use plonky2::field::types::PrimeField64;
use plonky2::iop::generator::generate_partial_witness;
use plonky2::iop::witness::PartialWitness;
use plonky2::plonk::circuit_builder::CircuitBuilder;
use plonky2::plonk::circuit_data::{CircuitConfig, CircuitData};
use plonky2::plonk::config::PoseidonGoldilocksConfig as C;
use plonky2::plonk::prover::{prove_with_partition_witness, set_lookup_wires};
use plonky2::util::timing::TimingTree;
use plonky2_aes::circuit_aes::{
    byte_xor_lut, gf_2_8_mul_lut, sbox_lut, state_mix_matrix, ByteTarget, CircuitBuilderAESState,
    PartialWitnessAESState, PartialWitnessByteArray,
};
use plonky2_aes::native_aes::{encrypt_block, key_expansion};
use plonky2_aes::{AesGcmTarget, D};

struct Fixture { sections: Vec<(String, u32, u64, Vec<u8>)> }
impl Fixture {
    fn new() -> Self { Self { sections: vec![] } }
    fn i32s(&mut self, name: &str, v: &[i32]) {
        self.sections.push((name.into(), 4, v.len() as u64, v.iter().flat_map(|x| x.to_le_bytes()).collect()));
    }
    fn u16s(&mut self, name: &str, v: &[u16]) {
        self.sections.push((name.into(), 2, v.len() as u64, v.iter().flat_map(|x| x.to_le_bytes()).collect()));
    }
    fn u64s(&mut self, name: &str, v: &[u64]) {
        self.sections.push((name.into(), 8, v.len() as u64, v.iter().flat_map(|x| x.to_le_bytes()).collect()));
    }
    fn bytes(&mut self, name: &str, v: &[u8]) { self.sections.push((name.into(), 1, v.len() as u64, v.to_vec())); }
    fn write(&self, path: &str) -> Result<()> {
        let mut f = File::create(path)?;
        f.write_all(b"P2GFIX1\0")?;
        f.write_all(&(self.sections.len() as u32).to_le_bytes())?;
        for (name, esz, cnt, data) in &self.sections {
            f.write_all(&(name.len() as u32).to_le_bytes())?;
            f.write_all(name.as_bytes())?;
            f.write_all(&esz.to_le_bytes())?;
            f.write_all(&cnt.to_le_bytes())?;
            f.write_all(data)?;
        }
        Ok(())
    }
}

/// number after `key` in a gate id such as "ArithmeticGate { num_ops: 20 }"
fn id_param(id: &str, key: &str) -> i32 {
    id.split(key).nth(1).map(|s| s.chars().take_while(|c| c.is_ascii_digit()).collect::<String>())
        .and_then(|s| s.parse().ok()).unwrap_or(0)
}

/// include/p2gpu.h gate kinds; -1 = a gate the backend does not evaluate
fn gate_kind(id: &str) -> (i32, i32) {
    if id.starts_with("NoopGate") { (0, 0) }
    else if id.starts_with("ConstantGate") { (1, id_param(id, "num_consts: ")) }
    else if id.starts_with("PublicInputGate") { (2, 0) }
    else if id.starts_with("ArithmeticGate") { (3, id_param(id, "num_ops: ")) }
    else if id.starts_with("LookupGate") { (4, 0) }
    else if id.starts_with("LookupTableGate") { (5, 0) }
    else if id.starts_with("PoseidonGate") { (6, 0) }
    else { (-1, 0) }
}

fn dump(data: &CircuitData<F, C, D>, pw: PartialWitness<F>, path: &str) -> Result<()> {
    let common = &data.common;
    let cfg = &common.config;
    let n = common.degree();
    let mut fx = Fixture::new();
    let sel = &common.selectors_info;
    fx.i32s("config", &[
        common.degree_bits() as i32, cfg.num_wires as i32, cfg.num_routed_wires as i32, cfg.num_constants as i32,
        cfg.num_challenges as i32, common.quotient_degree_factor as i32, cfg.fri_config.rate_bits as i32,
        cfg.fri_config.cap_height as i32, cfg.fri_config.proof_of_work_bits as i32, cfg.fri_config.num_query_rounds as i32,
        sel.num_selectors() as i32, common.num_lookup_selectors as i32, common.num_gate_constraints as i32,
        common.num_partial_products as i32, common.num_public_inputs as i32,
    ]);
    fx.i32s("reduction_arity_bits", &common.fri_params.reduction_arity_bits.iter().map(|&x| x as i32).collect::<Vec<_>>());
    let mut gates = vec![];
    let mut ids = String::new();
    for (i, g) in common.gates.iter().enumerate() {
        let id = g.0.id();
        let (kind, param) = gate_kind(&id);
        let s = sel.selector_indices[i];
        let grp = &sel.groups[s];
        gates.extend_from_slice(&[kind, s as i32, grp.start as i32, grp.end as i32, g.0.num_constraints() as i32, param]);
        ids.push_str(&id);
        ids.push('\n');
    }
    fx.i32s("gates", &gates);
    fx.bytes("gate_ids", ids.as_bytes());
    fx.i32s("lut_lens", &common.luts.iter().map(|l| l.len() as i32).collect::<Vec<_>>());
    fx.u16s("lut_data", &common.luts.iter().flat_map(|l| l.iter().flat_map(|&(a, b)| [a, b])).collect::<Vec<_>>());
    fx.i32s("lookup_rows", &data.prover_only.lookup_rows.iter()
        .flat_map(|r| [r.last_lu_gate as i32, r.last_lut_gate as i32, r.first_lut_gate as i32]).collect::<Vec<_>>());
    fx.u64s("k_is", &common.k_is.iter().map(|x| x.to_canonical_u64()).collect::<Vec<_>>());
    // constants_sigmas: the committed polynomials are stored as coefficients; the backend takes VALUES on H
    let mut cs = vec![];
    for p in &data.prover_only.constants_sigmas_commitment.polynomials {
        let vals = p.clone().fft();
        assert_eq!(vals.values.len(), n);
        cs.extend(vals.values.iter().map(|x| x.to_canonical_u64()));
    }
    fx.u64s("constants_sigmas", &cs);
    fx.u64s("circuit_digest", &data.verifier_only.circuit_digest.elements.map(|x| x.to_canonical_u64()));
    fx.u64s("constants_sigmas_cap", &data.verifier_only.constants_sigmas_cap.0.iter()
        .flat_map(|h| h.elements.map(|x| x.to_canonical_u64())).collect::<Vec<_>>());

    // the witness exactly as prove() builds it: generators, set_lookup_wires, full_witness
    let partition = generate_partial_witness(pw, &data.prover_only, common)?;
    let mut filled = partition.clone();
    if !common.luts.is_empty() {
        set_lookup_wires(&data.prover_only, common, &mut filled)?;
    }
    let public_inputs: Vec<u64> = filled.get_targets(&data.prover_only.public_inputs).iter().map(|x| x.to_canonical_u64()).collect();
    let matrix = filled.full_witness();
    let mut wires = Vec::with_capacity(cfg.num_wires * n);
    for col in 0..cfg.num_wires {
        for row in 0..n {
            wires.push(matrix.get_wire(row, col).to_canonical_u64());
        }
    }
    fx.u64s("wires", &wires);
    fx.u64s("public_inputs", &public_inputs);

    let mut timing = TimingTree::default();
    let proof = prove_with_partition_witness(&data.prover_only, common, partition, &mut timing)?;
    data.verify(proof.clone())?;
    fx.bytes("proof_bytes", &proof.to_bytes());
    fx.write(path)?;
    println!("wrote {path}: n = {n}, {} gates kinds, proof {} bytes, pow_witness {}", common.gates.len(),
             proof.to_bytes().len(), proof.proof.opening_proof.pow_witness.to_canonical_u64());
    Ok(())
}

/// C1: circuit_aes.rs:657-726 with the AES-128 vector of circuit_aes.rs:621-630
fn c1() -> Result<(CircuitData<F, C, D>, PartialWitness<F>)> {
    const NK: usize = 4; const NB: usize = 4; const NR: usize = 10;
    let input_state: [u8; 16] = [0x32, 0x43, 0xf6, 0xa8, 0x88, 0x5a, 0x30, 0x8d, 0x31, 0x31, 0x98, 0xa2, 0xe0, 0x37, 0x07, 0x34];
    let key: [u8; 16] = [0x2b, 0x7e, 0x15, 0x16, 0x28, 0xae, 0xd2, 0xa6, 0xab, 0xf7, 0x15, 0x88, 0x09, 0xcf, 0x4f, 0x3c];
    let mut builder = CircuitBuilder::<F, D>::new(CircuitConfig::standard_recursion_config());
    let key_target: [ByteTarget; NK * NB] = std::array::from_fn(|_| builder.add_virtual_byte_target_unsafe());
    let xor_lut_idx = byte_xor_lut(&mut builder);
    let gf_2_8_mul_lut_idx = gf_2_8_mul_lut(&mut builder);
    let sbox_lut_idx = sbox_lut(&mut builder);
    let mix_matrix = state_mix_matrix(&mut builder);
    let expanded_key_target: [[ByteTarget; 4]; 4 * (NR + 1)] =
        builder.key_expansion::<NK, NB, NR>(xor_lut_idx, sbox_lut_idx, key_target);
    let input_state_target = builder.add_virtual_state_target(sbox_lut_idx);
    let output = builder.encrypt_block(xor_lut_idx, gf_2_8_mul_lut_idx, sbox_lut_idx, mix_matrix, input_state_target, expanded_key_target);
    let data = builder.build::<C>();
    let mut m = [[0u8; 4]; 4];
    for i in 0..4 { for j in 0..4 { m[i][j] = input_state[i + 4 * j]; } }
    let expanded_key: [[u8; 4]; 4 * (NR + 1)] = key_expansion::<NK, NB, NR>(&key);
    let ct = encrypt_block::<NR>(&input_state, &expanded_key);
    let mut pw = PartialWitness::<F>::new();
    std::iter::zip(key_target, key).try_for_each(|(t, v)| pw.set_byte_target(t, v))?;
    pw.set_state_target(input_state_target, m)?;
    std::iter::zip(expanded_key_target, expanded_key).try_for_each(|(t, v)| std::iter::zip(t, v).try_for_each(|(t, v)| pw.set_byte_target(t, v)))?;
    std::iter::zip(output.0, ct).try_for_each(|(o, e)| std::iter::zip(o, e).try_for_each(|(o, e)| pw.set_byte_target(o, e)))?;
    Ok((data, pw))
}

/// C2: circuit_gcm.rs:737-783 with L = 256, TAG = true
fn c2() -> Result<(CircuitData<F, C, D>, PartialWitness<F>)> {
    const L: usize = 256;
    let key: &[u8; 16] = &[42; 16];
    let nonce: &[u8; 12] = &[111; 12];
    let pt: &[u8; L] = &[42u8; L];
    let (ct, tag) = plonky2_aes::native_gcm::encrypt::<4, 4, 10>(key, nonce, pt);
    let mut builder = CircuitBuilder::<F, D>::new(CircuitConfig::standard_recursion_config());
    let targets = AesGcmTarget::<4, 4, 10, L, true>::build(&mut builder);
    let data = builder.build::<C>();
    let mut pw = PartialWitness::<F>::new();
    targets.set_targets(&mut pw, key, nonce, pt, &ct, &tag)?;
    Ok((data, pw))
}

fn main() -> Result<()> {
    let dir = std::env::args().nth(1).unwrap_or_else(|| ".".into());
    std::fs::create_dir_all(&dir)?;
    if std::env::var("RAYON_NUM_THREADS").map(|v| v != "1").unwrap_or(true) {
        eprintln!("warning: RAYON_NUM_THREADS != 1 -- the proof-of-work witness may not be the lowest nonce");
    }
    let (d1, pw1) = c1()?;
    dump(&d1, pw1, &format!("{dir}/reference_c1_aes128_block.p2gfix"))?;
    let (d2, pw2) = c2()?;
    dump(&d2, pw2, &format!("{dir}/reference_c2_aes_gcm_256.p2gfix"))?;
    Ok(())
}
