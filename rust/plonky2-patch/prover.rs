// Patch for plonky2/src/plonk/prover.rs of the fork pinned at /root/reference/Cargo.toml:12
// (0xPARC/plonky2 @ 109d517).  NOT compiled here (no rustc); shown as the drop-in a maintainer
// applies.  Everything before the witness matrix and everything in verify() is unchanged upstream
// code; only the body between `full_witness()` and the returned proof is replaced by the GPU call.
//
// pub fn prove_with_partition_witness<F, C, const D: usize>(
//     prover_data: &ProverOnlyCircuitData<F, C, D>,
//     common_data: &CommonCircuitData<F, D>,
//     mut partition_witness: PartitionWitness<F>,
//     timing: &mut TimingTree,
// ) -> Result<ProofWithPublicInputs<F, C, D>>
// {
//     let has_lookup = !common_data.luts.is_empty();
//     if has_lookup { set_lookup_wires(prover_data, common_data, &mut partition_witness)?; }   // host
//     let public_inputs = partition_witness.get_targets(&prover_data.public_inputs);
//     let witness = partition_witness.full_witness();                                           // host
//
//     // ---- GPU hot path --------------------------------------------------------------------
//     let wires: Vec<u64> = witness.wire_values.iter()
//         .flat_map(|col| col.iter().map(|x| x.to_canonical_u64())).collect();                 // [135][n]
//     let pis: Vec<u64> = public_inputs.iter().map(|x| x.to_canonical_u64()).collect();
//     let gpu = prover_data.gpu.get_or_try_init(|| GpuCircuit::load(prover_data, common_data))?; // once
//     match gpu.lock().prove(&wires, &pis) {
//         Ok(words) => Ok(proof_from_words::<F, C, D>(&words, common_data, public_inputs)),
//         Err(GpuError::Unsatisfied) => Err(anyhow!("witness does not satisfy the circuit")),
//         Err(GpuError::Backend(code, msg)) if code == P2G_E_BADARG =>
//             // a gate the backend does not evaluate (pod2 custom gates): stock CPU prover
//             cpu::prove_with_partition_witness(prover_data, common_data, partition_witness, timing),
//         Err(GpuError::Backend(code, msg)) => Err(anyhow!("p2gpu error {code}: {msg}")),
//     }
// }
//
// GpuCircuit::load fills p2g_circuit_desc from:
//   common_data.config, .fri_params.reduction_arity_bits, .degree_bits(), .gates (kind by gate id,
//   selectors_info.selector_indices / groups, num_constraints), .num_gate_constraints, .num_constants,
//   .num_partial_products, .k_is, .num_lookup_selectors, .luts;
//   prover_data.constants_sigmas_commitment.polynomials (coefficients -> values with one FFT each),
//   prover_data.lookup_rows, prover_data.circuit_digest;
//   expected cap = verifier_only.constants_sigmas_cap.
