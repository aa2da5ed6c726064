//! Golden-vector generator for tests/test_ark_fixture.py (TEST INFRASTRUCTURE; nothing in the product depends on it).
//!
//! For every (curve, circuit) it runs arkworks' own setup and prover with a seeded RNG and writes one directory
//!     <out>/<curve>_<circuit>/{meta.json, pk.bin, vk.bin, matrices.bin, witness.bin, rs.bin, public.bin, proof.bin,
//!                              proof_uncompressed.bin, h.bin}
//! in the formats groth16_b200/serialize.py reads:
//!   pk.bin                  ProvingKey<E>::serialize_uncompressed        (data_structures.rs:125 derive)
//!   vk.bin                  VerifyingKey<E>::serialize_compressed         (data_structures.rs:31)
//!   proof.bin               Proof<E>::serialize_compressed               (data_structures.rs:8)
//!   proof_uncompressed.bin  Proof<E>::serialize_uncompressed
//!   witness.bin / public.bin / rs.bin / h.bin    Vec<Fr> (u64 LE length, then 32-byte LE canonical scalars); rs = [r, s];
//!                           h = LibsnarkReduction::witness_map_from_matrices output (pins the NTT path separately)
//!   matrices.bin            u64 num_instance, u64 num_witness, u64 num_constraints, then for a, b, c: for every row
//!                           u64 len, then len x (32-byte LE canonical coefficient, u64 column)
//! `r` is sampled before `s` (src/prover.rs:146-147).  Circuits: MySillyCircuit (src/test.rs:14-43), the MiMC demo of
//! tests/mimc.rs:65-143 (322 rounds) and a non-degenerate synthetic circuit (z_p + k) * z_q = z_new.
use ark_ec::pairing::Pairing;
use ark_ff::{Field, PrimeField, UniformRand};
use ark_groth16::{r1cs_to_qap::{LibsnarkReduction, R1CSToQAP}, Groth16};
use ark_poly::GeneralEvaluationDomain;
use ark_relations::{
    lc,
    r1cs::{ConstraintMatrices, ConstraintSynthesizer, ConstraintSystem, ConstraintSystemRef, OptimizationGoal, SynthesisError, Variable},
};
use ark_serialize::CanonicalSerialize;
use ark_crypto_primitives::snark::SNARK; // the trait ark-groth16 implements (src/lib.rs:47,59)
use ark_std::rand::{rngs::StdRng, Rng, SeedableRng};
use std::{fs, io::Write, path::Path};

// ---- circuits ----------------------------------------------------------------------------------------------------
#[derive(Clone)]
struct Silly<F: Field> { a: Option<F>, b: Option<F> }
impl<F: Field> ConstraintSynthesizer<F> for Silly<F> {
    fn generate_constraints(self, cs: ConstraintSystemRef<F>) -> Result<(), SynthesisError> {
        let a = cs.new_witness_variable(|| self.a.ok_or(SynthesisError::AssignmentMissing))?;
        let b = cs.new_witness_variable(|| self.b.ok_or(SynthesisError::AssignmentMissing))?;
        let c = cs.new_input_variable(|| Ok(self.a.ok_or(SynthesisError::AssignmentMissing)? * self.b.ok_or(SynthesisError::AssignmentMissing)?))?;
        for _ in 0..6 { cs.enforce_constraint(lc!() + a, lc!() + b, lc!() + c)?; }
        Ok(())
    }
}

const MIMC_ROUNDS: usize = 322;
#[derive(Clone)]
struct Mimc<F: Field> { xl: Option<F>, xr: Option<F>, constants: Vec<F> }
impl<F: Field> ConstraintSynthesizer<F> for Mimc<F> {
    fn generate_constraints(self, cs: ConstraintSystemRef<F>) -> Result<(), SynthesisError> {
        let (mut xl_v, mut xr_v) = (self.xl, self.xr);
        let mut xl = cs.new_witness_variable(|| xl_v.ok_or(SynthesisError::AssignmentMissing))?;
        let mut xr = cs.new_witness_variable(|| xr_v.ok_or(SynthesisError::AssignmentMissing))?;
        for i in 0..MIMC_ROUNDS {
            let k = self.constants[i];
            let tmp_v = xl_v.map(|e| (e + k).square());
            let tmp = cs.new_witness_variable(|| tmp_v.ok_or(SynthesisError::AssignmentMissing))?;
            cs.enforce_constraint(lc!() + xl + (k, Variable::One), lc!() + xl + (k, Variable::One), lc!() + tmp)?;
            let new_v = xl_v.map(|e| (e + k) * tmp_v.unwrap() + xr_v.unwrap());
            let new_xl = if i == MIMC_ROUNDS - 1 {
                cs.new_input_variable(|| new_v.ok_or(SynthesisError::AssignmentMissing))?
            } else {
                cs.new_witness_variable(|| new_v.ok_or(SynthesisError::AssignmentMissing))?
            };
            cs.enforce_constraint(lc!() + tmp, lc!() + xl + (k, Variable::One), lc!() + new_xl - xr)?;
            xr = xl; xr_v = xl_v; xl = new_xl; xl_v = new_v;
        }
        Ok(())
    }
}

/// constraint i: (z_p + k_i) * z_q = z_new, p and q uniform over earlier witnesses; the last product is the public input
#[derive(Clone)]
struct Synthetic<F: Field> { seeds: Option<(F, F)>, ks: Vec<F>, ps: Vec<usize>, qs: Vec<usize> }
impl<F: Field> ConstraintSynthesizer<F> for Synthetic<F> {
    fn generate_constraints(self, cs: ConstraintSystemRef<F>) -> Result<(), SynthesisError> {
        let n = self.ks.len();
        let mut vars = Vec::with_capacity(n + 2);
        let mut vals: Vec<Option<F>> = Vec::with_capacity(n + 2);
        for j in 0..2 {
            let v = self.seeds.map(|s| if j == 0 { s.0 } else { s.1 });
            vars.push(cs.new_witness_variable(|| v.ok_or(SynthesisError::AssignmentMissing))?);
            vals.push(v);
        }
        for i in 0..n {
            let (p, q, k) = (self.ps[i], self.qs[i], self.ks[i]);
            let v = vals[p].and_then(|a| vals[q].map(|b| (a + k) * b));
            let nv = if i == n - 1 {
                cs.new_input_variable(|| v.ok_or(SynthesisError::AssignmentMissing))?
            } else {
                cs.new_witness_variable(|| v.ok_or(SynthesisError::AssignmentMissing))?
            };
            cs.enforce_constraint(lc!() + vars[p] + (k, Variable::One), lc!() + vars[q], lc!() + nv)?;
            vars.push(nv);
            vals.push(v);
        }
        Ok(())
    }
}

// ---- writers -----------------------------------------------------------------------------------------------------
fn write_scalars<F: PrimeField>(path: &Path, xs: &[F]) {
    let mut f = fs::File::create(path).unwrap();
    xs.to_vec().serialize_compressed(&mut f).unwrap(); // Vec<F>: u64 LE length + canonical LE scalars
}
fn write_matrices<F: PrimeField>(path: &Path, m: &ConstraintMatrices<F>) {
    let mut f = fs::File::create(path).unwrap();
    for v in [m.num_instance_variables, m.num_witness_variables, m.num_constraints] {
        f.write_all(&(v as u64).to_le_bytes()).unwrap();
    }
    for mat in [&m.a, &m.b, &m.c] {
        for row in mat.iter() {
            f.write_all(&(row.len() as u64).to_le_bytes()).unwrap();
            for (coeff, col) in row {
                coeff.serialize_compressed(&mut f).unwrap();
                f.write_all(&(*col as u64).to_le_bytes()).unwrap();
            }
        }
    }
}

fn run<E: Pairing, C: ConstraintSynthesizer<E::ScalarField> + Clone>(out: &Path, curve: &str, name: &str, circuit: C, seed: u64)
where
    E::ScalarField: PrimeField,
{
    let dir = out.join(format!("{curve}_{name}"));
    fs::create_dir_all(&dir).unwrap();
    let mut rng = StdRng::seed_from_u64(seed);
    let (pk, vk) = Groth16::<E>::circuit_specific_setup(circuit.clone(), &mut rng).unwrap();
    let r = E::ScalarField::rand(&mut rng); // r before s: src/prover.rs:146-147
    let s = E::ScalarField::rand(&mut rng);
    let proof = Groth16::<E>::create_proof_with_reduction(circuit.clone(), &pk, r, s).unwrap();
    // the same synthesis the prover does (src/prover.rs:185-204) to export matrices and assignment
    let cs = ConstraintSystem::<E::ScalarField>::new_ref();
    cs.set_optimization_goal(OptimizationGoal::Constraints);
    circuit.generate_constraints(cs.clone()).unwrap();
    cs.finalize();
    let matrices = cs.to_matrices().unwrap();
    let inner = cs.borrow().unwrap();
    let full: Vec<E::ScalarField> = [inner.instance_assignment.as_slice(), inner.witness_assignment.as_slice()].concat();
    let public = inner.instance_assignment[1..].to_vec();
    assert!(Groth16::<E>::verify(&vk, &public, &proof).unwrap());
    let h = LibsnarkReduction::witness_map_from_matrices::<E::ScalarField, GeneralEvaluationDomain<E::ScalarField>>(
        &matrices, matrices.num_instance_variables, matrices.num_constraints, &full).unwrap();
    pk.serialize_uncompressed(fs::File::create(dir.join("pk.bin")).unwrap()).unwrap();
    vk.serialize_compressed(fs::File::create(dir.join("vk.bin")).unwrap()).unwrap();
    proof.serialize_compressed(fs::File::create(dir.join("proof.bin")).unwrap()).unwrap();
    proof.serialize_uncompressed(fs::File::create(dir.join("proof_uncompressed.bin")).unwrap()).unwrap();
    write_matrices(&dir.join("matrices.bin"), &matrices);
    write_scalars(&dir.join("witness.bin"), &full);
    write_scalars(&dir.join("public.bin"), &public);
    write_scalars(&dir.join("rs.bin"), &[r, s]);
    write_scalars(&dir.join("h.bin"), &h);
    fs::write(
        dir.join("meta.json"),
        format!(
            "{{\"producer\": \"ark-groth16 0.5.0 (oracle/ark_fixture)\", \"curve\": \"{curve}\", \"circuit\": \"{name}\", \"seed\": {seed}, \
             \"num_instance_variables\": {}, \"num_witness_variables\": {}, \"num_constraints\": {}, \"pk\": \"uncompressed\", \
             \"proof\": \"compressed\"}}\n",
            matrices.num_instance_variables, matrices.num_witness_variables, matrices.num_constraints
        ),
    )
    .unwrap();
    println!("wrote {}", dir.display());
}

fn synthetic<F: PrimeField>(rng: &mut StdRng, n: usize) -> Synthetic<F> {
    let ks = (0..n).map(|_| F::rand(rng)).collect();
    let ps = (0..n).map(|i| rng.gen_range(0..i + 2)).collect();
    let qs = (0..n).map(|i| rng.gen_range(0..i + 2)).collect();
    Synthetic { seeds: Some((F::rand(rng), F::rand(rng))), ks, ps, qs }
}

fn all<E: Pairing>(out: &Path, curve: &str)
where
    E::ScalarField: PrimeField,
{
    let mut rng = StdRng::seed_from_u64(42);
    type Fr<E> = <E as Pairing>::ScalarField;
    run::<E, _>(out, curve, "silly", Silly { a: Some(Fr::<E>::rand(&mut rng)), b: Some(Fr::<E>::rand(&mut rng)) }, 1);
    let constants: Vec<Fr<E>> = (0..MIMC_ROUNDS).map(|_| Fr::<E>::rand(&mut rng)).collect();
    run::<E, _>(out, curve, "mimc", Mimc { xl: Some(Fr::<E>::rand(&mut rng)), xr: Some(Fr::<E>::rand(&mut rng)), constants }, 2);
    run::<E, _>(out, curve, "synthetic_2p10", synthetic::<Fr<E>>(&mut rng, (1 << 10) - 2), 3);
}

fn main() {
    let out = std::env::args().nth(1).unwrap_or_else(|| "../../tests/golden/ark".to_string());
    let out = Path::new(&out);
    all::<ark_bls12_381::Bls12_381>(out, "bls12_381");
    all::<ark_bn254::Bn254>(out, "bn254");
    all::<ark_bls12_377::Bls12_377>(out, "bls12_377"); // tests/mimc.rs:21 runs MiMC on this curve
}
