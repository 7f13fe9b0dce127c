//! The batched form of the reference's examples/crank-out-pitchypl.rs:139-195: draw a block of
//! points, ONE call for the block, the same TSV columns and `{:.16e}` rows.  (The tested
//! implementation of this tool is rimphony_b200/crank_out.py; this file shows the Rust side.)
use rimphony_b200::{compute_all_dimensionless_batch, Mode};
use std::time::Instant;

fn main() -> Result<(), String> {
    let n: usize = std::env::args().nth(1).and_then(|a| a.parse().ok()).unwrap_or(4096);
    // rimphony_test_support::Sampler (test-support/src/lib.rs:39-63) is used in the real tool;
    // a fixed grid keeps this example dependency-free.
    let s: Vec<f64> = (0..n).map(|i| (0.07f64.ln() + (1e4f64 / 0.07).ln() * (i as f64 + 0.5) / n as f64).exp()).collect();
    let theta: Vec<f64> = (0..n).map(|i| 0.003 + 1.5675 * ((i * 7919) % n) as f64 / n as f64).collect();
    let p: Vec<f64> = (0..n).map(|i| 1.5 + 2.5 * ((i * 104729) % n) as f64 / n as f64).collect();
    let k: Vec<f64> = (0..n).map(|i| 3.0 * ((i * 1299709) % n) as f64 / n as f64).collect();

    let t0 = Instant::now();
    let res = compute_all_dimensionless_batch(2, &s, &theta, &[&p, &k], Mode::Fast)?;
    let ms = t0.elapsed().as_secs_f64() * 1e3 / n as f64;

    println!("s(log)\ttheta(lin)\tp(lin)\tk(lin)\ttime_ms(meta)\tj_I(res)\talpha_I(res)\tj_Q(res)\talpha_Q(res)\tj_V(res)\talpha_V(res)\trho_Q(res)\trho_V(res)");
    for i in 0..n {
        print!("{:.16e}\t{:.16e}\t{:.16e}\t{:.16e}\t{:.16e}", s[i], theta[i], p[i], k[i], ms);
        for v in res.values[i].iter() {
            print!("\t{:.16e}", v);
        }
        println!();
    }
    Ok(())
}
