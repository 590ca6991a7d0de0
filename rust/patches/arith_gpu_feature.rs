// arith_gpu_feature.rs -- the `#[cfg(feature = "gpu")]` bodies a maintainer adds to crate `arith` so that the
// UNCHANGED public API (Rq / Tn / NTT, operator overloading) runs on libfhe_b200.  Each function names the
// reference site it replaces.  Not compiled in this repository (no Rust toolchain in the build image); the same
// calls, in the same order, are exercised by include/fhe_b200.hpp + tests/cpp/test_host_api.cpp.
//
// Cargo.toml of `arith`:
//   [features]
//   gpu = ["dep:fhe-b200-sys"]
//   [dependencies]
//   fhe-b200-sys = { path = "../fhe-b200-sys", optional = true }

#[cfg(feature = "gpu")]
mod gpu {
    use crate::{ring::RingParam, ring_nq::Rq, ring_torus::Tn, torus::T64, zq::Zq};
    use fhe_b200_sys as sys;
    use std::{collections::HashMap, sync::{Mutex, OnceLock}};

    /// (q, n) -> plan, the GPU twin of the reference's CACHE (arith/src/ntt.rs:18-25).
    fn plan(p: &RingParam) -> *mut sys::FheNttPlan {
        static PLANS: OnceLock<Mutex<HashMap<(u64, usize), usize>>> = OnceLock::new();
        let mut m = PLANS.get_or_init(Default::default).lock().unwrap();
        *m.entry((p.q, p.n)).or_insert_with(|| {
            let mut h = std::ptr::null_mut();
            sys::check(unsafe { sys::fhe_ntt_plan_create(p.q, p.n as u64, &mut h) }); // panics like ntt.rs:116-130
            h as usize
        }) as *mut sys::FheNttPlan
    }
    /// Vec<Zq> is AoS {q, v} (16 B): gather v.
    fn vals(c: &[Zq]) -> Vec<u64> { c.iter().map(|z| z.v).collect() }
    fn zqs(q: u64, v: Vec<u64>) -> Vec<Zq> { v.into_iter().map(|v| Zq { q, v }).collect() }

    /// replaces NTT::ntt (arith/src/ntt.rs:44-73)
    pub fn ntt(a: &Rq) -> Rq {
        let (src, mut out) = (vals(&a.coeffs), vec![0u64; a.param.n]);
        sys::check(unsafe { sys::fhe_ntt_fwd(plan(&a.param), src.as_ptr(), out.as_mut_ptr(), 1) });
        Rq { param: a.param, coeffs: zqs(a.param.q, out), evals: None }
    }
    /// replaces NTT::intt (arith/src/ntt.rs:78-110)
    pub fn intt(a: &Rq) -> Rq {
        let (src, mut out) = (vals(&a.coeffs), vec![0u64; a.param.n]);
        sys::check(unsafe { sys::fhe_ntt_inv(plan(&a.param), src.as_ptr(), out.as_mut_ptr(), 1) });
        Rq { param: a.param, coeffs: zqs(a.param.q, out), evals: None }
    }
    /// replaces ring_nq::mul (arith/src/ring_nq.rs:586-607): reuses cached evals, returns the product with its evals
    pub fn mul(lhs: &Rq, rhs: &Rq) -> Rq {
        assert_eq!(lhs.param, rhs.param);
        let n = lhs.param.n;
        let a = vals(lhs.evals.as_ref().unwrap_or(&lhs.coeffs));
        let b = vals(rhs.evals.as_ref().unwrap_or(&rhs.coeffs));
        let flags = (lhs.evals.is_some() as i32) * sys::FHE_A_IS_EVALS + (rhs.evals.is_some() as i32) * sys::FHE_B_IS_EVALS;
        let (mut c, mut ce) = (vec![0u64; n], vec![0u64; n]);
        sys::check(unsafe { sys::fhe_rq_mul(plan(&lhs.param), a.as_ptr(), b.as_ptr(), c.as_mut_ptr(), 1, flags, ce.as_mut_ptr()) });
        Rq { param: lhs.param, coeffs: zqs(lhs.param.q, c), evals: Some(zqs(lhs.param.q, ce)) }
    }
    /// batched twin added BESIDE the operator API: one call, `a.len()` independent products
    pub fn mul_batch(a: &[Rq], b: &[Rq]) -> Vec<Rq> {
        let p = a[0].param;
        let fa: Vec<u64> = a.iter().flat_map(|x| x.coeffs.iter().map(|z| z.v)).collect();
        let fb: Vec<u64> = b.iter().flat_map(|x| x.coeffs.iter().map(|z| z.v)).collect();
        let mut c = vec![0u64; fa.len()];
        sys::check(unsafe { sys::fhe_rq_mul(plan(&p), fa.as_ptr(), fb.as_ptr(), c.as_mut_ptr(), a.len(), 0, std::ptr::null_mut()) });
        c.chunks(p.n).map(|v| Rq { param: p, coeffs: zqs(p.q, v.to_vec()), evals: None }).collect()
    }
    /// same, over the packed 32-bit wire (q <= 2^32: half the PCIe bytes; 5.7 M vs 3.0 M polymul/s end to end)
    pub fn mul_batch_u32(a: &[Rq], b: &[Rq]) -> Vec<Rq> {
        let p = a[0].param;
        assert!(p.q <= 1u64 << 32);
        let fa: Vec<u32> = a.iter().flat_map(|x| x.coeffs.iter().map(|z| z.v as u32)).collect();
        let fb: Vec<u32> = b.iter().flat_map(|x| x.coeffs.iter().map(|z| z.v as u32)).collect();
        let mut c = vec![0u32; fa.len()];
        sys::check(unsafe { sys::fhe_rq_mul_u32(plan(&p), fa.as_ptr(), fb.as_ptr(), c.as_mut_ptr(), a.len(), 0, std::ptr::null_mut()) });
        c.chunks(p.n).map(|v| Rq { param: p, coeffs: v.iter().map(|&x| Zq { q: p.q, v: x as u64 }).collect(), evals: None }).collect()
    }
    /// replaces ring_nq::mul_mut (arith/src/ring_nq.rs:564-583): like `mul`, and additionally stores the operands'
    /// transforms back into them (so that later products skip those transforms, exactly as on the CPU)
    pub fn mul_mut(lhs: &mut Rq, rhs: &mut Rq) -> Rq {
        assert_eq!(lhs.param, rhs.param);
        let (p, n) = (lhs.param, lhs.param.n);
        if lhs.evals.is_none() { lhs.evals = Some(ntt(lhs).coeffs); }   // NTT::ntt(lhs).coeffs, ring_nq.rs:569
        if rhs.evals.is_none() { rhs.evals = Some(ntt(rhs).coeffs); }   // ring_nq.rs:572
        let (a, b) = (vals(lhs.evals.as_ref().unwrap()), vals(rhs.evals.as_ref().unwrap()));
        let (mut c, mut ce) = (vec![0u64; n], vec![0u64; n]);
        sys::check(unsafe {
            sys::fhe_rq_mul(plan(&p), a.as_ptr(), b.as_ptr(), c.as_mut_ptr(), 1, sys::FHE_A_IS_EVALS | sys::FHE_B_IS_EVALS, ce.as_mut_ptr())
        });
        Rq { param: p, coeffs: zqs(p.q, c), evals: Some(zqs(p.q, ce)) }
    }
    /// batched products over the bit-packed wire (ceil(log2 q) bits per coefficient; n >= 1024, q < 2^30): the host path
    /// is PCIe-bound, so bytes on the wire are throughput (10.5 M polymul/s end to end against 5.9 M for u32 words)
    pub fn mul_batch_packed(a: &[Rq], b: &[Rq]) -> Vec<Rq> {
        let p = a[0].param;
        let bits = (64 - (p.q - 1).leading_zeros()).max(16) as i32;
        let words = p.n / 32 * bits as usize;
        let gather = |xs: &[Rq]| -> Vec<u32> {
            let flat: Vec<u64> = xs.iter().flat_map(|x| x.coeffs.iter().map(|z| z.v)).collect();
            let mut w = vec![0u32; xs.len() * words];
            sys::check(unsafe { sys::fhe_pack_bits(bits, flat.as_ptr(), w.as_mut_ptr(), flat.len()) });
            w
        };
        let (fa, fb) = (gather(a), gather(b));
        let mut c = vec![0u32; fa.len()];
        sys::check(unsafe { sys::fhe_rq_mul_packed(plan(&p), bits, fa.as_ptr(), fb.as_ptr(), c.as_mut_ptr(), a.len(), 0, std::ptr::null_mut()) });
        let mut flat = vec![0u64; a.len() * p.n];
        sys::check(unsafe { sys::fhe_unpack_bits(bits, c.as_ptr(), flat.as_mut_ptr(), flat.len()) });
        flat.chunks(p.n).map(|v| Rq { param: p, coeffs: zqs(p.q, v.to_vec()), evals: None }).collect()
    }
    /// replaces ring_torus::naive_poly_mul (arith/src/ring_torus.rs:266-298).  The pointer casts below need
    ///     #[repr(transparent)] pub struct T64(pub u64);          // arith/src/torus.rs:12-13, the one-line patch
    /// (a tuple struct of one u64 has that layout in practice, the attribute makes it a guarantee).
    pub fn tn_mul(a: &Tn, b: &Tn) -> Tn {
        let n = a.param.n;
        let mut c = vec![T64(0); n];
        sys::check(unsafe { sys::fhe_tn_mul(n as u64, a.coeffs.as_ptr() as *const u64, b.coeffs.as_ptr() as *const u64, c.as_mut_ptr() as *mut u64, 1) });
        Tn { param: a.param, coeffs: c }
    }
    /// flat u64 words of a slice of torus polynomials (used by the tfhe patch to build the layouts of SURVEY 8b)
    pub fn tn_words(polys: &[Tn]) -> Vec<u64> { polys.iter().flat_map(|p| p.coeffs.iter().map(|c| c.0)).collect() }
    pub fn tn_from_words(param: RingParam, w: &[u64]) -> Tn { Tn { param, coeffs: w.iter().map(|&x| T64(x)).collect() } }
}

// call sites (each guarded so that the CPU body stays the default):
//   arith/src/ntt.rs:44         pub fn ntt(a: &Rq) -> Rq  { #[cfg(feature = "gpu")] return gpu::ntt(a);  /* CPU body */ }
//   arith/src/ntt.rs:78         pub fn intt(a: &Rq) -> Rq { #[cfg(feature = "gpu")] return gpu::intt(a); /* CPU body */ }
//   arith/src/ring_nq.rs:564    fn mul_mut(lhs: &mut Rq, rhs: &mut Rq) -> Rq { #[cfg(feature = "gpu")] return gpu::mul_mut(lhs, rhs); ... }
//   arith/src/ring_nq.rs:586    fn mul(lhs: &Rq, rhs: &Rq) -> Rq { #[cfg(feature = "gpu")] return gpu::mul(lhs, rhs); ... }
//   arith/src/ring_torus.rs:266 fn naive_poly_mul(..)     { #[cfg(feature = "gpu")] return gpu::tn_mul(..); ... }
//   arith/src/torus.rs:12       #[repr(transparent)] on T64
// The scheme crates have their own patch files beside this one: tfhe_gpu_feature.rs (TGGSW x TGLWE, cmux, key_switch,
// bootstrapping), bfv_gpu_feature.rs (RLWE::mul, mul_const, keys), gfhe_gpu_feature.rs (GLWE<Rq>::key_switch, GLev * Vec<Rq>).
