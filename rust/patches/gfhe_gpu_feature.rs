// gfhe_gpu_feature.rs -- the `#[cfg(feature = "gpu")]` bodies for crate `gfhe` (feature gpu = ["arith/gpu", "dep:fhe-b200-sys"]).
// `GLWE<R>` / `GLev<R>` are generic over `R: Ring`; the GPU bodies exist for R = Rq (the gadget products there are NTT
// polymuls).  Rust has no specialisation on stable, so the generic functions dispatch on `R::gpu_kind()` -- a defaulted
// associated function the patch adds to the `Ring` trait (arith/src/ring.rs:16-55), `None` by default and
// `Some(GpuRing::Rq)` in `impl Ring for Rq` -- and fall through to the CPU body for every other ring (GLWE<Tn> is
// accelerated transitively: its `Tn * Tn` is arith's patched naive_poly_mul).  Not compiled in this repository;
// mirrored by include/fhe_b200.hpp.

#[cfg(feature = "gpu")]
pub(crate) mod gpu {
    use crate::{glev::GLev, glwe::{Param, GLWE, KSK}};
    use arith::{Ring, RingParam, Rq, TR};
    use fhe_b200_sys as sys;
    use std::{collections::HashMap, sync::{Mutex, OnceLock}};

    fn vals(p: &Rq) -> impl Iterator<Item = u64> + '_ { p.coeffs().iter().map(|z| z.v) }
    fn glwe_words(c: &GLWE<Rq>) -> Vec<u64> { c.0.r.iter().flat_map(vals).chain(vals(&c.1)).collect() }
    fn glwe_from(k: usize, ring: &RingParam, w: &[u64]) -> GLWE<Rq> {
        let n = ring.n;
        let poly = |i: usize| Rq::from_vec_u64(ring, w[i * n..(i + 1) * n].to_vec());
        GLWE(TR { k, r: (0..k).map(poly).collect() }, poly(k))
    }
    fn cache() -> &'static Mutex<HashMap<usize, usize>> {
        static C: OnceLock<Mutex<HashMap<usize, usize>>> = OnceLock::new();
        C.get_or_init(Default::default)
    }
    /// rows (GLWEs) of a gadget key, transformed once and kept resident (fhe_rq_glev_load)
    fn glev_handle(first_row: *const GLWE<Rq>, ring: &RingParam, k: usize, rows: &mut dyn Iterator<Item = &GLWE<Rq>>) -> *const sys::FheRqGlev {
        let mut m = cache().lock().unwrap();
        *m.entry(first_row as usize).or_insert_with(|| {
            let w: Vec<u64> = rows.flat_map(glwe_words).collect();
            let nrows = w.len() / ((k + 1) * ring.n);
            let mut h = std::ptr::null_mut();
            sys::check(unsafe { sys::fhe_rq_glev_load(arith::gpu_plan(ring), k as u64, nrows as u64, w.as_ptr(), &mut h) });
            h as usize
        }) as *const sys::FheRqGlev
    }

    /// replaces `impl Mul<Vec<R>> for GLev<R>` (gfhe/src/glev.rs:67-80) for R = Rq: sum_j GLWE_j * v_j
    pub fn glev_mul(lev: &GLev<Rq>, v: &[Rq]) -> GLWE<Rq> {
        let (k, ring) = (lev.0[0].0.k, lev.0[0].1.param);
        let h = glev_handle(lev.0.as_ptr(), &ring, k, &mut lev.0.iter());
        let (src, mut out) = (v.iter().flat_map(vals).collect::<Vec<u64>>(), vec![0u64; (k + 1) * ring.n]);
        sys::check(unsafe { sys::fhe_rq_glev_mul(h, src.as_ptr(), out.as_mut_ptr(), 1) });
        glwe_from(k, &ring, &out)
    }
    /// replaces GLWE<R>::key_switch (gfhe/src/glwe.rs:126-137) for R = Rq: (0, b) - sum_i KSK_i * decompose(a_i)
    pub fn key_switch(c: &GLWE<Rq>, param: &Param, beta: u32, l: u32, ksk: &KSK<Rq>) -> GLWE<Rq> {
        key_switch_batch(std::slice::from_ref(c), param, beta, l, ksk).pop().unwrap()
    }
    pub fn key_switch_batch(cs: &[GLWE<Rq>], param: &Param, beta: u32, l: u32, ksk: &KSK<Rq>) -> Vec<GLWE<Rq>> {
        let (k, ring) = (param.k, param.ring);
        let levs = ksk.levs();                                                     // accessor added by the patch: `&self.0`
        let h = glev_handle(levs[0].0.as_ptr(), &ring, k, &mut levs.iter().flat_map(|g| g.0.iter()));   // row i*l + j = GLWE_{i,j}
        let src: Vec<u64> = cs.iter().flat_map(glwe_words).collect();
        let mut out = vec![0u64; src.len()];
        sys::check(unsafe { sys::fhe_glwe_rq_key_switch(h, beta, l, src.as_ptr(), out.as_mut_ptr(), cs.len()) });
        out.chunks((k + 1) * ring.n).map(|w| glwe_from(k, &ring, w)).collect()
    }
    /// replaces GLWE<R> * R (gfhe/src/glwe.rs:263-280) for R = Rq: every component times one polynomial (FHE_B_BROADCAST)
    pub fn mul_poly(c: &GLWE<Rq>, p: &Rq) -> GLWE<Rq> {
        let (k, ring) = (c.0.k, c.1.param);
        let (src, b, mut out) = (glwe_words(c), vals(p).collect::<Vec<u64>>(), vec![0u64; (k + 1) * ring.n]);
        sys::check(unsafe {
            sys::fhe_rq_mul(arith::gpu_plan(&ring), src.as_ptr(), b.as_ptr(), out.as_mut_ptr(), k + 1, sys::FHE_B_BROADCAST, std::ptr::null_mut())
        });
        glwe_from(k, &ring, &out)
    }
    #[allow(dead_code)] fn _ring<R: Ring>() {}
}

// call sites (R generic; `as_rq` is the checked downcast the patch adds next to `gpu_kind`):
//   gfhe/src/glev.rs:69   fn mul(self, v: Vec<R>) -> GLWE<R> { #[cfg(feature = "gpu")] if let Some((lev, v)) = as_rq(&self, &v) { return from_rq(gpu::glev_mul(lev, v)); } ... }
//   gfhe/src/glwe.rs:126  pub fn key_switch(&self, param, beta, l, ksk) -> Self { #[cfg(feature = "gpu")] if R::gpu_kind() == Some(GpuRing::Rq) { return ...gpu::key_switch(..) } ... }
//   gfhe/src/glwe.rs:263  impl Mul<R> for GLWE<R>                  { #[cfg(feature = "gpu")] ... gpu::mul_poly(..) ... }
//   gfhe/src/glwe.rs:66   impl<R: Ring> KSK<R> { pub(crate) fn levs(&self) -> &Vec<GLev<R>> { &self.0 } }
//   arith/src/ring.rs:16  trait Ring { ...; #[cfg(feature = "gpu")] fn gpu_kind() -> Option<GpuRing> { None } }
