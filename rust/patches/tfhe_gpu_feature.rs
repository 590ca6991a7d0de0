// tfhe_gpu_feature.rs -- the `#[cfg(feature = "gpu")]` bodies for crate `tfhe` (feature gpu = ["arith/gpu", "dep:fhe-b200-sys"]).
// The public API is unchanged: `TGGSW * TGLWE`, `TGGSW::cmux`, `TLWE::key_switch`, `blind_rotation`, `bootstrapping` keep their
// signatures; only their bodies call libfhe_b200.  Resident device copies of the keys are cached per key object (the
// reference deep-clones BSK and KSK twice per bootstrapping call, tlwe.rs:107,157; here a clone of a key shares its handle).
// Not compiled in this repository (no Rust toolchain in the build image); include/fhe_b200.hpp mirrors every function
// below call for call and tests/cpp/test_host_api.cpp runs them.
//
// Layouts: TGLWE = k mask polynomials then the body, (k+1)*n words; TGLev = l TGLWEs (level 0 = MSB digit);
// TGGSW = the k mask TGLevs then the body TGLev; TLWE = kn mask words then b; KSK = kn_in * l rows of kn_out + 1 words.

#[cfg(feature = "gpu")]
pub(crate) mod gpu {
    use crate::{tggsw::{TGGSW, TGLev}, tglwe::TGLWE, tlev::TLev, tlwe::{BootstrappingKey, KSK, TLWE}};
    use arith::{Ring, RingParam, Tn, T64, TR};
    use fhe_b200_sys as sys;
    use gfhe::glwe::{Param, GLWE};
    use std::{collections::HashMap, sync::{Mutex, OnceLock}};

    // ---- flat words <-> the reference's nested types -------------------------------------------------------------
    fn tglwe_words(c: &TGLWE) -> Vec<u64> {
        let mut w: Vec<u64> = c.0 .0.r.iter().flat_map(|p| p.coeffs().iter().map(|x| x.0)).collect();   // mask a_0..a_{k-1}
        w.extend(c.0 .1.coeffs().iter().map(|x| x.0));                                                    // body
        w
    }
    fn tglwe_from_words(k: usize, ring: &RingParam, w: &[u64]) -> TGLWE {
        let n = ring.n;
        let poly = |i: usize| Tn::from_vec(ring, w[i * n..(i + 1) * n].iter().map(|&x| T64(x)).collect());
        TGLWE(GLWE(TR { k, r: (0..k).map(poly).collect() }, poly(k)))
    }
    fn tlwe_words(c: &TLWE) -> Vec<u64> {
        let mut w: Vec<u64> = c.0 .0.r.iter().map(|x| x.0).collect();
        w.push(c.0 .1 .0);
        w
    }
    fn tlwe_from_words(w: &[u64]) -> TLWE {
        let kn = w.len() - 1;
        TLWE(GLWE(TR { k: kn, r: w[..kn].iter().map(|&x| T64(x)).collect() }, T64(w[kn])))
    }
    fn tlev_rows(lev: &TGLev, out: &mut Vec<u64>) { for row in &lev.0 { out.extend(tglwe_words(row)); } }

    // ---- resident keys: one device copy per key object, found again through the address of its first row ----------
    // (a maintainer who may add a field stores the handle in the struct instead: `TGGSW(.., OnceLock<Handle>)`)
    fn cache() -> &'static Mutex<HashMap<(usize, usize), usize>> {
        static C: OnceLock<Mutex<HashMap<(usize, usize), usize>>> = OnceLock::new();
        C.get_or_init(Default::default)
    }
    fn tggsw_handle(g: &TGGSW) -> *const sys::FheTggsw {
        let first = &g.0[0].0[0];
        let (n, k) = (first.0 .1.param().n, first.0 .0.k);
        let key = (first as *const TGLWE as usize, 0usize);
        let mut m = cache().lock().unwrap();
        *m.entry(key).or_insert_with(|| {
            let mut rows = Vec::with_capacity((k + 1) * 64 * (k + 1) * n);
            for lev in &g.0 { tlev_rows(lev, &mut rows); }                 // rows of the k mask polynomials first
            tlev_rows(&g.1, &mut rows);                                     // the body row last (tggsw.rs:52-54)
            let mut h = std::ptr::null_mut();
            sys::check(unsafe { sys::fhe_tggsw_load(n as u64, k as u64, rows.as_ptr(), &mut h) });
            h as usize
        }) as *const sys::FheTggsw
    }
    fn ksk_handle(ksk: &KSK, kn_out: usize, l: usize) -> *const sys::FheKsk {
        let levs: &Vec<TLev> = ksk.levs();                                   // accessor added by the patch: `&self.0`
        let key = (levs.as_ptr() as usize, 1usize);
        let mut m = cache().lock().unwrap();
        *m.entry(key).or_insert_with(|| {
            let mut rows = Vec::with_capacity(levs.len() * l * (kn_out + 1));
            for lev in levs { for row in &lev.0 { rows.extend(tlwe_words(row)); } }   // row i*l + j = TLWE_{i,j} (tlwe.rs:84-100)
            let mut h = std::ptr::null_mut();
            sys::check(unsafe { sys::fhe_ksk_load(levs.len() as u64, kn_out as u64, l as u64, rows.as_ptr(), &mut h) });
            h as usize
        }) as *const sys::FheKsk
    }

    /// replaces `impl Mul<TGLWE> for TGGSW` (tfhe/src/tggsw.rs:45-62): decompose(2, 64) of every component, the
    /// (k+1)*64 TGLWE * Tn products and their sum are ONE fused kernel (fhe_extprod)
    pub fn external_product(g: &TGGSW, ct: &TGLWE) -> TGLWE {
        let (k, ring) = (ct.0 .0.k, *ct.0 .1.param());
        let (src, mut out) = (tglwe_words(ct), vec![0u64; (k + 1) * ring.n]);
        sys::check(unsafe { sys::fhe_extprod(tggsw_handle(g), src.as_ptr(), out.as_mut_ptr(), 1) });
        tglwe_from_words(k, &ring, &out)
    }
    /// replaces TGGSW::cmux (tfhe/src/tggsw.rs:39-41): ct1 + bit * (ct2 - ct1), subtraction and addition fused in
    pub fn cmux(bit: &TGGSW, ct1: &TGLWE, ct2: &TGLWE) -> TGLWE {
        let (k, ring) = (ct1.0 .0.k, *ct1.0 .1.param());
        let (a, b, mut out) = (tglwe_words(ct1), tglwe_words(ct2), vec![0u64; (k + 1) * ring.n]);
        sys::check(unsafe { sys::fhe_cmux(tggsw_handle(bit), a.as_ptr(), b.as_ptr(), out.as_mut_ptr(), 1) });
        tglwe_from_words(k, &ring, &out)
    }
    /// batched twin beside the operator API: `cts.len()` accumulators against one TGGSW in one launch (config 4)
    pub fn cmux_batch(bit: &TGGSW, ct1: &[TGLWE], ct2: &[TGLWE]) -> Vec<TGLWE> {
        let (k, ring) = (ct1[0].0 .0.k, *ct1[0].0 .1.param());
        let glwe = (k + 1) * ring.n;
        let a: Vec<u64> = ct1.iter().flat_map(tglwe_words).collect();
        let b: Vec<u64> = ct2.iter().flat_map(tglwe_words).collect();
        let mut out = vec![0u64; a.len()];
        sys::check(unsafe { sys::fhe_cmux(tggsw_handle(bit), a.as_ptr(), b.as_ptr(), out.as_mut_ptr(), ct1.len()) });
        out.chunks(glwe).map(|w| tglwe_from_words(k, &ring, w)).collect()
    }
    /// replaces TLWE::key_switch (tfhe/src/tlwe.rs:101-112); the library's engines need beta = 2 (the only value the
    /// reference passes, tlwe.rs:159)
    pub fn key_switch(c: &TLWE, param: &Param, beta: u32, l: u32, ksk: &KSK) -> TLWE {
        assert_eq!(beta, 2, "fhe_b200: key_switch is built for beta = 2 (tfhe/src/tlwe.rs:159)");
        let kn_out = param.k * param.ring.n;
        let (src, mut out) = (tlwe_words(c), vec![0u64; kn_out + 1]);
        sys::check(unsafe { sys::fhe_key_switch(ksk_handle(ksk, kn_out, l as usize), src.as_ptr(), out.as_mut_ptr(), 1) });
        tlwe_from_words(&out)
    }
    /// replaces blind_rotation (tfhe/src/tlwe.rs:121-148) AS EXECUTED: the CMux closure is a lazy iterator that is never
    /// consumed, so the result is table.left_rotate(mod_switch(c).b)
    pub fn blind_rotation(param: &Param, c: &TLWE, _btk: &BootstrappingKey, table: &TGLWE) -> TGLWE {
        let (n, k) = (param.ring.n, param.k);
        let (t, src, mut out) = (tglwe_words(table), tlwe_words(c), vec![0u64; (k + 1) * n]);
        sys::check(unsafe {
            sys::fhe_blind_rotate(n as u64, k as u64, std::ptr::null(), 0, t.as_ptr(), src.as_ptr(), (src.len() - 1) as u64, out.as_mut_ptr(), 1)
        });
        tglwe_from_words(k, &param.ring, &out)
    }
    /// replaces bootstrapping (tfhe/src/tlwe.rs:150-161): mod_switch, rotation, sample extraction and key switch in one call
    pub fn bootstrapping(param: &Param, btk: &BootstrappingKey, table: &TGLWE, c: &TLWE) -> TLWE {
        bootstrapping_batch(param, btk, table, std::slice::from_ref(c)).pop().unwrap()
    }
    /// batched twin (config 5): all ciphertexts in one call, keys resident
    pub fn bootstrapping_batch(param: &Param, btk: &BootstrappingKey, table: &TGLWE, cs: &[TLWE]) -> Vec<TLWE> {
        let (n, k) = (param.ring.n, param.k);
        let kn = k * n;
        let c_kn = cs[0].0 .0.r.len();
        let t = tglwe_words(table);
        let src: Vec<u64> = cs.iter().flat_map(tlwe_words).collect();
        let mut out = vec![0u64; cs.len() * (kn + 1)];
        sys::check(unsafe {
            sys::fhe_bootstrap(n as u64, k as u64, ksk_handle(&btk.1, kn, 64), t.as_ptr(), src.as_ptr(), c_kn as u64, out.as_mut_ptr(), cs.len())
        });
        out.chunks(kn + 1).map(tlwe_from_words).collect()
    }
    /// replaces compute_lookup_table (tfhe/src/tlwe.rs:196-214)
    pub fn compute_lookup_table(param: &Param) -> TGLWE {
        let (n, k) = (param.ring.n, param.k);
        let mut w = vec![0u64; (k + 1) * n];
        sys::check(unsafe { sys::fhe_compute_lookup_table(n as u64, k as u64, param.t, w.as_mut_ptr()) });
        tglwe_from_words(k, &param.ring, &w)
    }
}

// call sites:
//   tfhe/src/tggsw.rs:48   fn mul(self, tglwe: TGLWE) -> TGLWE { #[cfg(feature = "gpu")] return gpu::external_product(&self, &tglwe); ... }
//   tfhe/src/tggsw.rs:39   pub fn cmux(bit: Self, ct1: TGLWE, ct2: TGLWE) -> TGLWE { #[cfg(feature = "gpu")] return gpu::cmux(&bit, &ct1, &ct2); ... }
//   tfhe/src/tlwe.rs:101   pub fn key_switch(&self, param, beta, l, ksk) -> Self { #[cfg(feature = "gpu")] return gpu::key_switch(self, param, beta, l, ksk); ... }
//   tfhe/src/tlwe.rs:121   pub fn blind_rotation(..) -> TGLWE { #[cfg(feature = "gpu")] return gpu::blind_rotation(param, &c, &btk, &table); ... }
//   tfhe/src/tlwe.rs:150   pub fn bootstrapping(..) -> TLWE   { #[cfg(feature = "gpu")] return gpu::bootstrapping(param, &btk, &table, &c); ... }
//   tfhe/src/tlwe.rs:196   pub fn compute_lookup_table(param) -> TGLWE { #[cfg(feature = "gpu")] return gpu::compute_lookup_table(param); ... }
//   tfhe/src/tlwe.rs:37    impl KSK { pub(crate) fn levs(&self) -> &Vec<TLev> { &self.0 } }    // the accessor the cache uses
//   new, beside the unchanged API: gpu::cmux_batch, gpu::bootstrapping_batch (slices of ciphertexts, one call)
