// bfv_gpu_feature.rs -- the `#[cfg(feature = "gpu")]` bodies for crate `bfv` (feature gpu = ["arith/gpu", "dep:fhe-b200-sys"]).
// RLWE::tensor / RLWE::mul / BFV::relinearize_204 / mul_const are private in the reference (bfv/src/lib.rs:59-90,189-271):
// their bodies are swapped, and one PUBLIC batched entry point is added beside them ("a public batched BFV-mul
// entry point has to be added, not preserved").  Not compiled in this repository; mirrored by include/fhe_b200.hpp.

#[cfg(feature = "gpu")]
pub(crate) mod gpu {
    use crate::{Param, PublicKey, SecretKey, RLK, RLWE};
    use arith::{Rq, RingParam, Zq};
    use fhe_b200_sys as sys;

    fn vals(p: &Rq) -> impl Iterator<Item = u64> + '_ { p.coeffs().iter().map(|z| z.v) }
    fn rq(param: &RingParam, w: &[u64]) -> Rq { Rq::from_vec_u64(param, w.to_vec()) }   // values are canonical: no reduction happens
    fn rlwe_words(c: &RLWE) -> Vec<u64> { vals(&c.0).chain(vals(&c.1)).collect() }
    fn rlwe_from(param: &RingParam, w: &[u64]) -> RLWE { RLWE(rq(param, &w[..param.n]), rq(param, &w[param.n..])) }
    fn plan(p: &RingParam) -> *mut sys::FheNttPlan { arith::gpu_plan(p) }               // re-export of arith's (q, n) -> plan cache

    /// replaces RLWE::tensor (bfv/src/lib.rs:59-85)
    pub fn tensor(t: u64, a: &RLWE, b: &RLWE) -> (Rq, Rq, Rq) {
        let p = a.0.param;
        let (wa, wb, mut out) = (rlwe_words(a), rlwe_words(b), vec![0u64; 3 * p.n]);
        sys::check(unsafe { sys::fhe_bfv_tensor(p.q, p.n as u64, t, wa.as_ptr(), wb.as_ptr(), out.as_mut_ptr(), 1) });
        (rq(&p, &out[..p.n]), rq(&p, &out[p.n..2 * p.n]), rq(&p, &out[2 * p.n..]))
    }
    /// replaces BFV::relinearize_204 (bfv/src/lib.rs:251-271)
    pub fn relinearize_204(rlk: &RLK, c0: &Rq, c1: &Rq, c2: &Rq) -> RLWE {
        let p = c0.param;
        let k: Vec<u64> = vals(&rlk.0).chain(vals(&rlk.1)).collect();
        let c: Vec<u64> = vals(c0).chain(vals(c1)).chain(vals(c2)).collect();
        let mut out = vec![0u64; 2 * p.n];
        sys::check(unsafe { sys::fhe_bfv_relinearize(p.q, p.n as u64, rlk.0.param.q, k.as_ptr(), c.as_ptr(), out.as_mut_ptr(), 1) });
        rlwe_from(&p, &out)
    }
    /// replaces RLWE::mul (bfv/src/lib.rs:87-90): tensor + relinearisation in one kernel
    pub fn mul(t: u64, rlk: &RLK, a: &RLWE, b: &RLWE) -> RLWE { mul_batch(t, rlk, std::slice::from_ref(a), std::slice::from_ref(b)).pop().unwrap() }
    /// the added public batched entry point (config 3: batch of 4096)
    pub fn mul_batch(t: u64, rlk: &RLK, a: &[RLWE], b: &[RLWE]) -> Vec<RLWE> {
        let p = a[0].0.param;
        let k: Vec<u64> = vals(&rlk.0).chain(vals(&rlk.1)).collect();
        let wa: Vec<u64> = a.iter().flat_map(rlwe_words).collect();
        let wb: Vec<u64> = b.iter().flat_map(rlwe_words).collect();
        let mut out = vec![0u64; wa.len()];
        sys::check(unsafe { sys::fhe_bfv_mul_relin(p.q, p.n as u64, t, rlk.0.param.q, k.as_ptr(), wa.as_ptr(), wb.as_ptr(), out.as_mut_ptr(), a.len()) });
        out.chunks(2 * p.n).map(|w| rlwe_from(&p, w)).collect()
    }
    /// replaces BFV::mul_const (bfv/src/lib.rs:189-200)
    pub fn mul_const(rlk: &RLK, c: &RLWE, m: &Rq) -> RLWE {
        let (p, t) = (c.0.param, m.param.q);
        let k: Vec<u64> = vals(&rlk.0).chain(vals(&rlk.1)).collect();
        let (wc, wm, mut out) = (rlwe_words(c), vals(m).collect::<Vec<u64>>(), vec![0u64; 2 * p.n]);
        sys::check(unsafe { sys::fhe_bfv_mul_const(p.q, p.n as u64, t, rlk.0.param.q, k.as_ptr(), wc.as_ptr(), wm.as_ptr(), out.as_mut_ptr(), 1) });
        rlwe_from(&p, &out)
    }
    /// device-sampled twins of BFV::new_key / rlk_key / encrypt (bfv/src/lib.rs:120-160,202-225).  The reference draws from
    /// the caller's `impl Rng`; a device sampler cannot consume that stream, so these take a seed and are ADDED beside the
    /// rng-taking functions (whose Rq products already run on the GPU through arith's patch).
    pub fn new_key_seeded(seed: u64, param: &Param) -> (SecretKey, PublicKey) {
        let (p, n) = (param.ring, param.ring.n);
        let (mut sk, mut pk) = (vec![0u64; n], vec![0u64; 2 * n]);
        sys::check(unsafe { sys::fhe_bfv_keygen(plan(&p), p.q, n as u64, crate::ERR_SIGMA, seed, sk.as_mut_ptr(), pk.as_mut_ptr()) });
        let mut s = rq(&p, &sk);
        s.compute_evals();                                                             // lib.rs:132
        (SecretKey(s), PublicKey(rq(&p, &pk[..n]), rq(&p, &pk[n..])))
    }
    pub fn rlk_key_seeded(seed: u64, param: &Param, s: &SecretKey) -> RLK {
        let (p, n) = (param.ring, param.ring.n);
        let pq = RingParam { q: param.p * p.q, n };
        let (sk, mut out) = (vals(&s.0).collect::<Vec<u64>>(), vec![0u64; 2 * n]);
        sys::check(unsafe { sys::fhe_bfv_rlk_generate(p.q, n as u64, param.p, crate::ERR_SIGMA, seed, sk.as_ptr(), out.as_mut_ptr()) });
        RLK(rq(&pq, &out[..n]), rq(&pq, &out[n..]))
    }
    pub fn encrypt_batch_seeded(seed: u64, param: &Param, pk: &PublicKey, ms: &[Rq]) -> Vec<RLWE> {
        let (p, n) = (param.ring, param.ring.n);
        let k: Vec<u64> = vals(&pk.0).chain(vals(&pk.1)).collect();
        let m: Vec<u64> = ms.iter().flat_map(|x| vals(x)).collect();
        let mut out = vec![0u64; ms.len() * 2 * n];
        sys::check(unsafe { sys::fhe_bfv_encrypt(plan(&p), p.q, n as u64, param.t, k.as_ptr(), m.as_ptr(), crate::ERR_SIGMA, seed, out.as_mut_ptr(), ms.len()) });
        out.chunks(2 * n).map(|w| rlwe_from(&p, w)).collect()
    }
    pub fn decrypt_batch(param: &Param, sk: &SecretKey, cs: &[RLWE]) -> Vec<Rq> {
        let (p, n) = (param.ring, param.ring.n);
        let (s, c) = (vals(&sk.0).collect::<Vec<u64>>(), cs.iter().flat_map(rlwe_words).collect::<Vec<u64>>());
        let mut m = vec![0u64; cs.len() * n];
        sys::check(unsafe { sys::fhe_bfv_decrypt(plan(&p), p.q, n as u64, param.t, s.as_ptr(), c.as_ptr(), m.as_mut_ptr(), cs.len()) });
        let pt = RingParam { q: param.t, n };
        m.chunks(n).map(|w| rq(&pt, w)).collect()
    }
    #[allow(dead_code)] fn _zq(_: Zq) {}
}

// call sites:
//   bfv/src/lib.rs:59    fn tensor(t, a, b) -> (Rq, Rq, Rq)       { #[cfg(feature = "gpu")] return gpu::tensor(t, a, b); ... }
//   bfv/src/lib.rs:87    fn mul(t, rlk, a, b) -> Self             { #[cfg(feature = "gpu")] return gpu::mul(t, rlk, a, b); ... }
//   bfv/src/lib.rs:189   fn mul_const(rlk, c, m) -> RLWE          { #[cfg(feature = "gpu")] return gpu::mul_const(rlk, c, m); ... }
//   bfv/src/lib.rs:251   fn relinearize_204(rlk, c0, c1, c2)      { #[cfg(feature = "gpu")] return gpu::relinearize_204(rlk, c0, c1, c2); ... }
//   bfv/src/lib.rs:142,164 encrypt / decrypt: unchanged source -- their `&pk.0 * &u`, `&c.1 * &sk.0` products run on the GPU through arith's patch
//   arith/src/lib.rs     #[cfg(feature = "gpu")] pub fn gpu_plan(p: &RingParam) -> *mut fhe_b200_sys::FheNttPlan   // re-export of gpu::plan
//   new, beside the private API: pub fn mul_batch / new_key_seeded / rlk_key_seeded / encrypt_batch_seeded / decrypt_batch
