//! Drop-in replacement for the public API of `spread_spectrum_watermarking`
//! (reference: src/lib.rs:75-85, src/algorithm.rs) whose arithmetic runs in the sm_100a CUDA
//! library `libssw` through the C ABI of include/ssw.h.
//!
//! Same type and method names, same signatures, same panics.  Deviations (all documented in DESIGN.md):
//! `Insertion::Custom`, `Extraction::Custom`, `OrderingMethod::Custom` panic (host closures cannot run on the device
//! and there is no CPU fallback); `Reader::indices()` computes the full ordering on first use instead of in `base`;
//! `dct2d::dct2_2d` is `f32` only and its `DctPlanner` is a unit stand-in for `rustdct::DctPlanner<f32>` (the plans
//! live in the library's per-thread context).
//!
//! UNVERIFIED SKETCH: this repository's build image has no cargo / rustc, so the crate has never been compiled.
//! The C ABI underneath IS exercised from C (tests/cabi/cabi_flow.c) and Python (ctypes).
pub mod ffi;

use ffi::*;
use image::{DynamicImage, ImageBuffer, Luma, Rgb32FImage};
use std::cell::OnceCell;
use std::ptr;
use std::rc::Rc;

pub mod prelude {
    pub use crate::Mark;
}

pub type Luma32FImage = ImageBuffer<Luma<f32>, Vec<f32>>;

// ---- context: one per thread, kept alive by every object created from it (Writer/Reader are !Send + !Sync in the
//      reference as well: they own a DctPlanner and boxed closures, src/algorithm.rs:286-291,441-445) -----------------
struct Ctx(*mut ssw_ctx);
impl Drop for Ctx {
    fn drop(&mut self) { unsafe { ssw_ctx_destroy(self.0); } }
}
thread_local! {
    static CTX: OnceCell<Rc<Ctx>> = OnceCell::new();
}
fn ctx() -> Rc<Ctx> {
    CTX.with(|c| {
        c.get_or_init(|| {
            let mut h = ptr::null_mut();
            let dev = std::env::var("SSW_DEVICE").ok().and_then(|s| s.parse().ok()).unwrap_or(0);
            check(unsafe { ssw_ctx_create(dev, &mut h) });
            Rc::new(Ctx(h))
        })
        .clone()
    })
}

fn dims(width: u32, height: u32) -> usize { width as usize * height as usize }

// ---- configuration (src/algorithm.rs:24-152) -----------------------------------------------------------------------
pub type InsertFunction = Box<dyn Fn(usize, f32, f32) -> f32>;
pub type ExtractFunction = Box<dyn Fn(usize, f32, f32) -> f32>;
pub type OrderingFunction = Box<dyn Fn(usize, f32, usize, f32) -> std::cmp::Ordering>;

pub enum Insertion { Option1(f32), Option2(f32), Option3(f32), Custom(InsertFunction) }
pub enum Extraction { Option1(f32), Option2(f32), Option3(f32), Custom(ExtractFunction) }
pub enum OrderingMethod { Energy, EnergyOrthogonal, Legacy, Custom(OrderingFunction) }

pub struct WriteConfig { pub insertion: Insertion, pub ordering: OrderingMethod }
pub struct ReadConfig { pub extraction: Extraction, pub ordering: OrderingMethod }
impl Default for WriteConfig {
    fn default() -> Self { WriteConfig { insertion: Insertion::Option2(0.1), ordering: OrderingMethod::Energy } }
}
impl Default for ReadConfig {
    fn default() -> Self { ReadConfig { extraction: Extraction::Option2(0.1), ordering: OrderingMethod::Energy } }
}

fn ordering_code(o: &OrderingMethod) -> i32 {
    match o {
        OrderingMethod::Energy => 0,
        OrderingMethod::EnergyOrthogonal => 1,
        OrderingMethod::Legacy => 2,
        OrderingMethod::Custom(_) => panic!("OrderingMethod::Custom is a host closure; libssw has no CPU path"),
    }
}
fn write_cfg(c: &WriteConfig) -> ssw_config {
    let (method, alpha) = match c.insertion {
        Insertion::Option1(a) => (1, a),
        Insertion::Option2(a) => (2, a),
        Insertion::Option3(a) => (3, a),
        Insertion::Custom(_) => panic!("Insertion::Custom is a host closure; libssw has no CPU path"),
    };
    ssw_config { method, alpha, ordering: ordering_code(&c.ordering) }
}
fn read_cfg(c: &ReadConfig) -> ssw_config {
    let (method, alpha) = match c.extraction {
        Extraction::Option1(a) => (1, a),
        Extraction::Option2(a) => (2, a),
        Extraction::Option3(a) => (3, a),
        Extraction::Custom(_) => panic!("Extraction::Custom is a host closure; libssw has no CPU path"),
    };
    ssw_config { method, alpha, ordering: ordering_code(&c.ordering) }
}

// ---- marks (src/algorithm.rs:596-666) --------------------------------------------------------------------------------
pub trait Mark { fn data(&self) -> &[f32]; }

#[derive(Clone, Debug, Default)]
pub struct MarkBuf { data: Vec<f32> }
impl MarkBuf {
    pub fn new() -> Self { MarkBuf { data: vec![] } }
    /// N(0,1) samples drawn on the device (Philox + Box-Muller), OS-seeded like `thread_rng` (:619-626).
    pub fn generate_normal(length: usize) -> Self {
        let mut data = vec![0f32; length];
        check(unsafe { ssw_mark_generate_normal(ctx().0, 0, length, data.as_mut_ptr()) });
        MarkBuf { data }
    }
    pub fn from(data: &[f32]) -> Self { MarkBuf { data: data.to_vec() } }
    pub fn data(&self) -> &[f32] { &self.data }
    pub fn set_data(&mut self, data: &[f32]) { self.data = data.to_vec(); }
}
impl Mark for MarkBuf { fn data(&self) -> &[f32] { &self.data } }
impl Mark for &MarkBuf { fn data(&self) -> &[f32] { &self.data } }
// the reference's blanket impl (:659-666): Vec<f32>, arrays, Box<[f32]>, slices ... (MarkBuf itself is not AsRef<[f32]>)
impl<T: AsRef<[f32]>> Mark for T {
    fn data(&self) -> &[f32] { self.as_ref() }
}

// ---- Writer (src/algorithm.rs:286-433) -------------------------------------------------------------------------------
pub struct Writer { h: *mut ssw_writer, width: u32, height: u32, coeff: OnceCell<Luma32FImage>, _ctx: Rc<Ctx> }
impl Writer {
    pub fn new(image: DynamicImage, config: WriteConfig) -> Self {
        let rgb = image.into_rgb32f();
        let (width, height) = (rgb.width(), rgb.height());
        let cfg = write_cfg(&config);
        let c = ctx();
        let mut h = ptr::null_mut();
        check(unsafe { ssw_writer_new_rgb32f(c.0, rgb.as_raw().as_ptr(), width, height, &cfg, &mut h) });
        Writer { h, width, height, coeff: OnceCell::new(), _ctx: c }
    }
    /// The coefficients of the Y channel (:319-321); downloaded on first use, again after `embed`.
    pub fn coefficient_image(&self) -> &Luma32FImage {
        self.coeff.get_or_init(|| {
            let mut v = vec![0f32; dims(self.width, self.height)];
            check(unsafe { ssw_writer_coefficients(self.h, v.as_mut_ptr()) });
            Luma32FImage::from_raw(self.width, self.height, v).unwrap()
        })
    }
    pub fn embed(&mut self, marks: &[&dyn Mark]) {
        let ptrs: Vec<*const f32> = marks.iter().map(|m| m.data().as_ptr()).collect();
        let lens: Vec<usize> = marks.iter().map(|m| m.data().len()).collect();
        check(unsafe { ssw_writer_embed(self.h, ptrs.as_ptr(), lens.as_ptr(), marks.len()) });
        self.coeff = OnceCell::new();
    }
    pub fn mark(mut self, marks: &[&dyn Mark]) -> DynamicImage {
        self.embed(marks);
        self.result()
    }
    pub fn result(self) -> DynamicImage {
        let mut v = vec![0f32; dims(self.width, self.height) * 3];
        check(unsafe { ssw_writer_result_rgb32f(self.h, v.as_mut_ptr()) });
        DynamicImage::ImageRgb32F(Rgb32FImage::from_raw(self.width, self.height, v).unwrap())
    }
}
impl Drop for Writer {
    fn drop(&mut self) { unsafe { ssw_writer_destroy(self.h); } }   // before `_ctx` (fields drop after this body)
}

// ---- Reader (src/algorithm.rs:435-594) -------------------------------------------------------------------------------
pub struct Reader { h: *mut ssw_reader, n: usize, coeff: OnceCell<Vec<f32>>, idx: OnceCell<Vec<usize>>, _ctx: Rc<Ctx> }
pub struct ReaderDerived(Reader);
impl ReaderDerived {
    pub fn new(image: DynamicImage) -> Self { Reader::derived(image) }
}
impl Reader {
    pub fn base(image: DynamicImage, config: ReadConfig) -> Self {
        let rgb = image.into_rgb32f();
        let cfg = read_cfg(&config);
        let c = ctx();
        let mut h = ptr::null_mut();
        check(unsafe { ssw_reader_base_rgb32f(c.0, rgb.as_raw().as_ptr(), rgb.width(), rgb.height(), &cfg, &mut h) });
        Reader { h, n: dims(rgb.width(), rgb.height()), coeff: OnceCell::new(), idx: OnceCell::new(), _ctx: c }
    }
    pub fn derived(image: DynamicImage) -> ReaderDerived {
        let rgb = image.into_rgb32f();
        let c = ctx();
        let mut h = ptr::null_mut();
        check(unsafe { ssw_reader_derived_rgb32f(c.0, rgb.as_raw().as_ptr(), rgb.width(), rgb.height(), &mut h) });
        ReaderDerived(Reader { h, n: dims(rgb.width(), rgb.height()), coeff: OnceCell::new(), idx: OnceCell::new(), _ctx: c })
    }
    pub fn coefficients(&self) -> &[f32] {
        self.coeff.get_or_init(|| {
            let mut v = vec![0f32; self.n];
            check(unsafe { ssw_reader_coefficients(self.h, v.as_mut_ptr()) });
            v
        })
    }
    /// All W*H-1 ordered indices, like the reference (:506-508); computed (full radix sort on the device) on first use.
    pub fn indices(&self) -> &[usize] {
        self.idx.get_or_init(|| {
            let mut v = vec![0u64; self.n - 1];
            check(unsafe { ssw_reader_indices(self.h, v.as_mut_ptr(), v.len()) });
            v.into_iter().map(|x| x as usize).collect()
        })
    }
    pub fn extract(&self, derived: &ReaderDerived, extracted: &mut [f32]) {
        check(unsafe { ssw_reader_extract(self.h, derived.0.h, extracted.as_mut_ptr(), extracted.len()) });
    }
}
impl Drop for Reader {
    fn drop(&mut self) { unsafe { ssw_reader_destroy(self.h); } }
}

// ---- Tester (src/algorithm.rs:668-715) -------------------------------------------------------------------------------
#[derive(Clone, Copy, Debug)]
pub struct Similarity { pub similarity: f32 }
impl Similarity {
    pub fn exceeds_sigma(&self, n_sigma: f32) -> bool { self.similarity > n_sigma }
}
pub struct Tester<'a> { extracted: &'a [f32] }
impl<'a> Tester<'a> {
    pub fn new(extracted_watermark: &'a [f32]) -> Self { Tester { extracted: extracted_watermark } }
    pub fn similarity(&self, comparison_watermark: &dyn Mark) -> Similarity {
        let m = comparison_watermark.data();
        assert_eq!(self.extracted.len(), m.len());
        let mut s = 0f32;
        check(unsafe { ssw_similarity(ctx().0, self.extracted.as_ptr(), m.as_ptr(), m.len(), &mut s) });
        Similarity { similarity: s }
    }
}

// ---- dct2d (src/dct2d.rs:71-219) -------------------------------------------------------------------------------------
pub mod dct2d {
    /// Stand-in for `rustdct::DctPlanner<f32>` (the first argument of the reference's `dct2_2d`, src/dct2d.rs:83-89):
    /// plans and twiddle tables are cached by the library's per-thread context.
    #[derive(Default)]
    pub struct DctPlanner;
    impl DctPlanner {
        pub fn new() -> Self { DctPlanner }
    }
    #[derive(Clone, Copy, Debug, PartialEq)]
    pub enum Type { DCT2, DCT2Orthogonal, DCT3 }
    pub fn dct2_2d(_planner: &mut DctPlanner, transform_type: Type, width: usize, height: usize, data: &mut [f32]) {
        assert_eq!(data.len(), width * height);
        let t = match transform_type { Type::DCT2 => 0, Type::DCT2Orthogonal => 1, Type::DCT3 => 2 };
        crate::ffi::check(unsafe { crate::ffi::ssw_dct2_2d(crate::ctx().0, t, width as u32, height as u32, data.as_mut_ptr()) });
    }
}
