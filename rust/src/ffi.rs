//! Raw bindings of include/ssw.h (one `extern "C"` item per entry point used by the shim).
#![allow(non_camel_case_types)]
use std::os::raw::{c_char, c_int, c_void};

#[repr(C)]
#[derive(Clone, Copy, Debug)]
pub struct ssw_config {
    pub method: i32,   // SSW_METHOD_OPTION{1,2,3}
    pub alpha: f32,    // OptionN(alpha)
    pub ordering: i32, // 0 Energy, 1 EnergyOrthogonal, 2 Legacy
}

#[repr(C)] pub struct ssw_ctx { _p: [u8; 0] }
#[repr(C)] pub struct ssw_writer { _p: [u8; 0] }
#[repr(C)] pub struct ssw_reader { _p: [u8; 0] }
#[repr(C)] pub struct ssw_bank { _p: [u8; 0] }

extern "C" {
    pub fn ssw_last_error() -> *const c_char;
    pub fn ssw_ctx_create(device: c_int, out: *mut *mut ssw_ctx) -> c_int;
    pub fn ssw_ctx_destroy(ctx: *mut ssw_ctx) -> c_int;

    pub fn ssw_dct2_2d(ctx: *mut ssw_ctx, ty: c_int, width: u32, height: u32, data: *mut f32) -> c_int;

    pub fn ssw_writer_new_rgb32f(ctx: *mut ssw_ctx, rgb: *const f32, w: u32, h: u32, cfg: *const ssw_config,
                                 out: *mut *mut ssw_writer) -> c_int;
    pub fn ssw_writer_new_rgb8(ctx: *mut ssw_ctx, rgb: *const u8, w: u32, h: u32, cfg: *const ssw_config,
                               out: *mut *mut ssw_writer) -> c_int;
    pub fn ssw_writer_embed(w: *mut ssw_writer, marks: *const *const f32, lens: *const usize, n_marks: usize) -> c_int;
    pub fn ssw_writer_coefficients(w: *mut ssw_writer, out: *mut f32) -> c_int;
    pub fn ssw_writer_result_rgb32f(w: *mut ssw_writer, out: *mut f32) -> c_int;
    pub fn ssw_writer_result_rgb8(w: *mut ssw_writer, out: *mut u8) -> c_int;
    pub fn ssw_writer_destroy(w: *mut ssw_writer) -> c_int;

    pub fn ssw_reader_base_rgb32f(ctx: *mut ssw_ctx, rgb: *const f32, w: u32, h: u32, cfg: *const ssw_config,
                                  out: *mut *mut ssw_reader) -> c_int;
    pub fn ssw_reader_derived_rgb32f(ctx: *mut ssw_ctx, rgb: *const f32, w: u32, h: u32,
                                     out: *mut *mut ssw_reader) -> c_int;
    pub fn ssw_reader_extract(base: *mut ssw_reader, derived: *mut ssw_reader, out: *mut f32, n: usize) -> c_int;
    pub fn ssw_reader_coefficients(r: *mut ssw_reader, out: *mut f32) -> c_int;
    pub fn ssw_reader_indices(r: *mut ssw_reader, out: *mut u64, n: usize) -> c_int;
    pub fn ssw_reader_destroy(r: *mut ssw_reader) -> c_int;

    pub fn ssw_similarity(ctx: *mut ssw_ctx, extracted: *const f32, mark: *const f32, n: usize, out: *mut f32) -> c_int;
    pub fn ssw_mark_generate_normal(ctx: *mut ssw_ctx, seed: u64, n: usize, out: *mut f32) -> c_int;

    pub fn ssw_bank_create(ctx: *mut ssw_ctx, marks: *const f32, n_marks: usize, n: usize, out: *mut *mut ssw_bank) -> c_int;
    pub fn ssw_bank_similarity(bank: *mut ssw_bank, extracted: *const f32, n_extracted: usize, out: *mut f32) -> c_int;
    pub fn ssw_bank_destroy(bank: *mut ssw_bank) -> c_int;
}

/// The reference panics on misuse; every non-zero status becomes a panic carrying libssw's message.
pub fn check(status: c_int) {
    if status != 0 {
        let msg = unsafe { std::ffi::CStr::from_ptr(ssw_last_error()) }.to_string_lossy().into_owned();
        panic!("libssw error {}: {}", status, msg);
    }
}

pub type Opaque = c_void;
