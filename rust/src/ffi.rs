//! Raw bindings of include/ssw.h: the reference-facing entry points (one per public item of the crate), the batch /
//! bank / asynchronous entry points and the sharded-frame entry points.  Signatures are transcribed from the header;
//! `*_dev` pointers are CUDA device pointers on the context's device.
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
#[repr(C)] pub struct ssw_sharded { _p: [u8; 0] }

pub const SSW_SHARDED_ID_BYTES: usize = 128;

extern "C" {
    pub fn ssw_last_error() -> *const c_char;
    pub fn ssw_version() -> *const c_char;
    pub fn ssw_ctx_create(device: c_int, out: *mut *mut ssw_ctx) -> c_int;
    pub fn ssw_ctx_create_on_stream(device: c_int, stream: *mut c_void, out: *mut *mut ssw_ctx) -> c_int;
    pub fn ssw_ctx_destroy(ctx: *mut ssw_ctx) -> c_int;
    pub fn ssw_ctx_synchronize(ctx: *mut ssw_ctx) -> c_int;
    pub fn ssw_ctx_stream(ctx: *mut ssw_ctx) -> *mut c_void;
    pub fn ssw_ctx_set_trace(ctx: *mut ssw_ctx, dev_buf: *mut c_void) -> c_int;
    pub fn ssw_ctx_marker(ctx: *mut ssw_ctx, marker: *mut u64) -> c_int;
    pub fn ssw_ctx_wait_marker(ctx: *mut ssw_ctx, marker: u64) -> c_int;
    pub fn ssw_ctx_last_topk_fallbacks(ctx: *mut ssw_ctx) -> c_int;
    pub fn ssw_host_alloc(bytes: usize, out: *mut *mut c_void) -> c_int;
    pub fn ssw_host_free(p: *mut c_void) -> c_int;

    // dct2d::dct2_2d
    pub fn ssw_dct2_2d(ctx: *mut ssw_ctx, ty: c_int, width: u32, height: u32, data: *mut f32) -> c_int;
    pub fn ssw_dct2_2d_dev(ctx: *mut ssw_ctx, ty: c_int, width: u32, height: u32, data_dev: *mut f32) -> c_int;
    // yiq
    pub fn ssw_rgb32f_to_yiq(ctx: *mut ssw_ctx, rgb: *const f32, w: u32, h: u32, y: *mut f32, i: *mut f32, q: *mut f32) -> c_int;
    pub fn ssw_yiq_to_rgb32f(ctx: *mut ssw_ctx, y: *const f32, i: *const f32, q: *const f32, w: u32, h: u32, rgb: *mut f32) -> c_int;

    // Writer
    pub fn ssw_writer_new_rgb32f(ctx: *mut ssw_ctx, rgb: *const f32, w: u32, h: u32, cfg: *const ssw_config, out: *mut *mut ssw_writer) -> c_int;
    pub fn ssw_writer_new_rgb8(ctx: *mut ssw_ctx, rgb: *const u8, w: u32, h: u32, cfg: *const ssw_config, out: *mut *mut ssw_writer) -> c_int;
    pub fn ssw_writer_new_rgb8_dev(ctx: *mut ssw_ctx, rgb_dev: *const u8, w: u32, h: u32, cfg: *const ssw_config, out: *mut *mut ssw_writer) -> c_int;
    pub fn ssw_writer_embed(w: *mut ssw_writer, marks: *const *const f32, lens: *const usize, n_marks: usize) -> c_int;
    pub fn ssw_writer_coefficients(w: *mut ssw_writer, out: *mut f32) -> c_int;
    pub fn ssw_writer_indices(w: *mut ssw_writer, out: *mut u64, n: usize) -> c_int;
    pub fn ssw_writer_result_rgb32f(w: *mut ssw_writer, out: *mut f32) -> c_int;
    pub fn ssw_writer_result_rgb8(w: *mut ssw_writer, out: *mut u8) -> c_int;
    pub fn ssw_writer_result_rgb8_dev(w: *mut ssw_writer, out_dev: *mut u8) -> c_int;
    pub fn ssw_writer_destroy(w: *mut ssw_writer) -> c_int;

    // Reader
    pub fn ssw_reader_base_rgb32f(ctx: *mut ssw_ctx, rgb: *const f32, w: u32, h: u32, cfg: *const ssw_config, out: *mut *mut ssw_reader) -> c_int;
    pub fn ssw_reader_base_rgb8(ctx: *mut ssw_ctx, rgb: *const u8, w: u32, h: u32, cfg: *const ssw_config, out: *mut *mut ssw_reader) -> c_int;
    pub fn ssw_reader_base_rgb8_dev(ctx: *mut ssw_ctx, rgb_dev: *const u8, w: u32, h: u32, cfg: *const ssw_config, out: *mut *mut ssw_reader) -> c_int;
    pub fn ssw_reader_derived_rgb32f(ctx: *mut ssw_ctx, rgb: *const f32, w: u32, h: u32, out: *mut *mut ssw_reader) -> c_int;
    pub fn ssw_reader_derived_rgb8(ctx: *mut ssw_ctx, rgb: *const u8, w: u32, h: u32, out: *mut *mut ssw_reader) -> c_int;
    pub fn ssw_reader_derived_rgb8_dev(ctx: *mut ssw_ctx, rgb_dev: *const u8, w: u32, h: u32, out: *mut *mut ssw_reader) -> c_int;
    pub fn ssw_reader_extract(base: *mut ssw_reader, derived: *mut ssw_reader, out: *mut f32, n: usize) -> c_int;
    pub fn ssw_reader_extract_dev(base: *mut ssw_reader, derived: *mut ssw_reader, out_dev: *mut f32, n: usize) -> c_int;
    pub fn ssw_reader_coefficients(r: *mut ssw_reader, out: *mut f32) -> c_int;
    pub fn ssw_reader_indices(r: *mut ssw_reader, out: *mut u64, n: usize) -> c_int;
    pub fn ssw_reader_destroy(r: *mut ssw_reader) -> c_int;

    // Tester / marks / bank
    pub fn ssw_similarity(ctx: *mut ssw_ctx, extracted: *const f32, mark: *const f32, n: usize, out: *mut f32) -> c_int;
    pub fn ssw_mark_generate_normal(ctx: *mut ssw_ctx, seed: u64, n: usize, out: *mut f32) -> c_int;
    pub fn ssw_bank_create(ctx: *mut ssw_ctx, marks: *const f32, n_marks: usize, n: usize, out: *mut *mut ssw_bank) -> c_int;
    pub fn ssw_bank_create_normal(ctx: *mut ssw_ctx, seed: u64, n_marks: usize, n: usize, out: *mut *mut ssw_bank) -> c_int;
    pub fn ssw_bank_row(bank: *mut ssw_bank, index: usize, out: *mut f32) -> c_int;
    pub fn ssw_bank_similarity(bank: *mut ssw_bank, extracted: *const f32, n_extracted: usize, out: *mut f32) -> c_int;
    pub fn ssw_bank_similarity_dev(bank: *mut ssw_bank, extracted_dev: *const f32, n_extracted: usize, out_dev: *mut f32) -> c_int;
    pub fn ssw_bank_destroy(bank: *mut ssw_bank) -> c_int;

    // fused batches: device-resident, host buffers, asynchronous host buffers
    pub fn ssw_embed_batch_rgb8_dev(ctx: *mut ssw_ctx, rgb_dev: *const u8, w: u32, h: u32, batch: u32, cfg: *const ssw_config,
                                    marks_dev: *const f32, n: usize, out_rgb_dev: *mut u8) -> c_int;
    pub fn ssw_extract_batch_rgb8_dev(ctx: *mut ssw_ctx, base_dev: *const u8, derived_dev: *const u8, w: u32, h: u32, batch: u32,
                                      cfg: *const ssw_config, n: usize, extracted_dev: *mut f32, marks_dev: *const f32, sim_dev: *mut f32) -> c_int;
    pub fn ssw_embed_batch_rgb8(ctx: *mut ssw_ctx, rgb: *const u8, w: u32, h: u32, batch: u32, cfg: *const ssw_config,
                                marks: *const f32, n: usize, out_rgb: *mut u8) -> c_int;
    pub fn ssw_extract_batch_rgb8(ctx: *mut ssw_ctx, base: *const u8, derived: *const u8, w: u32, h: u32, batch: u32,
                                  cfg: *const ssw_config, n: usize, extracted: *mut f32, marks: *const f32, sim: *mut f32) -> c_int;
    pub fn ssw_embed_batch_rgb8_async(ctx: *mut ssw_ctx, rgb: *const u8, w: u32, h: u32, batch: u32, cfg: *const ssw_config,
                                      marks: *const f32, n: usize, out_rgb: *mut u8) -> c_int;
    pub fn ssw_extract_batch_rgb8_async(ctx: *mut ssw_ctx, base: *const u8, derived: *const u8, w: u32, h: u32, batch: u32,
                                        cfg: *const ssw_config, n: usize, extracted: *mut f32, marks: *const f32, sim: *mut f32) -> c_int;

    // sharded single frames (one process per GPU)
    pub fn ssw_sharded_unique_id(id_out: *mut c_void) -> c_int;
    pub fn ssw_sharded_create(ctx: *mut ssw_ctx, id: *const c_void, rank: c_int, world: c_int, width: u32, height: u32,
                              out: *mut *mut ssw_sharded) -> c_int;
    pub fn ssw_sharded_destroy(s: *mut ssw_sharded) -> c_int;
    pub fn ssw_sharded_embed_rgb8_dev(s: *mut ssw_sharded, rows_dev: *const u8, cfg: *const ssw_config, mark_dev: *const f32, n: usize,
                                      out_rows_dev: *mut u8) -> c_int;
    pub fn ssw_sharded_extract_rgb8_dev(s: *mut ssw_sharded, base_rows_dev: *const u8, derived_rows_dev: *const u8, cfg: *const ssw_config,
                                        n: usize, extracted_dev: *mut f32) -> c_int;
    pub fn ssw_sharded_indices(s: *mut ssw_sharded, out: *mut u32, n: usize) -> c_int;
    pub fn ssw_sharded_coefficients(s: *mut ssw_sharded, which: c_int, out: *mut f32) -> c_int;
    pub fn ssw_sharded_overflow(s: *mut ssw_sharded, overflowed: *mut c_int) -> c_int;
}

/// The reference panics on misuse; every non-zero status becomes a panic carrying libssw's message.
pub fn check(status: c_int) {
    if status != 0 {
        let msg = unsafe { std::ffi::CStr::from_ptr(ssw_last_error()) }.to_string_lossy().into_owned();
        panic!("libssw error {}: {}", status, msg);
    }
}
