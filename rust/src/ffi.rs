//! src/ffi.rs -- bindings to librla_b200 (include/rla_b200.h).  Added to rulinalg by the integration;
//! `mod ffi;` goes into src/lib.rs next to `extern crate matrixmultiply;` (src/lib.rs:90).
#![allow(non_camel_case_types)]

use error::{Error, ErrorKind};

pub const RLA_OK: i32 = 0;
pub const RLA_ERR_SINGULAR: i32 = 1;
pub const RLA_ERR_NOT_POSITIVE: i32 = 3;

#[repr(C)]
pub struct rla_lu_handle {
    _private: [u8; 0],
}

extern "C" {
    /// Same argument order and meaning as matrixmultiply::dgemm (src/matrix/mat_mul.rs:57-67).
    pub fn rla_dgemm(m: usize, k: usize, n: usize, alpha: f64,
                     a: *const f64, rsa: isize, csa: isize,
                     b: *const f64, rsb: isize, csb: isize,
                     beta: f64, c: *mut f64, rsc: isize, csc: isize) -> i32;
    /// Same as matrixmultiply::sgemm (src/matrix/mat_mul.rs:33-43).
    pub fn rla_sgemm(m: usize, k: usize, n: usize, alpha: f32,
                     a: *const f32, rsa: isize, csa: isize,
                     b: *const f32, rsb: isize, csb: isize,
                     beta: f32, c: *mut f32, rsc: isize, csc: isize) -> i32;
    pub fn rla_dgetrf(n: usize, lu: *mut f64, perm: *mut usize) -> i32;
    pub fn rla_sgetrf(n: usize, lu: *mut f32, perm: *mut usize) -> i32;
    pub fn rla_dgetrs(n: usize, lu: *const f64, perm: *const usize, b: *mut f64) -> i32;
    pub fn rla_sgetrs(n: usize, lu: *const f32, perm: *const usize, b: *mut f32) -> i32;
    pub fn rla_dgetrf_keep(n: usize, lu: *mut f64, perm: *mut usize, out: *mut *mut rla_lu_handle) -> i32;
    pub fn rla_dlu_solve(h: *const rla_lu_handle, b: *mut f64) -> i32;
    pub fn rla_lu_free(h: *mut rla_lu_handle);
    /// SURVEY 8f "next" rows: inverse, triangular solves, matrix-vector product, Cholesky (cholesky.rs:116-233).
    pub fn rla_dgetri(n: usize, lu: *const f64, perm: *const usize, inv: *mut f64) -> i32;
    pub fn rla_sgetri(n: usize, lu: *const f32, perm: *const usize, inv: *mut f32) -> i32;
    pub fn rla_dtrsv(lower: i32, n: usize, a: *const f64, rs: isize, x: *mut f64) -> i32;
    pub fn rla_strsv(lower: i32, n: usize, a: *const f32, rs: isize, x: *mut f32) -> i32;
    pub fn rla_dgemv(m: usize, n: usize, a: *const f64, rs: isize, x: *const f64, y: *mut f64) -> i32;
    pub fn rla_sgemv(m: usize, n: usize, a: *const f32, rs: isize, x: *const f32, y: *mut f32) -> i32;
    pub fn rla_dpotrf(n: usize, a: *mut f64) -> i32;
    pub fn rla_spotrf(n: usize, a: *mut f32) -> i32;
    pub fn rla_dpotrs(n: usize, l: *const f64, b: *mut f64) -> i32;
    pub fn rla_spotrs(n: usize, l: *const f32, b: *mut f32) -> i32;
    pub fn rla_dpotri(n: usize, l: *const f64, inv: *mut f64) -> i32;
    pub fn rla_spotri(n: usize, l: *const f32, inv: *mut f32) -> i32;
    pub fn rla_strerror(status: i32) -> *const ::std::os::raw::c_char;
    /// One host process, N GPUs: after `rla_set_devices(n)` large `rla_dgemm` / `rla_sgemm` / `rla_dgetrf` calls shard over
    /// GPUs 0..n-1 (row panels + peer-memory fan-out of B; 1D block-cyclic LU).  Results are bit-identical to n = 1.
    pub fn rla_set_devices(n_gpus: i32) -> i32;
    pub fn rla_get_devices() -> i32;
    pub fn rla_device_count() -> i32;
    /// Returns the calling thread's streams / buffers / staging rings and the multi-GPU contexts (also runs at thread exit).
    pub fn rla_shutdown() -> i32;
    /// Held operands: between hold and release the byte range is promised immutable, so products that
    /// read an operand inside it keep the operand's device copy and skip its upload on later calls.
    pub fn rla_operand_hold(host: *const ::std::os::raw::c_void, bytes: usize) -> i32;
    pub fn rla_operand_release(host: *const ::std::os::raw::c_void) -> i32;
    pub fn rla_operand_resident_bytes() -> usize;
}

/// `let _g = ffi::Held::new(&b);` -- while the guard lives, `b` is shared-borrowed (so nothing can mutate or drop it: the
/// borrow checker enforces the promise the C ABI asks for) and every `&a * &b` re-uses b's copy in HBM instead of
/// uploading it again (the repeated products of lu.rs:789,907, eigen.rs:114-148).  Dropping the guard frees the copy.
pub struct Held<'a, T: 'a> {
    data: &'a [T],
}
impl<'a, T> Held<'a, T> {
    pub fn new(m: &'a ::matrix::Matrix<T>) -> Held<'a, T> {
        let data = m.data().as_slice();
        let st = unsafe { rla_operand_hold(data.as_ptr() as *const _, data.len() * ::std::mem::size_of::<T>()) };
        assert!(st == RLA_OK, "librla_b200: rla_operand_hold -> {}", st);
        Held { data: data }
    }
}
impl<'a, T> Drop for Held<'a, T> {
    fn drop(&mut self) {
        unsafe { rla_operand_release(self.data.as_ptr() as *const _) };
    }
}

/// Opt-in for multi-GPU boxes, e.g. from the embedding application's start-up:
/// `ffi::use_gpus(8)` -> every later `&a * &b` / `PartialPivLu::decompose` large enough to profit is sharded.
pub fn use_gpus(n: i32) {
    let st = unsafe { rla_set_devices(n) };
    if st != RLA_OK {
        let msg = unsafe { ::std::ffi::CStr::from_ptr(rla_strerror(st)) };
        panic!("librla_b200: rla_set_devices({}) -> {} ({})", n, st, msg.to_string_lossy())
    }
}

/// Cholesky::decompose statuses (cholesky.rs:151-158): both are ErrorKind::DecompFailure.
pub fn check_potrf(status: i32) -> Result<(), Error> {
    match status {
        RLA_OK => Ok(()),
        RLA_ERR_SINGULAR => Err(Error::new(ErrorKind::DecompFailure, "Matrix is singular to working precision.")),
        RLA_ERR_NOT_POSITIVE => Err(Error::new(ErrorKind::DecompFailure, "Diagonal entries of matrix are not all positive.")),
        s => check(s, ""),
    }
}

/// Numerical statuses become `Err(Error)`, environment failures (no device, CUDA error) panic:
/// north_star forbids a CPU fallback, and the reference has no error kind for them.
pub fn check(status: i32, singular_msg: &'static str) -> Result<(), Error> {
    match status {
        RLA_OK => Ok(()),
        RLA_ERR_SINGULAR => Err(Error::new(ErrorKind::DivByZero, singular_msg)),
        s => {
            let msg = unsafe { ::std::ffi::CStr::from_ptr(rla_strerror(s)) };
            panic!("librla_b200: status {} ({})", s, msg.to_string_lossy())
        }
    }
}
