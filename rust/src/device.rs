//! `src/device.rs` -- the device mirror a solver struct owns: one `libsla_b200` context (created lazily, freed in
//! `Drop`, NOT copied by `Clone`: a cloned solver re-uploads on its first solve), the widening of `u16` indices to the
//! `u32` the C ABI speaks, and the calls that replace the bodies of `KhoslaSolver::solve` (src/ksparse.rs:153-251) and
//! `ForwardAuctionSolver::solve_with_params` (src/symmetric.rs:217-332).
//!
//! Wholly new code (nothing in the reference corresponds); not compiled in this repository (no Rust toolchain in the
//! build image).  The Python and C++ mirrors of the same logic (`sparse_linear_assignment_b200/solver.py`,
//! `include/sla.hpp`) are what the tests here run.
use crate::ffi;
use crate::solution::{AuctionSolution, UnsignedInt};
use anyhow::{anyhow, Result};
use num_traits::AsPrimitive;
use std::os::raw::c_int;

pub(crate) struct DeviceMirror {
    ctx: *mut ffi::sla_ctx,
    /// the host CSR changed since the last upload (`init`, `add_value`, `extend_from_values`, any `*_mut` accessor)
    dirty: bool,
    /// the upload of the pending solve already negated the host `values` in place (solver.rs:214-216)
    pre_negated: bool,
    has_solution: bool,
    rows32: Vec<u32>,
    cols32: Vec<u32>,
    p2o32: Vec<u32>,
    o2p32: Vec<u32>,
}

impl Default for DeviceMirror {
    fn default() -> Self {
        DeviceMirror {
            ctx: std::ptr::null_mut(),
            dirty: true,
            pre_negated: false,
            has_solution: false,
            rows32: Vec::new(),
            cols32: Vec::new(),
            p2o32: Vec::new(),
            o2p32: Vec::new(),
        }
    }
}

/// `#[derive(Clone)]` on the solvers: the clone owns no device state yet.
impl Clone for DeviceMirror {
    fn clone(&self) -> Self {
        DeviceMirror::default()
    }
}

impl Drop for DeviceMirror {
    fn drop(&mut self) {
        if !self.ctx.is_null() {
            unsafe { ffi::sla_ctx_destroy(self.ctx) }
        }
    }
}

// One CUDA stream per context, `&mut self` on every mutating call: same threading contract as the reference's Vecs.
unsafe impl Send for DeviceMirror {}

/// `&[I]` as `*const u32`: in place for `u32`, through a scratch vector for `u16`.
fn widen<I: UnsignedInt>(src: &[I], scratch: &mut Vec<u32>) -> *const u32 {
    if std::mem::size_of::<I>() == 4 {
        src.as_ptr() as *const u32
    } else {
        scratch.clear();
        scratch.extend(src.iter().map(|x| {
            let u: usize = (*x).as_();
            u as u32
        }));
        scratch.as_ptr()
    }
}

/// `u32` results into a `Vec<I>`; `SLA_NONE` (`u32::MAX`) becomes `I::max_value()` (solution.rs:27-34).
fn narrow<I: UnsignedInt>(src: &[u32], dst: &mut Vec<I>) {
    dst.clear();
    dst.extend(src.iter().map(|&x| I::from_u32(x).unwrap_or_else(I::max_value)));
}

impl DeviceMirror {
    pub(crate) fn mark_dirty(&mut self) {
        self.dirty = true;
    }

    fn error(&self, rc: c_int) -> anyhow::Error {
        let msg = unsafe { std::ffi::CStr::from_ptr(ffi::sla_last_error(self.ctx)) };
        anyhow!("libsla_b200 error {}: {}", rc, msg.to_string_lossy())
    }

    fn check(&self, rc: c_int) -> Result<()> {
        if rc == ffi::SLA_OK {
            Ok(())
        } else {
            Err(self.error(rc))
        }
    }

    fn ensure_ctx(&mut self, rows: usize, cols: usize, arcs: usize) -> Result<()> {
        if self.ctx.is_null() {
            let mut ctx = std::ptr::null_mut();
            let rc = unsafe { ffi::sla_ctx_create(0, rows, cols, arcs, &mut ctx) };
            if rc != ffi::SLA_OK {
                let msg = unsafe { std::ffi::CStr::from_ptr(ffi::sla_last_error(std::ptr::null())) };
                return Err(anyhow!("sla_ctx_create failed ({}): {}", rc, msg.to_string_lossy()));
            }
            self.ctx = ctx;
            self.dirty = true;
        }
        Ok(())
    }

    /// Mirrors the host CSR into HBM when it changed.  When the coming solve will flip the sign
    /// (`maximize ^ (values[0] >= 0.0)`, solver.rs:209-216) the upload also negates the host `values` in place on the
    /// library's worker threads, overlapped with the PCIe copies and the solve (joined inside the solve call).
    fn sync<I: UnsignedInt>(
        &mut self,
        num_rows: I,
        num_cols: I,
        i_starts_stops: &[I],
        column_indices: &[I],
        values: &mut Vec<f64>,
        maximize: bool,
    ) -> Result<()> {
        let (n, m): (usize, usize) = (num_rows.as_(), num_cols.as_());
        self.ensure_ctx(n, m, column_indices.len())?;
        if !self.dirty {
            return Ok(());
        }
        let rows = widen(&i_starts_stops[..n + 1], &mut self.rows32);
        let cols = widen(column_indices, &mut self.cols32);
        let flip = maximize ^ (values.first().copied().unwrap_or(0.0) >= 0.0);
        let rc = unsafe {
            if flip {
                ffi::sla_upload_csr_negating(self.ctx, n as u32, m as u32, rows, cols, values.as_mut_ptr(), values.len() as u64, 0)
            } else {
                ffi::sla_upload_csr(self.ctx, n as u32, m as u32, rows, cols, values.as_ptr(), values.len() as u64)
            }
        };
        self.check(rc)?;
        self.pre_negated = flip;
        self.dirty = false;
        self.has_solution = false;
        Ok(())
    }

    /// Output pointers for the two assignment vectors: the caller's `Vec<u32>` in place, scratch for `u16`.
    fn outputs<I: UnsignedInt>(&mut self, solution: &mut AuctionSolution<I>, n: usize, m: usize) -> (*mut u32, *mut u32) {
        if std::mem::size_of::<I>() == 4 {
            solution.person_to_object.resize(n, I::max_value()); // post-condition of solver.rs:221-228
            solution.object_to_person.resize(m, I::max_value());
            (solution.person_to_object.as_mut_ptr() as *mut u32, solution.object_to_person.as_mut_ptr() as *mut u32)
        } else {
            self.p2o32.resize(n, ffi::SLA_NONE);
            self.o2p32.resize(m, ffi::SLA_NONE);
            (self.p2o32.as_mut_ptr(), self.o2p32.as_mut_ptr())
        }
    }

    fn finish<I: UnsignedInt>(&mut self, solution: &mut AuctionSolution<I>, st: &ffi::sla_stats, values: &mut Vec<f64>) {
        if std::mem::size_of::<I>() != 4 {
            narrow(&self.p2o32, &mut solution.person_to_object);
            narrow(&self.o2p32, &mut solution.object_to_person);
        }
        // in-place sign normalisation of the host copy (solver.rs:214-216) unless the upload already did it
        if st.values_negated == 1 && !self.pre_negated {
            unsafe { ffi::sla_host_negate_f64(values.as_mut_ptr(), values.len(), 8) };
        }
        self.pre_negated = false;
        self.has_solution = true;
        solution.num_unassigned = I::from_u32(st.num_unassigned).unwrap_or_else(I::max_value);
        solution.eps = st.eps;
    }

    /// Body of `KhoslaSolver::solve` after `validate_input` (src/ksparse.rs:153-251).  Prices stay in HBM.
    #[allow(clippy::too_many_arguments)]
    pub(crate) fn khosla_solve<I: UnsignedInt>(
        &mut self,
        num_rows: I,
        num_cols: I,
        i_starts_stops: &[I],
        column_indices: &[I],
        values: &mut Vec<f64>,
        solution: &mut AuctionSolution<I>,
        maximize: bool,
        eps: Option<f64>,
    ) -> Result<ffi::sla_stats> {
        self.sync(num_rows, num_cols, i_starts_stops, column_indices, values, maximize)?;
        let (n, m): (usize, usize) = (num_rows.as_(), num_cols.as_());
        let (p2o, o2p) = self.outputs(solution, n, m);
        let mut st = ffi::sla_stats::default();
        let rc = unsafe {
            ffi::sla_khosla_solve(self.ctx, maximize as c_int, eps.unwrap_or(f64::NAN), p2o, o2p, std::ptr::null_mut(), &mut st)
        };
        if rc != ffi::SLA_OK {
            self.pre_negated = false; // host and device agree about the sign; only the marker must not survive
            return Err(self.error(rc));
        }
        self.finish(solution, &st, values);
        Ok(st)
    }

    /// Body of `ForwardAuctionSolver::solve_with_params` after `validate_input` (src/symmetric.rs:217-332).
    #[allow(clippy::too_many_arguments)]
    pub(crate) fn forward_solve<I: UnsignedInt>(
        &mut self,
        num_rows: I,
        num_cols: I,
        i_starts_stops: &[I],
        column_indices: &[I],
        values: &mut Vec<f64>,
        solution: &mut AuctionSolution<I>,
        maximize: bool,
        eps: Option<f64>,
        start_eps: Option<f64>,
        max_iterations: u32,
    ) -> Result<ffi::sla_stats> {
        self.sync(num_rows, num_cols, i_starts_stops, column_indices, values, maximize)?;
        let (n, m): (usize, usize) = (num_rows.as_(), num_cols.as_());
        let (p2o, o2p) = self.outputs(solution, n, m);
        let mut st = ffi::sla_stats::default();
        let rc = unsafe {
            ffi::sla_forward_solve(
                self.ctx,
                maximize as c_int,
                eps.unwrap_or(f64::NAN),
                start_eps.unwrap_or(f64::NAN),
                max_iterations.max(1), // Some(0) behaves like Some(1): the check runs after the first round (symmetric.rs:326)
                p2o,
                o2p,
                std::ptr::null_mut(),
                &mut st,
            )
        };
        if rc != ffi::SLA_OK {
            self.pre_negated = false;
            return Err(self.error(rc));
        }
        self.finish(solution, &st, values);
        Ok(st)
    }

    /// Lazy fill of `prices()` (solver.rs:28): the final prices stay in HBM until somebody reads them.
    pub(crate) fn download_prices(&self, num_cols: usize) -> Vec<f64> {
        let mut prices = vec![0.0; num_cols];
        if self.has_solution && !self.ctx.is_null() {
            let rc = unsafe { ffi::sla_download_solution(self.ctx, std::ptr::null_mut(), std::ptr::null_mut(), prices.as_mut_ptr()) };
            debug_assert_eq!(rc, ffi::SLA_OK);
        }
        prices
    }
}
