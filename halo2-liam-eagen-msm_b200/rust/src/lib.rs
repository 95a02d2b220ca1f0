//! Drop-in module layout of levs57/halo2-liam-eagen-msm for the witness path
//! (reference: src/lib.rs:1-11).  `config`, `scripts` and `testing_stuff` are not part of the path and stay as they
//! are in the reference crate; they only consume the functions re-exported here.
pub mod ffi;
pub mod gpu;
pub mod negbase_utils;
pub mod regular_functions_utils;
pub mod argument_witness_calc;
