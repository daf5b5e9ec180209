//! Rust binding of libqwgpu for quickwit-search: the two seams of SURVEY.md §3.4 / §8b.
pub mod context;
pub mod ffi;
pub mod invoker;
pub mod service;

pub use context::GpuContext;
pub use invoker::GpuLeafSearchInvoker;
pub use service::GpuSearchService;
