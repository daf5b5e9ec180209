//! Seam B: the GPU as a `LambdaLeafSearchInvoker` (quickwit-search/src/invoker.rs:27-38). The reference keeps its
//! request rewriting, partial-result cache, permit scheduling, per-split failure handling and `IncrementalCollector`
//! merge (leaf.rs:1449-1598) and hands the splits it would have offloaded to a Lambda to the GPU instead.

use std::sync::Arc;

use async_trait::async_trait;
use prost::Message;
use quickwit_proto::search::{LambdaSearchResponses, LambdaSingleSplitResult, LeafSearchRequest};
use quickwit_common::thread_pool::run_cpu_intensive;
use quickwit_search::{LambdaLeafSearchInvoker, SearchError};

use crate::context::GpuContext;
use crate::ffi;

pub struct GpuLeafSearchInvoker {
    ctx: Arc<GpuContext>,
}

impl GpuLeafSearchInvoker {
    pub fn new(ctx: Arc<GpuContext>) -> Self {
        GpuLeafSearchInvoker { ctx }
    }
}

#[async_trait]
impl LambdaLeafSearchInvoker for GpuLeafSearchInvoker {
    /// One `LambdaSingleSplitResult` per split of the request; a split that is not resident or fails on the device
    /// comes back as an error entry (retryable), never as an `Err` of the whole call.
    async fn invoke_leaf_search(&self, request: LeafSearchRequest) -> Result<Vec<LambdaSingleSplitResult>, SearchError> {
        let request_bytes = request.encode_to_vec();
        let ctx = self.ctx.clone();
        // the call blocks on the device: keep it off the async runtime, like the reference's CPU searches
        let response_bytes = run_cpu_intensive(move || unsafe { ffi::bytes_call(ffi::qwgpu_invoke_leaf_search, ctx.raw(), &request_bytes) })
            .await
            .map_err(|panicked| SearchError::Internal(format!("gpu leaf search task failed: {panicked}")))??;
        let responses = LambdaSearchResponses::decode(response_bytes.as_slice())
            .map_err(|err| SearchError::Internal(format!("undecodable LambdaSearchResponses from libqwgpu: {err}")))?;
        Ok(responses.split_results)
    }
}
