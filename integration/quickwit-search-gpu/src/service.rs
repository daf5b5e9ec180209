//! Seam A: a `SearchService` whose `leaf_search` runs on the GPU (quickwit-search/src/service.rs:64-136); every other
//! method is the wrapped `SearchServiceImpl`'s. `LeafSearchResponse.intermediate_aggregation_result` produced by the
//! GPU leaf is libqwgpu's own layout (not tantivy's postcard): a root that receives it finalizes through
//! `qwgpu_merge_leaf_responses` + `qwgpu_finalize_aggregation`.

use std::sync::Arc;

use async_trait::async_trait;
use prost::Message;
use quickwit_proto::search::{
    FetchDocsRequest, FetchDocsResponse, GetKvRequest, LeafListFieldsRequest, LeafListTermsRequest, LeafListTermsResponse,
    LeafSearchRequest, LeafSearchResponse, ListFieldsRequest, ListFieldsResponse, ListTermsRequest, ListTermsResponse,
    PutKvRequest, ReportSplitsRequest, ReportSplitsResponse, ScrollRequest, SearchPlanResponse, SearchRequest, SearchResponse,
};
use quickwit_common::thread_pool::run_cpu_intensive;
use quickwit_search::{SearchError, SearchService};

use crate::context::GpuContext;
use crate::ffi;

pub struct GpuSearchService<S: SearchService> {
    inner: S,
    ctx: Arc<GpuContext>,
}

impl<S: SearchService> GpuSearchService<S> {
    pub fn new(inner: S, ctx: Arc<GpuContext>) -> Self {
        GpuSearchService { inner, ctx }
    }
}

#[async_trait]
impl<S: SearchService> SearchService for GpuSearchService<S> {
    async fn leaf_search(&self, request: LeafSearchRequest) -> quickwit_search::Result<LeafSearchResponse> {
        let request_bytes = request.encode_to_vec();
        let ctx = self.ctx.clone();
        let outcome = run_cpu_intensive(move || unsafe { ffi::bytes_call(ffi::qwgpu_leaf_search, ctx.raw(), &request_bytes) })
            .await
            .map_err(|panicked| SearchError::Internal(format!("gpu leaf search task failed: {panicked}")))?;
        match outcome {
            Ok(response_bytes) => LeafSearchResponse::decode(response_bytes.as_slice())
                .map_err(|err| SearchError::Internal(format!("undecodable LeafSearchResponse from libqwgpu: {err}"))),
            // a request shape the GPU path does not take is deterministic: serve it on the CPU searcher
            Err(SearchError::Internal(msg)) if msg.starts_with("unsupported on the GPU path") => self.inner.leaf_search(request).await,
            Err(err) => Err(err),
        }
    }

    async fn root_search(&self, request: SearchRequest) -> quickwit_search::Result<SearchResponse> {
        self.inner.root_search(request).await
    }
    async fn fetch_docs(&self, request: FetchDocsRequest) -> quickwit_search::Result<FetchDocsResponse> {
        self.inner.fetch_docs(request).await
    }
    async fn root_list_terms(&self, request: ListTermsRequest) -> quickwit_search::Result<ListTermsResponse> {
        self.inner.root_list_terms(request).await
    }
    async fn leaf_list_terms(&self, request: LeafListTermsRequest) -> quickwit_search::Result<LeafListTermsResponse> {
        self.inner.leaf_list_terms(request).await
    }
    async fn scroll(&self, scroll_request: ScrollRequest) -> quickwit_search::Result<SearchResponse> {
        self.inner.scroll(scroll_request).await
    }
    async fn put_kv(&self, put_kv: PutKvRequest) {
        self.inner.put_kv(put_kv).await
    }
    async fn get_kv(&self, get_kv: GetKvRequest) -> Option<Vec<u8>> {
        self.inner.get_kv(get_kv).await
    }
    async fn report_splits(&self, report_splits: ReportSplitsRequest) -> ReportSplitsResponse {
        self.inner.report_splits(report_splits).await
    }
    async fn root_list_fields(&self, list_fields: ListFieldsRequest) -> quickwit_search::Result<ListFieldsResponse> {
        self.inner.root_list_fields(list_fields).await
    }
    async fn leaf_list_fields(&self, list_fields: LeafListFieldsRequest) -> quickwit_search::Result<ListFieldsResponse> {
        self.inner.leaf_list_fields(list_fields).await
    }
    async fn search_plan(&self, request: SearchRequest) -> quickwit_search::Result<SearchPlanResponse> {
        self.inner.search_plan(request).await
    }
}
