//! Route registration (OperationBuilder pattern of `modules/file-parser/src/api/rest/routes.rs:49-71`; versioned path: DE0801).

use std::sync::Arc;

use axum::{Extension, Router};
use modkit::api::{operation_builder::LicenseFeature, OpenApiRegistry, OperationBuilder};

use crate::api::rest::{dto, handlers};
use crate::domain::service::TokenizerService;

struct License;

impl AsRef<str> for License {
    fn as_ref(&self) -> &'static str {
        "gts.x.core.lic.feat.v1~x.core.global.base.v1"
    }
}

impl LicenseFeature for License {}

#[allow(clippy::needless_pass_by_value)] // Arc is intentionally passed by value for the Extension layer
pub fn register_routes(mut router: Router, openapi: &dyn OpenApiRegistry, service: Arc<TokenizerService>) -> Router {
    // POST /llm-gateway/v1/tokenize - token ids / counts of a list of texts under a model's vocabulary
    router = OperationBuilder::post("/llm-gateway/v1/tokenize")
        .operation_id("llm_gateway.tokenize")
        .summary("Tokenize texts with the vocabulary of a model")
        .tag("LLM Gateway")
        .authenticated()
        .require_license_features::<License>([])
        .json_request::<dto::TokenizeRequest>(openapi, "Texts and the model whose vocabulary applies")
        .allow_content_types(&["application/json"])
        .handler(handlers::tokenize)
        .json_response_with_schema::<dto::TokenizeResponse>(openapi, http::StatusCode::OK, "Token counts (and ids)")
        .standard_errors(openapi)
        .error_415(openapi)
        .register(router, openapi);
    // POST /llm-gateway/v1/count-tokens - Usage.input_tokens of a chat request before it is sent (budget / context-window checks)
    router = OperationBuilder::post("/llm-gateway/v1/count-tokens")
        .operation_id("llm_gateway.count_tokens")
        .summary("Count the input tokens of a chat request, with the provider's framing if a template is given")
        .tag("LLM Gateway")
        .authenticated()
        .require_license_features::<License>([])
        .json_request::<dto::CountTokensRequest>(openapi, "Messages, the model whose vocabulary applies, and optionally its chat template")
        .allow_content_types(&["application/json"])
        .handler(handlers::count_tokens)
        .json_response_with_schema::<dto::CountTokensResponse>(openapi, http::StatusCode::OK, "Usage.input_tokens")
        .standard_errors(openapi)
        .error_415(openapi)
        .register(router, openapi);
    router.layer(Extension(service))
}
