//! Handlers: translate DTOs, map `TokenizerError` to RFC 9457 `Problem` (`libs/modkit-errors/src/problem.rs:41-53`).

use std::sync::Arc;

use axum::{Extension, Json};
use llm_gateway_sdk::{ChatTemplate, TokenizerClient, TokenizerError};
use modkit_errors::Problem;
use modkit_security::SecurityContext;

use crate::api::rest::dto::{ChatTemplateDto, CountTokensRequest, CountTokensResponse, TokenizeRequest, TokenizeResponse};
use crate::domain::service::TokenizerService;

fn problem(e: TokenizerError) -> Problem {
    let (status, title) = match &e {
        TokenizerError::InvalidInput(_) => (http::StatusCode::BAD_REQUEST, "Invalid input"),
        TokenizerError::VocabNotFound { .. } => (http::StatusCode::NOT_FOUND, "Vocabulary not found"),
        TokenizerError::NoPluginAvailable | TokenizerError::ServiceUnavailable(_) => (http::StatusCode::SERVICE_UNAVAILABLE, "Tokenizer unavailable"),
        TokenizerError::Internal(_) => (http::StatusCode::INTERNAL_SERVER_ERROR, "Internal error"),
    };
    Problem::new(status, title, e.to_string())
}

pub async fn tokenize(
    Extension(ctx): Extension<SecurityContext>,
    Extension(service): Extension<Arc<TokenizerService>>,
    Json(req): Json<TokenizeRequest>,
) -> Result<Json<TokenizeResponse>, Problem> {
    // only sizes are logged, never the text (docs/DESIGN.md:120-124)
    tracing::debug!(model = %req.model, texts = req.texts.len(), bytes = req.texts.iter().map(String::len).sum::<usize>(), "tokenize");
    let ids = service.encode(&ctx, &req.model, &req.texts).await.map_err(problem)?;
    let counts: Vec<u32> = ids.iter().map(|v| v.len() as u32).collect();
    let input_tokens = counts.iter().map(|c| u64::from(*c)).sum();
    Ok(Json(TokenizeResponse { counts, input_tokens, ids: req.return_ids.then_some(ids) }))
}

pub async fn count_tokens(
    Extension(ctx): Extension<SecurityContext>,
    Extension(service): Extension<Arc<TokenizerService>>,
    Json(req): Json<CountTokensRequest>,
) -> Result<Json<CountTokensResponse>, Problem> {
    tracing::debug!(model = %req.model, messages = req.messages.len(), "count_tokens");
    let usage = match req.template {
        None => service.count_tokens(&ctx, &req.model, &req.messages).await,
        Some(t) => {
            let template = match t {
                ChatTemplateDto::Overhead { tokens_per_message, tokens_per_name, reply_priming } => ChatTemplate::Overhead { tokens_per_message, tokens_per_name, reply_priming },
                ChatTemplateDto::Rendered { bos, message_prefix, message_suffix, generation_prompt, special_tokens } => {
                    ChatTemplate::Rendered { bos, message_prefix, message_suffix, generation_prompt, special_tokens }
                }
            };
            service.count_chat_tokens(&ctx, &req.model, &req.messages, &template).await
        }
    }
    .map_err(problem)?;
    Ok(Json(CountTokensResponse { input_tokens: usage.input_tokens }))
}
