//! Public API trait for consumers of the gateway's tokenizer (unscoped in `ClientHub`).

use async_trait::async_trait;
use modkit_security::SecurityContext;
use serde_json::Value;

use crate::error::TokenizerError;
use crate::models::{ChatTemplate, SpecialTokens, Usage};

#[async_trait]
pub trait TokenizerClient: Send + Sync {
    /// `encode_ordinary` of every text under the vocabulary of `model` (canonical id or vocabulary name).
    async fn encode(&self, ctx: &SecurityContext, model: &str, texts: &[String]) -> Result<Vec<Vec<u32>>, TokenizerError>;

    /// tiktoken's `encode(text, allowed_special = …, disallowed_special = …)`.
    async fn encode_with_special(&self, ctx: &SecurityContext, model: &str, texts: &[String], special: &SpecialTokens)
        -> Result<Vec<Vec<u32>>, TokenizerError>;

    /// `Usage.input_tokens` of one chat request: the sum over its `TextContent.text` parts
    /// (`schemas/core/message.v1.schema.json`, `schemas/content/text_content.v1.schema.json`); `messages` are the request's
    /// message objects as JSON.
    async fn count_tokens(&self, ctx: &SecurityContext, model: &str, messages: &[Value]) -> Result<Usage, TokenizerError>;

    /// Pre-call estimate against a remaining budget (`docs/DESIGN.md:833-855`).
    /// `Usage.input_tokens` as the provider counts it: content AND the framing of `template`
    async fn count_chat_tokens(&self, ctx: &SecurityContext, model: &str, messages: &[Value], template: &ChatTemplate) -> Result<Usage, TokenizerError>;
    async fn check_budget(&self, ctx: &SecurityContext, model: &str, messages: &[Value], remaining_tokens: u64) -> Result<bool, TokenizerError>;
}
