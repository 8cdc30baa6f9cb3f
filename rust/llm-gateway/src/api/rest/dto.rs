//! REST DTOs of `POST /llm-gateway/v1/tokenize` and `POST /llm-gateway/v1/count-tokens`.

use schemars::JsonSchema;
use serde::{Deserialize, Serialize};

#[derive(Debug, Deserialize, JsonSchema)]
#[serde(deny_unknown_fields)]
pub struct TokenizeRequest {
    /// canonical model id (`{provider_slug}::{provider_model_id}`) or vocabulary name
    pub model: String,
    /// texts to tokenize (one entry per prompt)
    pub texts: Vec<String>,
    /// return the ids as well as the counts
    #[serde(default)]
    pub return_ids: bool,
}

#[derive(Debug, Serialize, JsonSchema)]
pub struct TokenizeResponse {
    /// token count of every text (`Usage.input_tokens` is their sum)
    pub counts: Vec<u32>,
    pub input_tokens: u64,
    /// token ids of every text, when asked for
    #[serde(skip_serializing_if = "Option::is_none")]
    pub ids: Option<Vec<Vec<u32>>>,
}

/// How the provider frames the messages (`llm_gateway_sdk::ChatTemplate`); absent: content only.
#[derive(Debug, Deserialize, JsonSchema)]
#[serde(tag = "kind", rename_all = "snake_case", deny_unknown_fields)]
pub enum ChatTemplateDto {
    Overhead { tokens_per_message: u32, tokens_per_name: u32, reply_priming: u32 },
    Rendered { bos: String, message_prefix: String, message_suffix: String, generation_prompt: String, special_tokens: Vec<String> },
}

#[derive(Debug, Deserialize, JsonSchema)]
#[serde(deny_unknown_fields)]
pub struct CountTokensRequest {
    /// canonical model id or vocabulary name
    pub model: String,
    /// chat messages (`llm-gateway-sdk/schemas/core/message.v1.schema.json`): only their `text` content parts are counted
    pub messages: Vec<serde_json::Value>,
    /// the provider's framing; absent: the sum over the text parts
    pub template: Option<ChatTemplateDto>,
}

/// `gts.x.llmgw.core.usage.v1~` restricted to what a pre-call estimate knows
#[derive(Debug, Serialize, JsonSchema)]
pub struct CountTokensResponse {
    pub input_tokens: u64,
}
