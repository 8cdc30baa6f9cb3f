//! REST DTOs of `POST /llm-gateway/v1/tokenize`.

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
