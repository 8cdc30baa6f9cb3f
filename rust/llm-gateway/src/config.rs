//! Module configuration.

use serde::Deserialize;

#[derive(Debug, Clone, Deserialize)]
#[serde(default, deny_unknown_fields)]
pub struct LlmGatewayConfig {
    /// Vendor whose tokenizer plugin the gateway selects (lowest priority wins among its instances).
    pub tokenizer_vendor: String,
}

impl Default for LlmGatewayConfig {
    fn default() -> Self {
        Self { tokenizer_vendor: "cyberfabric".to_owned() }
    }
}
