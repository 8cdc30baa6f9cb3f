//! Error types of the tokenizer API (shape of `tenant-resolver-sdk/src/error.rs:7-34`).

use thiserror::Error;

/// Errors of the tokenizer API.  REST mapping (RFC 9457 `Problem`, `libs/modkit-errors/src/problem.rs:41-53`):
/// `InvalidInput` 400, `VocabNotFound` 404, `NoPluginAvailable` / `ServiceUnavailable` 503, `Internal` 500.
#[derive(Debug, Error)]
pub enum TokenizerError {
    /// Bad offsets, malformed UTF-8, a batch beyond the plugin's limits, a disallowed special token in the text
    /// (`CFBPE_EINVAL`, `CFBPE_EILSEQ`, `CFBPE_ENOSPC`).
    #[error("invalid input: {0}")]
    InvalidInput(String),

    /// The model or vocabulary is not known to / not loaded on the plugin (`CFBPE_ENOENT`).
    #[error("vocabulary not found: {vocab}")]
    VocabNotFound {
        /// what the request named
        vocab: String,
    },

    /// No plugin is available to handle the request.
    #[error("no plugin available")]
    NoPluginAvailable,

    /// The plugin is not available yet, or has no device (`CFBPE_ENODEV`, `CFBPE_ENOMEM`): there is no CPU fallback.
    #[error("service unavailable: {0}")]
    ServiceUnavailable(String),

    /// An internal error occurred (`CFBPE_EIO`: CUDA / NCCL failure).
    #[error("internal error: {0}")]
    Internal(String),
}
