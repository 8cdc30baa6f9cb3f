//! LLM Gateway — the tokenizer slice: `llm-gateway::tokenizer`, `llm-gateway::usage::count_tokens`, and the REST endpoint
//! `POST /llm-gateway/v1/tokenize`.  The rest of the module (providers, chat completions, budgets) is specified in
//! `modules/llm-gateway/docs/` and not part of this slice.  NOT COMPILED where this file lives (no Rust toolchain);
//! Python mirror: `cyberfabric-core_b200/cfbpe/plugin.py:LlmGatewayTokenizerService`.

pub mod api;
pub mod config;
pub mod domain;
pub mod module;

pub use module::LlmGateway;
