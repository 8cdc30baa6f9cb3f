//! Module declaration: registers the plugin SCHEMA (plugins register their instances), the public client and the REST routes
//! (gateway side of `docs/MODKIT_PLUGINS.md:152-164`).

use std::sync::Arc;

use async_trait::async_trait;
use llm_gateway_sdk::{TokenizerClient, TokenizerPluginSpecV1};
use modkit::api::OpenApiRegistry;
use modkit::context::ModuleCtx;
use modkit::{Module, RestApiCapability};
use tracing::info;
use types_registry_sdk::{RegisterResult, TypesRegistryClient};

use crate::config::LlmGatewayConfig;
use crate::domain::service::TokenizerService;

#[modkit::module(
    name = "llm-gateway",
    deps = ["types-registry"],
    capabilities = [rest]
)]
#[derive(Default)]
pub struct LlmGateway {
    service: std::sync::OnceLock<Arc<TokenizerService>>,
}

#[async_trait]
impl Module for LlmGateway {
    async fn init(&self, ctx: &ModuleCtx) -> anyhow::Result<()> {
        let cfg: LlmGatewayConfig = ctx.config()?;
        let registry = ctx.client_hub().get::<dyn TypesRegistryClient>()?;
        // the gateway registers the plugin SCHEMA (same sequence as tenant-resolver/src/module.rs:55-70, including the
        // additionalProperties patch a derived GTS schema needs until gts-macros emits it)
        let mut schema: serde_json::Value = serde_json::from_str(&TokenizerPluginSpecV1::gts_schema_with_refs_as_string())?;
        if let Some(o) = schema.as_object_mut() {
            o.insert("additionalProperties".to_owned(), serde_json::Value::Bool(false));
        }
        RegisterResult::ensure_all_ok(&registry.register(vec![schema]).await?)?;
        let service = Arc::new(TokenizerService::new(ctx.client_hub(), cfg.tokenizer_vendor));
        self.service.set(service.clone()).map_err(|_| anyhow::anyhow!("llm-gateway already initialized"))?;
        let api: Arc<dyn TokenizerClient> = service;
        ctx.client_hub().register::<dyn TokenizerClient>(api);
        info!("llm-gateway tokenizer service ready (plugin resolution is lazy)");
        Ok(())
    }
}

impl RestApiCapability for LlmGateway {
    fn register_rest(&self, _ctx: &ModuleCtx, router: axum::Router, openapi: &dyn OpenApiRegistry) -> anyhow::Result<axum::Router> {
        let service = self.service.get().ok_or_else(|| anyhow::anyhow!("llm-gateway not initialized"))?.clone();
        Ok(crate::api::rest::routes::register_routes(router, openapi, service))
    }
}
