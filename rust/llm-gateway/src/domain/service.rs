//! `llm-gateway::tokenizer` + `llm-gateway::usage::count_tokens`: the gateway-side domain service.
//! Plugin resolution is lazy and single-flight (`libs/modkit/src/plugins/mod.rs:44-78`): vendor match, lowest priority wins
//! (`:136-191`).  Prompt text is never logged (`docs/DESIGN.md:120-124`): only sizes and counts.

use std::sync::Arc;

use async_trait::async_trait;
use bytes::Bytes;
use llm_gateway_sdk::{
    ChatTemplate, CountTokensRequest, EncodeBatchRequest, SpecialTokens, TokenizerClient, TokenizerError, TokenizerPluginClient, TokenizerPluginSpecV1, Usage,
    VocabRef,
};
use modkit::client_hub::{ClientHub, ClientScope};
use modkit::plugins::{choose_plugin_instance, GtsPluginSelector};
use modkit_security::SecurityContext;
use serde_json::Value;
use types_registry_sdk::{ListQuery, TypesRegistryClient};

pub struct TokenizerService {
    hub: Arc<ClientHub>,
    vendor: String,
    selector: GtsPluginSelector,
}

/// list[str] -> (packed bytes, n + 1 offsets): the packed multi-tenant prompt buffer the device reads
pub fn pack_texts<S: AsRef<str>>(texts: &[S]) -> (Bytes, Vec<u64>) {
    let mut bytes = Vec::with_capacity(texts.iter().map(|t| t.as_ref().len()).sum());
    let mut offsets = Vec::with_capacity(texts.len() + 1);
    offsets.push(0u64);
    for t in texts {
        bytes.extend_from_slice(t.as_ref().as_bytes());
        offsets.push(bytes.len() as u64);
    }
    (Bytes::from(bytes), offsets)
}

impl TokenizerService {
    pub fn new(hub: Arc<ClientHub>, vendor: String) -> Self {
        Self { hub, vendor, selector: GtsPluginSelector::new() }
    }

    async fn plugin(&self) -> Result<Arc<dyn TokenizerPluginClient>, TokenizerError> {
        let instance_id = self
            .selector
            .get_or_init(|| async {
                let registry = self.hub.get::<dyn TypesRegistryClient>().map_err(|e| TokenizerError::Internal(e.to_string()))?;
                let plugin_type_id = TokenizerPluginSpecV1::gts_schema_id().clone();
                let instances = registry
                    .list(ListQuery::new().with_pattern(format!("{plugin_type_id}*")).with_is_type(false))
                    .await
                    .map_err(|e| TokenizerError::Internal(e.to_string()))?;
                choose_plugin_instance::<TokenizerPluginSpecV1>(&self.vendor, instances.iter().map(|e| (e.gts_id.as_str(), &e.content)))
                    .map_err(|_| TokenizerError::NoPluginAvailable)
            })
            .await?;
        self.hub
            .try_get_scoped::<dyn TokenizerPluginClient>(&ClientScope::gts_id(instance_id.as_ref()))
            .ok_or_else(|| TokenizerError::ServiceUnavailable(format!("tokenizer plugin {instance_id} is not registered yet")))
    }

    /// the `TextContent.text` parts of a request's messages (`schemas/core/message.v1.schema.json`)
    fn text_parts(messages: &[Value]) -> Vec<String> {
        messages
            .iter()
            .filter_map(|m| m.get("content")?.as_array())
            .flatten()
            .filter(|p| p.get("type").and_then(Value::as_str) == Some("text"))
            .filter_map(|p| p.get("text")?.as_str().map(str::to_owned))
            .collect()
    }
}

#[async_trait]
impl TokenizerClient for TokenizerService {
    async fn encode(&self, ctx: &SecurityContext, model: &str, texts: &[String]) -> Result<Vec<Vec<u32>>, TokenizerError> {
        let (bytes, offsets) = pack_texts(texts);
        let r = self.plugin().await?.encode_batch(ctx, EncodeBatchRequest { vocab: VocabRef(model.to_owned()), bytes, offsets, vocabs_per_prompt: None, vocab_index: None }).await?;
        Ok((0..texts.len()).map(|i| r.ids[r.offsets[i] as usize..r.offsets[i + 1] as usize].to_vec()).collect())
    }

    async fn encode_with_special(&self, ctx: &SecurityContext, model: &str, texts: &[String], special: &SpecialTokens)
        -> Result<Vec<Vec<u32>>, TokenizerError> {
        // cut every text at the allowed special tokens (leftmost, longest first), send all stretches of all texts through ONE
        // plugin batch, put the special ids back; a text that spells a token that is not allowed is refused
        let mut allowed: Vec<&String> = special.allowed.iter().collect();
        allowed.sort_by_key(|t| std::cmp::Reverse(t.len()));
        if let Some(unknown) = allowed.iter().find(|t| !special.ids.contains_key(**t)) {
            return Err(TokenizerError::InvalidInput(format!("allowed special token without an id: {unknown}")));
        }
        if special.disallow_all_others {
            for t in texts {
                if let Some(bad) = special.ids.keys().find(|k| !special.allowed.contains(*k) && t.contains(k.as_str())) {
                    return Err(TokenizerError::InvalidInput(format!("the text holds the special token {bad:?}, which is not allowed here")));
                }
            }
        }
        enum Step { Stretch(usize), Special(u32) }
        let (mut plan, mut stretches): (Vec<Vec<Step>>, Vec<String>) = (Vec::new(), Vec::new());
        for t in texts {
            let (mut steps, mut pos) = (Vec::new(), 0usize);
            while pos < t.len() {
                let next = allowed.iter().filter_map(|tok| t[pos..].find(tok.as_str()).map(|i| (pos + i, *tok))).min_by_key(|(i, tok)| (*i, std::cmp::Reverse(tok.len())));
                match next {
                    Some((i, tok)) => {
                        if i > pos { steps.push(Step::Stretch(stretches.len())); stretches.push(t[pos..i].to_owned()); }
                        steps.push(Step::Special(special.ids[tok]));
                        pos = i + tok.len();
                    }
                    None => { steps.push(Step::Stretch(stretches.len())); stretches.push(t[pos..].to_owned()); pos = t.len(); }
                }
            }
            plan.push(steps);
        }
        let enc = if stretches.is_empty() { Vec::new() } else { self.encode(ctx, model, &stretches).await? };
        Ok(plan.into_iter().map(|steps| steps.into_iter().flat_map(|s| match s { Step::Stretch(i) => enc[i].clone(), Step::Special(id) => vec![id] }).collect()).collect())
    }

    async fn count_tokens(&self, ctx: &SecurityContext, model: &str, messages: &[Value]) -> Result<Usage, TokenizerError> {
        let texts = Self::text_parts(messages);
        if texts.is_empty() {
            return Ok(Usage::default());
        }
        let (bytes, offsets) = pack_texts(&texts);
        let counts = self.plugin().await?.count_tokens(ctx, CountTokensRequest { vocab: VocabRef(model.to_owned()), bytes, offsets, vocabs_per_prompt: None, vocab_index: None }).await?;
        Ok(Usage { input_tokens: counts.iter().map(|c| u64::from(*c)).sum(), output_tokens: 0 })
    }

    async fn count_chat_tokens(&self, ctx: &SecurityContext, model: &str, messages: &[Value], template: &ChatTemplate) -> Result<Usage, TokenizerError> {
        // every stretch of ordinary text of the whole request goes to the device in ONE batch; control tokens count one each
        let content = |m: &Value| -> String {
            m.get("content").and_then(Value::as_array).map(|parts| {
                parts.iter().filter(|p| p.get("type").and_then(Value::as_str) == Some("text")).filter_map(|p| p.get("text").and_then(Value::as_str)).collect::<String>()
            }).unwrap_or_default()
        };
        let role = |m: &Value| m.get("role").and_then(Value::as_str).unwrap_or("").to_owned();
        let (mut fixed, mut texts): (u64, Vec<String>) = (0, Vec::new());
        match template {
            ChatTemplate::Overhead { tokens_per_message, tokens_per_name, reply_priming } => {
                for m in messages {
                    fixed += u64::from(*tokens_per_message);
                    texts.push(role(m));
                    if let Some(name) = m.get("name").and_then(Value::as_str) {
                        fixed += u64::from(*tokens_per_name);
                        texts.push(name.to_owned());
                    }
                    texts.extend(Self::text_parts(std::slice::from_ref(m)));
                }
                fixed += u64::from(*reply_priming);
            }
            ChatTemplate::Rendered { bos, message_prefix, message_suffix, generation_prompt, special_tokens } => {
                // framing text and content between two control tokens form one stretch (the pre-tokenizer may join them)
                let mut run = vec![String::new()];
                let feed = |framing: &str, fixed: &mut u64, run: &mut Vec<String>| {
                    let mut rest = framing;
                    loop {
                        let next = special_tokens.iter().filter_map(|t| rest.find(t.as_str()).map(|i| (i, t.len()))).min_by_key(|(i, l)| (*i, std::cmp::Reverse(*l)));
                        match next {
                            Some((i, l)) => { run.last_mut().expect("never empty").push_str(&rest[..i]); run.push(String::new()); *fixed += 1; rest = &rest[i + l..]; }
                            None => { run.last_mut().expect("never empty").push_str(rest); break; }
                        }
                    }
                };
                feed(bos, &mut fixed, &mut run);
                for m in messages {
                    feed(&message_prefix.replace("{role}", &role(m)), &mut fixed, &mut run);
                    run.last_mut().expect("never empty").push_str(&content(m));
                    feed(message_suffix, &mut fixed, &mut run);
                }
                feed(generation_prompt, &mut fixed, &mut run);
                texts = run;
            }
        }
        texts.retain(|t| !t.is_empty());
        if texts.is_empty() {
            return Ok(Usage { input_tokens: fixed, output_tokens: 0 });
        }
        let (bytes, offsets) = pack_texts(&texts);
        let counts = self.plugin().await?.count_tokens(ctx, CountTokensRequest { vocab: VocabRef(model.to_owned()), bytes, offsets, vocabs_per_prompt: None, vocab_index: None }).await?;
        Ok(Usage { input_tokens: fixed + counts.iter().map(|c| u64::from(*c)).sum::<u64>(), output_tokens: 0 })
    }

    async fn check_budget(&self, ctx: &SecurityContext, model: &str, messages: &[Value], remaining_tokens: u64) -> Result<bool, TokenizerError> {
        Ok(self.count_tokens(ctx, model, messages).await?.input_tokens <= remaining_tokens)
    }
}
