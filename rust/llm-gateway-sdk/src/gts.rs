//! GTS schema of tokenizer plugin instances (shape of `tenant-resolver-sdk/src/gts.rs:40-47`).
//!
//! Instance id format: `gts.x.core.modkit.plugin.v1~<vendor>.<package>.tokenizer.plugin.v1~`; the gateway registers the schema,
//! a plugin registers its instance `{id, vendor, priority, properties}` and a client scoped by that id
//! (`docs/MODKIT_PLUGINS.md:152-164`).

use gts_macros::struct_to_gts_schema;
use modkit::gts::BaseModkitPluginV1;

#[struct_to_gts_schema(
    dir_path = "schemas",
    base = BaseModkitPluginV1,
    schema_id = "gts.x.core.modkit.plugin.v1~x.llmgw.tokenizer.plugin.v1~",
    description = "LLM Gateway tokenizer plugin specification",
    properties = ""
)]
pub struct TokenizerPluginSpecV1;
