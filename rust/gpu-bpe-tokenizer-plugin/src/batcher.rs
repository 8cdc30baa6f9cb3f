//! Micro-batcher for `count_tokens`: a `stateful` lifecycle task in ModKit terms
//! (`docs/modkit_unified_system/08_lifecycle_stateful_tasks.md:14-58`).  Request handlers enqueue `(bytes, offsets, vocab ids)`
//! with a `oneshot` for the answer; ONE worker packs what is waiting -- until `batch_bytes` / `max_prompts` are reached or
//! `wait` after the first item -- into one `cfbpe_count_batch` call inside `spawn_blocking` and hands every caller exactly its
//! own counts.  Requests of different tenants share batches, so a failing batch is retried request by request: a bad request
//! (malformed UTF-8) fails its own caller only.  Python mirror with tests: `cfbpe/plugin.py:CountTokensMicroBatcher`.

use std::sync::Arc;
use std::time::Duration;

use bytes::Bytes;
use cfbpe_sys::Ctx;
use llm_gateway_sdk::TokenizerError;
use tokio::sync::{mpsc, oneshot};

use crate::service::map_native_error;

struct Item {
    bytes: Bytes,
    offsets: Vec<u64>,
    vocab_ids: Option<Vec<u8>>,
    reply: oneshot::Sender<Result<Vec<u32>, TokenizerError>>,
}

pub struct CountBatcher {
    tx: mpsc::Sender<Item>,
    batch_bytes: u64,
}

impl CountBatcher {
    pub fn start(native: Arc<Ctx>, batch_bytes: u64, max_prompts: u32, wait_us: u64) -> Self {
        let (tx, rx) = mpsc::channel::<Item>(65_536);
        tokio::spawn(run(native, rx, batch_bytes, max_prompts as usize, Duration::from_micros(wait_us)));
        Self { tx, batch_bytes }
    }

    /// requests of at least this size skip the queue (they are a batch of their own)
    pub fn direct_threshold(&self) -> u64 {
        self.batch_bytes / 4
    }

    pub async fn count(&self, bytes: Bytes, offsets: Vec<u64>, vocab_ids: Option<Vec<u8>>) -> Result<Vec<u32>, TokenizerError> {
        let (reply, answer) = oneshot::channel();
        self.tx
            .send(Item { bytes, offsets, vocab_ids, reply })
            .await
            .map_err(|_| TokenizerError::ServiceUnavailable("the count_tokens batcher stopped".to_owned()))?;
        answer.await.map_err(|_| TokenizerError::ServiceUnavailable("the count_tokens batcher stopped".to_owned()))?
    }
}

fn pack(items: &[Item]) -> (Vec<u8>, Vec<u64>, Vec<u8>) {
    let mut bytes = Vec::new();
    let mut offsets = vec![0u64];
    let mut vids = Vec::new();
    for it in items {
        let n = it.offsets.len() - 1;
        let base = bytes.len() as u64;
        bytes.extend_from_slice(&it.bytes[..it.offsets[n] as usize]);
        offsets.extend(it.offsets[1..].iter().map(|o| base + o));
        match &it.vocab_ids {
            Some(v) => vids.extend_from_slice(&v[..n]),
            None => vids.extend(std::iter::repeat(0u8).take(n)),
        }
    }
    (bytes, offsets, vids)
}

async fn flush(native: &Arc<Ctx>, items: Vec<Item>) {
    let (bytes, offsets, vids) = pack(&items);
    let nat = native.clone();
    let whole = tokio::task::spawn_blocking(move || nat.count_batch(&bytes, &offsets, Some(&vids))).await;
    match whole {
        Ok(Ok(counts)) => {
            let mut k = 0;
            for it in items {
                let n = it.offsets.len() - 1;
                let _ = it.reply.send(Ok(counts[k..k + n].to_vec()));
                k += n;
            }
        }
        _ => {
            // one request of the batch is bad (or the device failed): every caller gets the outcome of ITS request
            for it in items {
                let nat = native.clone();
                let (b, o, v) = (it.bytes.clone(), it.offsets.clone(), it.vocab_ids.clone());
                let r = tokio::task::spawn_blocking(move || nat.count_batch(&b, &o, v.as_deref())).await;
                let _ = it.reply.send(match r {
                    Ok(Ok(c)) => Ok(c),
                    Ok(Err(e)) => Err(map_native_error(e)),
                    Err(e) => Err(TokenizerError::Internal(e.to_string())),
                });
            }
        }
    }
}

async fn run(native: Arc<Ctx>, mut rx: mpsc::Receiver<Item>, batch_bytes: u64, max_prompts: usize, wait: Duration) {
    let mut carry: Option<Item> = None;
    loop {
        let first = match carry.take() {
            Some(it) => it,
            None => match rx.recv().await {
                Some(it) => it,
                None => return,
            },
        };
        let mut size = first.bytes.len() as u64;
        let mut prompts = first.offsets.len() - 1;
        let mut batch = vec![first];
        let deadline = tokio::time::Instant::now() + wait;
        while size < batch_bytes && prompts < max_prompts {
            match tokio::time::timeout_at(deadline, rx.recv()).await {
                Ok(Some(it)) => {
                    let (s, p) = (it.bytes.len() as u64, it.offsets.len() - 1);
                    if size + s > batch_bytes || prompts + p > max_prompts {
                        carry = Some(it); // does not fit: it opens the next batch
                        break;
                    }
                    size += s;
                    prompts += p;
                    batch.push(it);
                }
                _ => break,
            }
        }
        flush(&native, batch).await;
    }
}
