// vocab.h -- host-side vocabulary loader and table builder (plain C++, no CUDA).
//
// This is the "model-registry vocab loader" row of SURVEY.md section 8(a5): the reference has
// no such component (modules/model-registry/docs/PRD.md:196-209 lists no tokenizer field),
// so the formats follow the stand-in oracle: ".tiktoken" rank files (tiktoken/load.py:160-172)
// and Mistral Tekken JSON (mistral_common/data/tekken_*.json).
#pragma once
#include <cstddef>
#include <cstdint>
#include <string>
#include <vector>

#include "tables.h"

namespace cfbpe {

// Parse a rank file into tokens[rank] = bytes.  Returns 0 or a negative CFBPE_* code; err gets a message.
int parse_tiktoken(const uint8_t* file, size_t len, uint32_t max_ranks, std::vector<std::string>& tokens,
                   std::string& err);
int parse_tekken_json(const uint8_t* file, size_t len, uint32_t max_ranks, std::vector<std::string>& tokens,
                      std::string& err);

// Build the packed table blob (TablesHeader + sections) for tokens[rank].
int build_tables(const std::vector<std::string>& tokens, uint32_t pattern_id, std::vector<uint8_t>& blob,
                 std::string& err);

// Structural validation of an imported blob (bounds, magic, capacities); 0 or CFBPE_EINVAL.
int validate_tables(const uint8_t* blob, uint64_t size, std::string& err);

}  // namespace cfbpe
