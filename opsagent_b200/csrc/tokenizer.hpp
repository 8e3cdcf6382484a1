// tokenizer.hpp — synthetic byte-level tokenizer and chat templates.
//
// No tokenizer files exist offline, so text maps to ids 0..255 (one per UTF-8 byte) and
// the chat-format control tokens keep their real ids inside the real vocabulary size.  Decode is
// many-to-one: a non-control id t renders as byte (t & 0xFF).  With config "tokenizer": "<tokenizer.json>" text goes through the
// checkpoint's own byte-level BPE instead (bpe.hpp) and the control ids come from its added_tokens.  The message order/roles templated here are
// the ones the ReAct loop sends (reference pkg/assistants/simple.go:358,496-501; seeds at
// pkg/handlers/execute.go:190-199).  oracle/oracle.py restates this file for the tests.
#pragma once
#include <cstdint>
#include <string>
#include <vector>

#include "bpe.hpp"
#include "config.hpp"

namespace oa {

struct ChatMessage { std::string role, content; };

class Tokenizer {
public:
    explicit Tokenizer(const ModelConfig& c, const std::string& tokenizer_json = "") : llama3_(c.chat_template == "llama3"), vocab_(c.vocab) {
        if (!tokenizer_json.empty()) {
            bpe_ = BpeTokenizer::load(tokenizer_json);
            if (bpe_->vocab_size() > vocab_) throw std::runtime_error("tokenizer has " + std::to_string(bpe_->vocab_size()) + " ids but the model's vocabulary is " + std::to_string(vocab_));
            auto need = [&](const char* name) { const int id = bpe_->special_id(name); if (id < 0) throw std::runtime_error(std::string("tokenizer lacks the control token ") + name); return id; };
            if (llama3_) { bot_ = need("<|begin_of_text|>"); eot_text_ = need("<|end_of_text|>"); sh_ = need("<|start_header_id|>"); eh_ = need("<|end_header_id|>"); eot_ = need("<|eot_id|>"); eos_ = {eot_, eot_text_}; }
            else { eot_text_ = need("<|endoftext|>"); im_start_ = need("<|im_start|>"); im_end_ = need("<|im_end|>"); eos_ = {im_end_, eot_text_}; }
            return;
        }
        if (llama3_) {
            int ids[5] = {128000, 128001, 128006, 128007, 128009};   // bot, eot_text, start_hdr, end_hdr, eot_id
            if (ids[4] >= vocab_) for (int i = 0; i < 5; ++i) ids[i] = vocab_ - 5 + i;   // tiny test vocabularies
            bot_ = ids[0]; eot_text_ = ids[1]; sh_ = ids[2]; eh_ = ids[3]; eot_ = ids[4];
            eos_ = {eot_, eot_text_};
        } else {
            int ids[3] = {151643, 151644, 151645};                    // endoftext, im_start, im_end
            if (ids[2] >= vocab_) for (int i = 0; i < 3; ++i) ids[i] = vocab_ - 3 + i;
            eot_text_ = ids[0]; im_start_ = ids[1]; im_end_ = ids[2];
            eos_ = {im_end_, eot_text_};
        }
    }
    bool byte_level() const { return !bpe_; }       // ids 0..255 are the bytes
    // bytes of every token a grammar-constrained completion may emit (token_mask.hpp): the checkpoint's text tokens, or ids 0..255 = one byte each
    std::vector<std::string> text_token_bytes() const {
        if (bpe_) return bpe_->text_token_bytes();
        std::vector<std::string> v(256);
        for (int b = 0; b < 256; ++b) v[b] = std::string(1, (char)b);
        return v;
    }
    int vocab() const { return vocab_; }
    void bytes(const std::string& s, std::vector<int32_t>& out) const {
        if (bpe_) { bpe_->encode(s, out); return; }
        for (unsigned char ch : s) out.push_back((int32_t)ch);
    }
    std::vector<int32_t> apply_chat_template(const std::vector<ChatMessage>& msgs) const {
        std::vector<int32_t> ids;
        size_t n = 16; for (auto& m : msgs) n += m.role.size() + m.content.size() + 8;
        ids.reserve(n);
        if (llama3_) {
            ids.push_back(bot_);
            for (auto& m : msgs) { ids.push_back(sh_); bytes(m.role, ids); ids.push_back(eh_); bytes("\n\n", ids); bytes(m.content, ids); ids.push_back(eot_); }
            ids.push_back(sh_); bytes("assistant", ids); ids.push_back(eh_); bytes("\n\n", ids);
        } else {
            for (auto& m : msgs) { ids.push_back(im_start_); bytes(m.role + "\n", ids); bytes(m.content, ids); ids.push_back(im_end_); bytes("\n", ids); }
            ids.push_back(im_start_); bytes("assistant\n", ids);
        }
        return ids;
    }
    bool is_eos(int32_t id) const { for (int e : eos_) if (e == id) return true; return false; }
    std::string detokenize(const std::vector<int32_t>& ids) const {
        if (bpe_) return bpe_->decode(ids.data(), ids.size());
        std::string s; s.reserve(ids.size());
        for (int32_t t : ids) s.push_back((char)(t & 0xFF));
        return s;
    }
    const std::vector<int32_t>& eos_ids() const { return eos_; }
private:
    bool llama3_; int vocab_;
    std::shared_ptr<BpeTokenizer> bpe_;
    int bot_ = 0, eot_text_ = 0, sh_ = 0, eh_ = 0, eot_ = 0, im_start_ = 0, im_end_ = 0;
    std::vector<int32_t> eos_;
};

}  // namespace oa
