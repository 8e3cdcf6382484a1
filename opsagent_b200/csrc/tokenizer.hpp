// tokenizer.hpp — synthetic byte-level tokenizer and chat templates.
//
// No tokenizer files exist offline, so text maps to ids 0..255 (one per UTF-8 byte) and
// the chat-format control tokens keep their real ids inside the real vocabulary size.  Decode is
// many-to-one: a non-control id t renders as byte (t & 0xFF).  The message order/roles templated here are
// the ones the ReAct loop sends (reference pkg/assistants/simple.go:358,496-501; seeds at
// pkg/handlers/execute.go:190-199).  oracle/oracle.py restates this file for the tests.
#pragma once
#include <cstdint>
#include <string>
#include <vector>

#include "config.hpp"

namespace oa {

struct ChatMessage { std::string role, content; };

class Tokenizer {
public:
    explicit Tokenizer(const ModelConfig& c) : llama3_(c.chat_template == "llama3"), vocab_(c.vocab) {
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
    void bytes(const std::string& s, std::vector<int32_t>& out) const { for (unsigned char ch : s) out.push_back((int32_t)ch); }
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
        std::string s; s.reserve(ids.size());
        for (int32_t t : ids) s.push_back((char)(t & 0xFF));
        return s;
    }
    const std::vector<int32_t>& eos_ids() const { return eos_; }
private:
    bool llama3_; int vocab_;
    int bot_ = 0, eot_text_ = 0, sh_ = 0, eh_ = 0, eot_ = 0, im_start_ = 0, im_end_ = 0;
    std::vector<int32_t> eos_;
};

}  // namespace oa
