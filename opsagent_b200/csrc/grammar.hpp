// grammar.hpp — byte-level grammar that forces a completion to parse as tools.ToolPrompt
// (reference pkg/tools/tool.go:29-38: {question, thought, action{name,input}, observation, final_answer}).
//
// Why: AssistantWithConfig json.Unmarshals every reply into ToolPrompt and gives up on the first reply that does not
// parse (reference pkg/assistants/simple.go:366-382); a random-init (or merely sloppy) model then ends the ReAct loop
// after one step and triggers the extra "Summarize…" Chat (simple.go:544-566).  With this mask the engine can only emit
// schema-valid JSON, tool names are restricted to the registry (pkg/tools/tool.go:20-26) and `final_answer` is either
// forced empty (tool-call step) or forced to be >= 10 bytes (final step; shorter values are treated as template
// placeholders by isTemplateValue, simple.go:640-654).  oracle/oracle.py restates the same automaton for the tests.
#pragma once
#include <cstdint>
#include <string>
#include <vector>

namespace oa {

// GRAMMAR_FUNCTION: OpenAI function calling for the swarm-go flows (reference pkg/workflows/swarm.go:14-78: kubectl{command},
// trivy{image}, python{code}) — the completion is {"name":"<one of the offered functions>","arguments":{"<its parameter>":"…"}}.
// GRAMMAR_TEXT: a bounded line of printable text (the flows' final answer) terminated by '\n'.
enum GrammarKind : int { GRAMMAR_NONE = 0, GRAMMAR_TOOLCALL = 1, GRAMMAR_FINAL = 2, GRAMMAR_FUNCTION = 3, GRAMMAR_TEXT = 4 };

class ToolPromptGrammar {
public:
    struct Seg { int type; std::string lit; int min_len, max_len; };   // type 0 literal, 1 string, 2 tool-name enum
    // `functions` (GRAMMAR_FUNCTION): "name:param,name:param,…" — what the request's `tools` array offers
    explicit ToolPromptGrammar(int kind = GRAMMAR_NONE, const std::string& functions = "") : kind_(kind), opts_(tools()) {
        auto L = [&](const char* s) { segs_.push_back(Seg{0, s, 0, 0}); };
        auto S = [&](int lo, int hi) { segs_.push_back(Seg{1, "", lo, hi}); };
        if (kind == GRAMMAR_TOOLCALL) {
            L("{\"question\":\""); S(1, 64); L("\",\"thought\":\""); S(1, 96); L("\",\"action\":{\"name\":\"");
            segs_.push_back(Seg{2, "", 0, 0}); L("\",\"input\":\""); S(1, 96); L("\"},\"observation\":\"\",\"final_answer\":\"\"}");
        } else if (kind == GRAMMAR_FINAL) {
            L("{\"question\":\""); S(1, 64); L("\",\"thought\":\""); S(1, 96);
            L("\",\"action\":{\"name\":\"\",\"input\":\"\"},\"observation\":\"\",\"final_answer\":\""); S(10, 160); L("\"}");
        } else if (kind == GRAMMAR_FUNCTION) {
            opts_.clear();
            size_t b = 0;
            while (b < functions.size() && opts_.size() < 32) {
                size_t e = functions.find(',', b); if (e == std::string::npos) e = functions.size();
                const std::string item = functions.substr(b, e - b); const size_t c = item.find(':');
                if (c != std::string::npos && c > 0 && c + 1 < item.size()) opts_.push_back(item.substr(0, c) + "\",\"arguments\":{\"" + item.substr(c + 1) + "\":\"");
                b = e + 1;
            }
            if (opts_.empty()) { kind_ = GRAMMAR_NONE; }
            else { L("{\"name\":\""); segs_.push_back(Seg{2, "", 0, 0}); S(1, 96); L("\"}}"); }
        } else if (kind == GRAMMAR_TEXT) {
            close_ = '\n'; S(10, 200); L("\n");
        }
        reset();
    }
    static const std::vector<std::string>& tools() { static const std::vector<std::string> t = {"kubectl", "python", "trivy", "jq", "search"}; return t; }
    void reset() { seg_ = 0; off_ = 0; cand_ = all_opts(); }
    bool active() const { return kind_ != GRAMMAR_NONE; }
    bool done() const { return !active() || seg_ >= (int)segs_.size(); }
    static bool string_byte(int b) { return b >= 0x20 && b <= 0x7E && b != '"' && b != '\\'; }

    // 256-bit set of bytes allowed next (8 x uint32, bit b of word b/32)
    void allowed(uint32_t out[8]) const {
        for (int i = 0; i < 8; ++i) out[i] = 0;
        if (done()) return;
        auto set = [&](int b) { out[b >> 5] |= 1u << (b & 31); };
        const Seg& s = segs_[seg_];
        if (s.type == 0) { set((unsigned char)s.lit[off_]); }
        else if (s.type == 1) {
            if (off_ < s.max_len) for (int b = 0x20; b <= 0x7E; ++b) if (string_byte(b)) set(b);
            if (off_ >= s.min_len) set((unsigned char)close_);   // the closing quote (or line end) opens the following literal
        } else {
            for (size_t t = 0; t < opts_.size(); ++t) if (cand_ & (1u << t)) set((unsigned char)opts_[t][off_]);
        }
    }
    // consume one emitted byte (must be in the allowed set); returns false if it was not
    bool advance(int b) {
        if (done()) return false;
        const Seg& s = segs_[seg_];
        if (s.type == 0) {
            if ((unsigned char)s.lit[off_] != b) return false;
            if (++off_ == (int)s.lit.size()) next();
        } else if (s.type == 1) {
            if (b == (unsigned char)close_ && off_ >= s.min_len) { next(); off_ = 1; if (off_ == (int)segs_[seg_].lit.size()) next(); }
            else if (string_byte(b) && off_ < s.max_len) ++off_;
            else return false;
        } else {
            uint32_t keep = 0;
            for (size_t t = 0; t < opts_.size(); ++t) if ((cand_ & (1u << t)) && (unsigned char)opts_[t][off_] == b) keep |= 1u << t;
            if (!keep) return false;
            cand_ = keep; ++off_;
            for (size_t t = 0; t < opts_.size(); ++t) if ((cand_ & (1u << t)) && off_ == (int)opts_[t].size()) { next(); break; }
        }
        return true;
    }
    int kind() const { return kind_; }

private:
    uint32_t all_opts() const { return opts_.size() >= 32 ? 0xffffffffu : ((1u << opts_.size()) - 1); }
    void next() { ++seg_; off_ = 0; cand_ = all_opts(); }
    int kind_; std::vector<std::string> opts_; char close_ = '"'; std::vector<Seg> segs_; int seg_ = 0, off_ = 0; uint32_t cand_ = 0;
};

}  // namespace oa
