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
    explicit ToolPromptGrammar(int kind = GRAMMAR_NONE, const std::string& functions = "") : kind_(kind), sig_(kind == GRAMMAR_FUNCTION ? functions : ""), opts_(tools()) {
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
    // The automaton's whole state is three integers; the segment table is immutable after construction.  Token-level masks
    // (token_mask.hpp) walk thousands of tokens' bytes from one state, so the transition function works on a POD cursor.
    struct Cursor { int seg, off; uint32_t cand; };
    Cursor cursor() const { return Cursor{seg_, off_, cand_}; }
    void set_cursor(const Cursor& c) { seg_ = c.seg; off_ = c.off; cand_ = c.cand; }
    Cursor start() const { return Cursor{0, 0, all_opts()}; }
    bool done_at(const Cursor& c) const { return !active() || c.seg >= (int)segs_.size(); }
    // identifies (grammar, state) for the mask cache: two sequences in the same state of the same grammar share one token mask
    std::string state_key() const { return state_key(cursor()); }
    std::string state_key(const Cursor& c) const {
        std::string k = std::to_string(kind_) + "|" + sig_ + "|" + std::to_string(c.seg) + "|" + std::to_string(c.off) + "|" + std::to_string(c.cand);
        return k;
    }
    void reset() { set_cursor(start()); }
    bool active() const { return kind_ != GRAMMAR_NONE; }
    bool done() const { return done_at(cursor()); }
    static bool string_byte(int b) { return b >= 0x20 && b <= 0x7E && b != '"' && b != '\\'; }

    // 256-bit set of bytes allowed next (8 x uint32, bit b of word b/32)
    void allowed(uint32_t out[8]) const { allowed_at(cursor(), out); }
    void allowed_at(const Cursor& c, uint32_t out[8]) const {
        for (int i = 0; i < 8; ++i) out[i] = 0;
        if (done_at(c)) return;
        auto set = [&](int b) { out[b >> 5] |= 1u << (b & 31); };
        const Seg& s = segs_[c.seg];
        if (s.type == 0) { set((unsigned char)s.lit[c.off]); }
        else if (s.type == 1) {
            if (c.off < s.max_len) for (int b = 0x20; b <= 0x7E; ++b) if (string_byte(b)) set(b);
            if (c.off >= s.min_len) set((unsigned char)close_);   // the closing quote (or line end) opens the following literal
        } else {
            for (size_t t = 0; t < opts_.size(); ++t) if (c.cand & (1u << t)) set((unsigned char)opts_[t][c.off]);
        }
    }
    // consume one emitted byte (must be in the allowed set); returns false if it was not
    bool advance(int b) { Cursor c = cursor(); if (!step(c, b)) return false; set_cursor(c); return true; }
    bool step(Cursor& c, int b) const {
        if (done_at(c)) return false;
        const Seg& s = segs_[c.seg];
        auto next = [&]() { ++c.seg; c.off = 0; c.cand = all_opts(); };
        if (s.type == 0) {
            if ((unsigned char)s.lit[c.off] != b) return false;
            if (++c.off == (int)s.lit.size()) next();
        } else if (s.type == 1) {
            if (b == (unsigned char)close_ && c.off >= s.min_len) { next(); c.off = 1; if (c.off == (int)segs_[c.seg].lit.size()) next(); }
            else if (string_byte(b) && c.off < s.max_len) ++c.off;
            else return false;
        } else {
            uint32_t keep = 0;
            for (size_t t = 0; t < opts_.size(); ++t) if ((c.cand & (1u << t)) && (unsigned char)opts_[t][c.off] == b) keep |= 1u << t;
            if (!keep) return false;
            c.cand = keep; ++c.off;
            for (size_t t = 0; t < opts_.size(); ++t) if ((c.cand & (1u << t)) && c.off == (int)opts_[t].size()) { next(); break; }
        }
        return true;
    }
    int kind() const { return kind_; }
    // inside a string segment: its [min_len, max_len] (token_mask.hpp canonicalises mask-cache keys with them)
    bool string_bounds(const Cursor& c, int& lo, int& hi) const {
        if (done_at(c) || segs_[c.seg].type != 1) return false;
        lo = segs_[c.seg].min_len; hi = segs_[c.seg].max_len; return true;
    }

private:
    uint32_t all_opts() const { return opts_.size() >= 32 ? 0xffffffffu : ((1u << opts_.size()) - 1); }
    int kind_; std::string sig_; std::vector<std::string> opts_; char close_ = '"'; std::vector<Seg> segs_; int seg_ = 0, off_ = 0; uint32_t cand_ = 0;
};

}  // namespace oa
