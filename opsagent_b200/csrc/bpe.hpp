// bpe.hpp — byte-level BPE tokenizer that reads a Hugging Face `tokenizer.json` (the Llama-3 and Qwen2.5 format):
//   pre-tokenizer  Sequence[ Split(Regex(<pattern>), Isolated), ByteLevel(add_prefix_space=false, use_regex=false) ]
//   model          BPE { vocab: {token: id}, merges: [[a, b] | "a b", ...], ignore_merges }
//   added_tokens   control tokens (looked up by content; never produced from text)
//   normalizer     null or NFC (already-NFC text passes, anything else is refused: see from_json)
// Host only (no CUDA).  The two published split patterns are recognised by text and matched by a hand-written scanner with the
// regex engine's semantics (leftmost alternative first, greedy with backtracking, Unicode \p{L} \p{N} \s); anything else is
// rejected loudly.  Pinned against the `tokenizers` library itself: tests/golden/gen_golden_bpe.py trains two tokenizers in exactly
// this configuration and records that library's encodings (tests/test_bpe.py).
//
// Why it exists: the reference counts tokens with tiktoken BPE for OpenAI models and silently skips truncation for Llama/Qwen names
// (pkg/llms/tokens.go:60-66,128-144); with a real checkpoint (config "weights") the engine needs the checkpoint's own vocabulary.
#pragma once
#include <algorithm>
#include <cstdint>
#include <cstring>
#include <fstream>
#include <memory>
#include <sstream>
#include <stdexcept>
#include <string>
#include <unordered_map>
#include <utility>
#include <vector>

namespace oa {

namespace bpe_detail {

#include "unicode_tables.inc"

inline bool in_ranges(const uint32_t (*r)[2], size_t n, uint32_t cp) {
    size_t lo = 0, hi = n;
    while (lo < hi) { const size_t mid = (lo + hi) / 2; if (cp < r[mid][0]) hi = mid; else if (cp > r[mid][1]) lo = mid + 1; else return true; }
    return false;
}
inline bool is_L(uint32_t cp) { return in_ranges(kUnicodeL, sizeof(kUnicodeL) / sizeof(kUnicodeL[0]), cp); }
inline bool is_N(uint32_t cp) { return in_ranges(kUnicodeN, sizeof(kUnicodeN) / sizeof(kUnicodeN[0]), cp); }
inline bool is_S(uint32_t cp) {      // Unicode White_Space
    return (cp >= 0x9 && cp <= 0xD) || cp == 0x20 || cp == 0x85 || cp == 0xA0 || cp == 0x1680 || (cp >= 0x2000 && cp <= 0x200A) || cp == 0x2028 ||
           cp == 0x2029 || cp == 0x202F || cp == 0x205F || cp == 0x3000;
}
inline bool is_NL(uint32_t cp) { return cp == '\r' || cp == '\n'; }
inline bool nfc_unsafe(uint32_t cp) { return cp >= 0x300 && in_ranges(kUnicodeNfcUnsafe, sizeof(kUnicodeNfcUnsafe) / sizeof(kUnicodeNfcUnsafe[0]), cp); }

inline void utf8_append(std::string& o, uint32_t cp) {
    if (cp < 0x80) o += (char)cp;
    else if (cp < 0x800) { o += (char)(0xC0 | (cp >> 6)); o += (char)(0x80 | (cp & 63)); }
    else if (cp < 0x10000) { o += (char)(0xE0 | (cp >> 12)); o += (char)(0x80 | ((cp >> 6) & 63)); o += (char)(0x80 | (cp & 63)); }
    else { o += (char)(0xF0 | (cp >> 18)); o += (char)(0x80 | ((cp >> 12) & 63)); o += (char)(0x80 | ((cp >> 6) & 63)); o += (char)(0x80 | (cp & 63)); }
}
// decode UTF-8 into code points + the byte offset of each (an invalid byte becomes one U+FFFD code point; its raw byte still reaches the
// byte-level model, so nothing is lost)
inline void utf8_decode(const std::string& s, std::vector<uint32_t>& cps, std::vector<uint32_t>& off) {
    size_t i = 0; const size_t n = s.size();
    while (i < n) {
        const unsigned char c = (unsigned char)s[i]; uint32_t cp = c; size_t len = 1;
        if (c >= 0xF0 && c < 0xF8) { len = 4; cp = c & 7; } else if (c >= 0xE0 && c < 0xF0) { len = 3; cp = c & 15; } else if (c >= 0xC0 && c < 0xE0) { len = 2; cp = c & 31; }
        bool ok = c < 0x80 || (c >= 0xC0 && i + len <= n);
        for (size_t k = 1; ok && k < len; ++k) { const unsigned char d = (unsigned char)s[i + k]; if ((d & 0xC0) != 0x80) ok = false; else cp = (cp << 6) | (d & 63); }
        if (!ok) { cp = 0xFFFD; len = 1; }
        cps.push_back(cp); off.push_back((uint32_t)i); i += len;
    }
    off.push_back((uint32_t)n);
}

// ---- a small JSON DOM (tokenizer.json is an object tree with a 100k-entry vocabulary and merge list) ----
struct JVal {
    enum Kind { Null, Bool, Num, Str, Arr, Obj } kind = Null;
    bool b = false; double num = 0; std::string str;
    std::vector<JVal> arr; std::vector<std::pair<std::string, JVal>> obj;
    const JVal* get(const char* k) const { for (auto& kv : obj) if (kv.first == k) return &kv.second; return nullptr; }
};
class JParser {
public:
    explicit JParser(const std::string& s) : s_(s) {}
    JVal parse() { JVal v = value(); ws(); if (i_ != s_.size()) fail("trailing characters"); return v; }
private:
    [[noreturn]] void fail(const char* m) const { throw std::runtime_error(std::string("tokenizer.json: ") + m + " at byte " + std::to_string(i_)); }
    void ws() { while (i_ < s_.size() && (s_[i_] == ' ' || s_[i_] == '\n' || s_[i_] == '\t' || s_[i_] == '\r')) ++i_; }
    uint32_t hex4() {
        if (i_ + 4 > s_.size()) fail("bad \\u escape");
        uint32_t v = 0;
        for (int k = 0; k < 4; ++k) { const char c = s_[i_++]; v = v * 16 + (c >= '0' && c <= '9' ? c - '0' : c >= 'a' && c <= 'f' ? c - 'a' + 10 : c >= 'A' && c <= 'F' ? c - 'A' + 10 : (fail("bad hex digit"), 0)); }
        return v;
    }
    std::string string() {
        if (s_[i_] != '"') fail("expected string");
        ++i_; std::string o;
        while (i_ < s_.size() && s_[i_] != '"') {
            if (s_[i_] == '\\') {
                ++i_; if (i_ >= s_.size()) fail("unterminated escape");
                const char c = s_[i_++];
                switch (c) {
                    case 'n': o += '\n'; break; case 't': o += '\t'; break; case 'r': o += '\r'; break; case 'b': o += '\b'; break; case 'f': o += '\f'; break;
                    case 'u': {
                        uint32_t cp = hex4();
                        if (cp >= 0xD800 && cp < 0xDC00 && i_ + 1 < s_.size() && s_[i_] == '\\' && s_[i_ + 1] == 'u') { i_ += 2; const uint32_t lo = hex4(); cp = 0x10000 + ((cp - 0xD800) << 10) + (lo - 0xDC00); }
                        utf8_append(o, cp); break;
                    }
                    default: o += c;      // \" \\ \/
                }
            } else o += s_[i_++];
        }
        if (i_ >= s_.size()) fail("unterminated string");
        ++i_; return o;
    }
    JVal value() {
        ws(); if (i_ >= s_.size()) fail("unexpected end");
        JVal v; const char c = s_[i_];
        if (c == '{') {
            v.kind = JVal::Obj; ++i_; ws();
            if (i_ < s_.size() && s_[i_] == '}') { ++i_; return v; }
            while (true) {
                ws(); std::string k = string(); ws();
                if (i_ >= s_.size() || s_[i_] != ':') fail("expected ':'");
                ++i_; v.obj.emplace_back(std::move(k), value()); ws();
                if (i_ < s_.size() && s_[i_] == ',') { ++i_; continue; }
                if (i_ < s_.size() && s_[i_] == '}') { ++i_; break; }
                fail("expected ',' or '}'");
            }
        } else if (c == '[') {
            v.kind = JVal::Arr; ++i_; ws();
            if (i_ < s_.size() && s_[i_] == ']') { ++i_; return v; }
            while (true) {
                v.arr.push_back(value()); ws();
                if (i_ < s_.size() && s_[i_] == ',') { ++i_; continue; }
                if (i_ < s_.size() && s_[i_] == ']') { ++i_; break; }
                fail("expected ',' or ']'");
            }
        } else if (c == '"') { v.kind = JVal::Str; v.str = string(); }
        else if (s_.compare(i_, 4, "true") == 0) { v.kind = JVal::Bool; v.b = true; i_ += 4; }
        else if (s_.compare(i_, 5, "false") == 0) { v.kind = JVal::Bool; v.b = false; i_ += 5; }
        else if (s_.compare(i_, 4, "null") == 0) { i_ += 4; }
        else {
            const size_t b = i_;
            while (i_ < s_.size() && (std::strchr("+-.eE", s_[i_]) || (s_[i_] >= '0' && s_[i_] <= '9'))) ++i_;
            if (b == i_) fail("unexpected character");
            v.kind = JVal::Num; v.num = std::stod(s_.substr(b, i_ - b));
        }
        return v;
    }
    const std::string& s_; size_t i_ = 0;
};

}  // namespace bpe_detail

class BpeTokenizer {
public:
    // throws std::runtime_error with a message naming what is unsupported or malformed
    static std::shared_ptr<BpeTokenizer> load(const std::string& path) {
        std::ifstream f(path, std::ios::binary);
        if (!f) throw std::runtime_error("tokenizer file not found: " + path);
        std::stringstream ss; ss << f.rdbuf();
        return from_json(ss.str());
    }
    static std::shared_ptr<BpeTokenizer> from_json(const std::string& text) {
        using namespace bpe_detail;
        const JVal root = JParser(text).parse();
        auto t = std::shared_ptr<BpeTokenizer>(new BpeTokenizer());
        const JVal* model = root.get("model");
        if (!model || model->kind != JVal::Obj) throw std::runtime_error("tokenizer.json: no model");
        const JVal* type = model->get("type");
        if (type && type->kind == JVal::Str && type->str != "BPE") throw std::runtime_error("tokenizer.json: model type '" + type->str + "' is not BPE");
        if (const JVal* ig = model->get("ignore_merges")) t->ignore_merges_ = ig->kind == JVal::Bool && ig->b;
        // ---- normaliser: none (Llama-3) or NFC (Qwen2.5).  NFC itself is not implemented: text that is already NFC — no code point that
        // can decompose, reorder or compose (table from tools/gen_unicode_tables.py) — passes unchanged, anything else is refused ----
        if (const JVal* nz = root.get("normalizer")) if (nz->kind == JVal::Obj) {
            const JVal* ty = nz->get("type");
            if (ty && ty->kind == JVal::Str && ty->str == "NFC") t->nfc_ = true;
            else throw std::runtime_error("tokenizer.json: unsupported normalizer" + (ty && ty->kind == JVal::Str ? " '" + ty->str + "'" : std::string()));
        }
        // ---- split pattern ----
        std::string pattern; bool byte_level = false;
        std::vector<const JVal*> pts;
        if (const JVal* pt = root.get("pre_tokenizer")) {
            if (const JVal* seq = pt->get("pretokenizers")) for (auto& e : seq->arr) pts.push_back(&e); else pts.push_back(pt);
        }
        for (const JVal* e : pts) {
            const JVal* ty = e->get("type");
            if (!ty || ty->kind != JVal::Str) continue;
            if (ty->str == "Split") { if (const JVal* p = e->get("pattern")) if (const JVal* r = p->get("Regex")) pattern = r->str; }
            else if (ty->str == "ByteLevel") {
                byte_level = true;
                const JVal* ur = e->get("use_regex"); const JVal* ps = e->get("add_prefix_space");
                if ((ur && ur->kind == JVal::Bool && ur->b && pts.size() > 1) || (ps && ps->kind == JVal::Bool && ps->b))
                    throw std::runtime_error("tokenizer.json: ByteLevel with use_regex/add_prefix_space is not supported");
            }
        }
        if (!byte_level) throw std::runtime_error("tokenizer.json: only byte-level BPE (ByteLevel pre-tokenizer) is supported");
        const std::string head = "(?i:'s|'t|'re|'ve|'m|'ll|'d)|[^\\r\\n\\p{L}\\p{N}]?\\p{L}+|";
        const std::string tail = "| ?[^\\s\\p{L}\\p{N}]+[\\r\\n]*|\\s*[\\r\\n]+|\\s+(?!\\S)|\\s+";
        if (pattern == head + "\\p{N}{1,3}" + tail) t->digits_ = 3;           // Llama-3
        else if (pattern == head + "\\p{N}" + tail) t->digits_ = 1;           // Qwen2 / Qwen2.5
        else throw std::runtime_error("tokenizer.json: unsupported split pattern: " + pattern);
        // ---- vocabulary: token strings are bytes written in the GPT-2 byte<->unicode alphabet ----
        uint32_t byte_of_cp[512]; for (auto& v : byte_of_cp) v = 0xFFFFFFFFu;
        {
            int n = 0;
            for (int b = 0; b < 256; ++b) {
                const bool printable = (b >= '!' && b <= '~') || (b >= 0xA1 && b <= 0xAC) || (b >= 0xAE && b <= 0xFF);
                const uint32_t cp = printable ? (uint32_t)b : (uint32_t)(256 + n++);
                byte_of_cp[cp] = (uint32_t)b;
            }
        }
        auto to_bytes = [&](const std::string& tok, std::string& out) -> bool {
            std::vector<uint32_t> cps, off; utf8_decode(tok, cps, off);
            out.clear();
            for (uint32_t cp : cps) { if (cp >= 512 || byte_of_cp[cp] == 0xFFFFFFFFu) return false; out += (char)byte_of_cp[cp]; }
            return true;
        };
        const JVal* vocab = model->get("vocab");
        if (!vocab || vocab->kind != JVal::Obj) throw std::runtime_error("tokenizer.json: model.vocab missing");
        int max_id = -1;
        for (auto& kv : vocab->obj) max_id = std::max(max_id, (int)kv.second.num);
        const JVal* added = root.get("added_tokens");
        if (added) for (auto& a : added->arr) if (const JVal* id = a.get("id")) max_id = std::max(max_id, (int)id->num);
        t->id_to_bytes_.assign((size_t)max_id + 1, std::string());
        t->is_special_.assign((size_t)max_id + 1, 0);
        if (added) for (auto& a : added->arr) {
            const JVal* id = a.get("id"); const JVal* c = a.get("content");
            if (!id || !c) continue;
            t->special_[c->str] = (int)id->num; t->is_special_[(size_t)id->num] = 1; t->id_to_bytes_[(size_t)id->num] = c->str;
        }
        t->vocab_.reserve(vocab->obj.size() * 2);
        std::string bytes;
        for (auto& kv : vocab->obj) {
            const int id = (int)kv.second.num;
            if (t->is_special_[(size_t)id]) continue;
            if (!to_bytes(kv.first, bytes)) throw std::runtime_error("tokenizer.json: vocabulary entry outside the byte-level alphabet: " + kv.first);
            t->vocab_[bytes] = id; t->id_to_bytes_[(size_t)id] = bytes;
        }
        for (int b = 0; b < 256; ++b) {
            auto it = t->vocab_.find(std::string(1, (char)b));
            if (it == t->vocab_.end()) throw std::runtime_error("tokenizer.json: byte " + std::to_string(b) + " has no token (not a byte-level vocabulary)");
            t->byte_id_[b] = it->second;
        }
        // ---- merges: rank = position; both halves and the result must be vocabulary entries ----
        const JVal* merges = model->get("merges");
        if (!merges || merges->kind != JVal::Arr) throw std::runtime_error("tokenizer.json: model.merges missing");
        t->merge_.reserve(merges->arr.size() * 2);
        std::string a, b;
        for (size_t r = 0; r < merges->arr.size(); ++r) {
            const JVal& m = merges->arr[r];
            std::string sa, sb;
            if (m.kind == JVal::Arr && m.arr.size() == 2) { sa = m.arr[0].str; sb = m.arr[1].str; }
            else if (m.kind == JVal::Str) { const size_t sp = m.str.find(' '); if (sp == std::string::npos) throw std::runtime_error("tokenizer.json: bad merge entry"); sa = m.str.substr(0, sp); sb = m.str.substr(sp + 1); }
            else throw std::runtime_error("tokenizer.json: bad merge entry");
            if (!to_bytes(sa, a) || !to_bytes(sb, b)) throw std::runtime_error("tokenizer.json: merge outside the byte-level alphabet");
            auto ia = t->vocab_.find(a), ib = t->vocab_.find(b), ic = t->vocab_.find(a + b);
            if (ia == t->vocab_.end() || ib == t->vocab_.end() || ic == t->vocab_.end()) continue;      // as the library does: unusable merges are skipped
            t->merge_.emplace(((uint64_t)(uint32_t)ia->second << 32) | (uint32_t)ib->second, std::make_pair((int)r, ic->second));
        }
        return t;
    }

    int vocab_size() const { return (int)id_to_bytes_.size(); }
    // raw bytes of every TEXT token (control / added tokens and unused ids: empty) — what a grammar may emit
    std::vector<std::string> text_token_bytes() const { std::vector<std::string> v = id_to_bytes_; for (size_t i = 0; i < v.size(); ++i) if (is_special_[i]) v[i].clear(); return v; }
    int special_id(const std::string& content) const { auto it = special_.find(content); return it == special_.end() ? -1 : it->second; }

    // text -> ids (control tokens are never produced from text: the chat template inserts them by id)
    void encode(const std::string& text, std::vector<int32_t>& out) const {
        using namespace bpe_detail;
        std::vector<uint32_t> cp, off; utf8_decode(text, cp, off);
        const size_t n = cp.size();
        if (nfc_) for (uint32_t c : cp) if (nfc_unsafe(c)) throw std::runtime_error("text is not in Unicode NFC (this tokenizer normalises to NFC; normalisation is not implemented)");
        size_t i = 0;
        std::vector<int32_t> word;
        while (i < n) {
            const size_t j = match(cp, i);
            encode_piece(text.data() + off[i], off[j] - off[i], word, out);
            i = j;
        }
    }
    // ids -> bytes (control tokens render as their literal content)
    std::string decode(const int32_t* ids, size_t n) const {
        std::string s;
        for (size_t k = 0; k < n; ++k) if (ids[k] >= 0 && (size_t)ids[k] < id_to_bytes_.size()) s += id_to_bytes_[(size_t)ids[k]];
        return s;
    }

private:
    BpeTokenizer() = default;
    // end (exclusive, in code points) of the pre-token starting at i — the split regex, alternative by alternative
    size_t match(const std::vector<uint32_t>& c, size_t i) const {
        using namespace bpe_detail;
        const size_t n = c.size();
        auto lower = [](uint32_t x) { return (x >= 'A' && x <= 'Z') ? x + 32 : x == 0x17F ? (uint32_t)'s' : x; };      // (?i) also folds U+017F to s
        // 1. (?i:'s|'t|'re|'ve|'m|'ll|'d)
        if (c[i] == '\'' && i + 1 < n) {
            const uint32_t a = lower(c[i + 1]), b = i + 2 < n ? lower(c[i + 2]) : 0;
            if (a == 's' || a == 't') return i + 2;
            if ((a == 'r' && b == 'e') || (a == 'v' && b == 'e')) return i + 3;
            if (a == 'm') return i + 2;
            if (a == 'l' && b == 'l') return i + 3;
            if (a == 'd') return i + 2;
        }
        // 2. [^\r\n\p{L}\p{N}]?\p{L}+
        {
            size_t p = i;
            if (!is_NL(c[p]) && !is_L(c[p]) && !is_N(c[p]) && p + 1 < n && is_L(c[p + 1])) ++p;
            if (is_L(c[p])) { while (p < n && is_L(c[p])) ++p; return p; }
        }
        // 3. \p{N}{1,3}  (Qwen2: \p{N})
        if (is_N(c[i])) { size_t p = i; while (p < n && p - i < (size_t)digits_ && is_N(c[p])) ++p; return p; }
        // 4.  ?[^\s\p{L}\p{N}]+[\r\n]*
        {
            auto punct = [&](uint32_t x) { return !is_S(x) && !is_L(x) && !is_N(x); };
            size_t p = i;
            if (c[p] == ' ' && p + 1 < n && punct(c[p + 1])) ++p;
            if (punct(c[p])) { while (p < n && punct(c[p])) ++p; while (p < n && is_NL(c[p])) ++p; return p; }
        }
        // 5-7 all start with whitespace: W = the maximal whitespace run at i
        size_t w = i; while (w < n && is_S(c[w])) ++w;
        if (w == i) return i + 1;      // unreachable for valid input (every code point is a letter, number, space or "other"); never loop
        // 5. \s*[\r\n]+ : up to and including the LAST newline of the run (greedy \s* gives back only what it must)
        for (size_t p = w; p > i; --p) if (is_NL(c[p - 1])) return p;
        // 6. \s+(?!\S) : the whole run at the end of the text, else all but its last character
        if (w == n) return w;
        if (w - i >= 2) return w - 1;
        // 7. \s+
        return w;
    }
    // one pre-token (raw bytes) -> ids appended to out
    void encode_piece(const char* p, size_t len, std::vector<int32_t>& word, std::vector<int32_t>& out) const {
        if (len == 0) return;
        if (ignore_merges_ || len == 1) {
            auto it = vocab_.find(std::string(p, len));
            if (it != vocab_.end()) { out.push_back(it->second); return; }
        }
        word.clear();
        for (size_t k = 0; k < len; ++k) word.push_back(byte_id_[(unsigned char)p[k]]);
        if (len > 24) { merge_long(word); out.insert(out.end(), word.begin(), word.end()); return; }
        while (word.size() > 1) {      // merge the lowest-ranked adjacent pair (leftmost on ties) until none is mergeable
            int best = -1, best_rank = 0x7fffffff, best_id = 0;
            for (size_t k = 0; k + 1 < word.size(); ++k) {
                auto it = merge_.find(((uint64_t)(uint32_t)word[k] << 32) | (uint32_t)word[k + 1]);
                if (it != merge_.end() && it->second.first < best_rank) { best = (int)k; best_rank = it->second.first; best_id = it->second.second; }
            }
            if (best < 0) break;
            word[(size_t)best] = best_id; word.erase(word.begin() + best + 1);
        }
        out.insert(out.end(), word.begin(), word.end());
    }
    // The same merge order for long pre-tokens in O(n log n): doubly linked symbols + a min-heap of candidate pairs keyed by
    // (rank, position of the left symbol); stale heap entries (a symbol changed since the push) are skipped on pop.  A 1 MB run of
    // '=' or 'A' in a tool observation must not pin the calling thread for minutes (the quadratic rescan above would).
    void merge_long(std::vector<int32_t>& word) const {
        const int n = (int)word.size();
        std::vector<int> prev(n), next(n);
        for (int i = 0; i < n; ++i) { prev[i] = i - 1; next[i] = i + 1 < n ? i + 1 : -1; }
        struct Cand { int rank, pos, left_id, right_id, merged; };
        auto worse = [](const Cand& a, const Cand& b) { return a.rank != b.rank ? a.rank > b.rank : a.pos > b.pos; };
        std::vector<Cand> heap;
        auto push = [&](int i) {
            const int j = next[i]; if (j < 0) return;
            auto it = merge_.find(((uint64_t)(uint32_t)word[i] << 32) | (uint32_t)word[j]);
            if (it != merge_.end()) { heap.push_back(Cand{it->second.first, i, word[i], word[j], it->second.second}); std::push_heap(heap.begin(), heap.end(), worse); }
        };
        for (int i = 0; i + 1 < n; ++i) push(i);
        std::vector<uint8_t> dead(n, 0);
        while (!heap.empty()) {
            std::pop_heap(heap.begin(), heap.end(), worse);
            const Cand c = heap.back(); heap.pop_back();
            const int i = c.pos, j = dead[i] ? -1 : next[i];
            if (j < 0 || word[i] != c.left_id || word[j] != c.right_id) continue;      // stale
            word[i] = c.merged; dead[j] = 1;
            next[i] = next[j]; if (next[j] >= 0) prev[next[j]] = i;
            if (prev[i] >= 0) push(prev[i]);
            push(i);
        }
        int w = 0;
        for (int i = 0; i >= 0 && i < n; i = next[i]) word[w++] = word[i];
        word.resize(w);
    }

    int digits_ = 3; bool ignore_merges_ = false, nfc_ = false;
    std::unordered_map<std::string, int> vocab_;                              // raw bytes -> id
    std::unordered_map<uint64_t, std::pair<int, int>> merge_;                 // (id_a, id_b) -> (rank, merged id)
    std::unordered_map<std::string, int> special_;
    std::vector<std::string> id_to_bytes_; std::vector<uint8_t> is_special_;
    int byte_id_[256] = {};
};

}  // namespace oa
