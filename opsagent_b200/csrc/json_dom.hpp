// json_dom.hpp — a small JSON reader / writer shared by the native HTTP front (http_server.cpp) and the host-side mirrors of the reference's caller
// (host/assistants.hpp).  Reader: full syntax check, \uXXXX incl. surrogate pairs, nesting bounded at 64, strict numbers; raw bytes >= 0x80 inside
// strings pass through.  Writer: well-formed UTF-8 passes through, ill-formed bytes become U+FFFD (a byte-level model can stop mid-character).
#pragma once
#include <cmath>
#include <cstdint>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <string>
#include <utility>
#include <vector>

namespace oa {

// ---- a small JSON DOM (requests are a few KB; 16k-token observations ~100 KB) ------------------------------------------------------------
struct Json {
    enum T { Null, Bool, Num, Str, Arr, Obj } t = Null;
    bool b = false; double n = 0; std::string s; std::vector<Json> a; std::vector<std::pair<std::string, Json>> o;
    const Json* get(const char* k) const { if (t != Obj) return nullptr; for (auto& kv : o) if (kv.first == k) return &kv.second; return nullptr; }
    std::string str(const char* k, const std::string& d = "") const { const Json* j = get(k); return j && j->t == Str ? j->s : d; }
};
struct JsonParser {
    const char* p; const char* e; std::string err; int depth = 0;
    void ws() { while (p < e && (*p == ' ' || *p == '\n' || *p == '\t' || *p == '\r')) ++p; }
    static void utf8(std::string& o, uint32_t c) {
        if (c < 0x80) o += (char)c;
        else if (c < 0x800) { o += (char)(0xC0 | (c >> 6)); o += (char)(0x80 | (c & 63)); }
        else if (c < 0x10000) { o += (char)(0xE0 | (c >> 12)); o += (char)(0x80 | ((c >> 6) & 63)); o += (char)(0x80 | (c & 63)); }
        else { o += (char)(0xF0 | (c >> 18)); o += (char)(0x80 | ((c >> 12) & 63)); o += (char)(0x80 | ((c >> 6) & 63)); o += (char)(0x80 | (c & 63)); }
    }
    bool hex4(uint32_t& v) {
        if (e - p < 4) return false;
        v = 0;
        for (int i = 0; i < 4; ++i) { const char c = *p++; v <<= 4; if (c >= '0' && c <= '9') v |= c - '0'; else if (c >= 'a' && c <= 'f') v |= c - 'a' + 10; else if (c >= 'A' && c <= 'F') v |= c - 'A' + 10; else return false; }
        return true;
    }
    bool string(std::string& o) {
        if (p >= e || *p != '"') { err = "expected string"; return false; }
        ++p;
        while (p < e && *p != '"') {
            if (*p == '\\') {
                if (++p >= e) break;
                const char c = *p++;
                switch (c) {
                    case 'n': o += '\n'; break; case 't': o += '\t'; break; case 'r': o += '\r'; break; case 'b': o += '\b'; break; case 'f': o += '\f'; break;
                    case '"': case '\\': case '/': o += c; break;
                    case 'u': {
                        uint32_t v; if (!hex4(v)) { err = "bad \\u escape"; return false; }
                        if (v >= 0xD800 && v <= 0xDBFF && e - p >= 6 && p[0] == '\\' && p[1] == 'u') {      // surrogate pair
                            const char* save = p; p += 2; uint32_t lo;
                            if (hex4(lo) && lo >= 0xDC00 && lo <= 0xDFFF) v = 0x10000 + ((v - 0xD800) << 10) + (lo - 0xDC00); else { p = save; v = 0xFFFD; }
                        } else if (v >= 0xD800 && v <= 0xDFFF) v = 0xFFFD;
                        utf8(o, v); break;
                    }
                    default: err = "bad escape"; return false;
                }
            } else o += *p++;
        }
        if (p >= e) { err = "unterminated string"; return false; }
        ++p; return true;
    }
    bool value(Json& j) {
        if (++depth > 64) { err = "nesting too deep"; return false; }
        ws();
        if (p >= e) { err = "unexpected end"; return false; }
        bool ok = true;
        if (*p == '"') { j.t = Json::Str; ok = string(j.s); }
        else if (*p == '{') {
            j.t = Json::Obj; ++p; ws();
            if (p < e && *p == '}') ++p;
            else while (ok) {
                ws(); std::string k; if (!string(k)) { ok = false; break; }
                ws(); if (p >= e || *p != ':') { err = "expected ':'"; ok = false; break; }
                ++p; Json v; if (!value(v)) { ok = false; break; }
                j.o.emplace_back(std::move(k), std::move(v)); ws();
                if (p < e && *p == ',') { ++p; continue; }
                if (p < e && *p == '}') { ++p; break; }
                err = "expected ',' or '}'"; ok = false;
            }
        } else if (*p == '[') {
            j.t = Json::Arr; ++p; ws();
            if (p < e && *p == ']') ++p;
            else while (ok) {
                Json v; if (!value(v)) { ok = false; break; }
                j.a.push_back(std::move(v)); ws();
                if (p < e && *p == ',') { ++p; continue; }
                if (p < e && *p == ']') { ++p; break; }
                err = "expected ',' or ']'"; ok = false;
            }
        } else if (e - p >= 4 && !std::strncmp(p, "true", 4)) { j.t = Json::Bool; j.b = true; p += 4; }
        else if (e - p >= 5 && !std::strncmp(p, "false", 5)) { j.t = Json::Bool; j.b = false; p += 5; }
        else if (e - p >= 4 && !std::strncmp(p, "null", 4)) { j.t = Json::Null; p += 4; }
        else {
            const char* b = p;
            while (p < e && *p != 0 && (std::strchr("+-.eE", *p) || (*p >= '0' && *p <= '9'))) ++p;
            const std::string tok(b, p); char* endp = nullptr;
            if (!tok.empty()) j.n = std::strtod(tok.c_str(), &endp);
            if (tok.empty() || tok[0] == '+' || tok[0] == '.' || endp != tok.c_str() + tok.size() || !std::isfinite(j.n)) { err = "unexpected character"; ok = false; }
            else j.t = Json::Num;
        }
        --depth; return ok;
    }
};
inline bool parse_json(const std::string& s, Json& out, std::string& err) {
    JsonParser P{s.data(), s.data() + s.size(), "", 0};
    if (!P.value(out)) { err = P.err; return false; }
    P.ws();
    if (P.p != P.e) { err = "trailing characters"; return false; }
    return true;
}
// length of the well-formed UTF-8 sequence at s[i..] (0: ill-formed); *bad = bytes of the maximal ill-formed subpart (what one U+FFFD replaces)
inline int utf8_seq(const std::string& s, size_t i, int* bad) {
    const unsigned char c = (unsigned char)s[i];
    auto at = [&](size_t k) -> int { return i + k < s.size() ? (unsigned char)s[i + k] : -1; };
    auto cont = [](int b) { return b >= 0x80 && b <= 0xBF; };
    *bad = 1;
    if (c < 0x80) return 1;
    if (c >= 0xC2 && c <= 0xDF) return cont(at(1)) ? 2 : 0;
    if (c >= 0xE0 && c <= 0xEF) {
        const int b1 = at(1), lo = c == 0xE0 ? 0xA0 : 0x80, hi = c == 0xED ? 0x9F : 0xBF;
        if (b1 < lo || b1 > hi) return 0;
        if (!cont(at(2))) { *bad = 2; return 0; }
        return 3;
    }
    if (c >= 0xF0 && c <= 0xF4) {
        const int b1 = at(1), lo = c == 0xF0 ? 0x90 : 0x80, hi = c == 0xF4 ? 0x8F : 0xBF;
        if (b1 < lo || b1 > hi) return 0;
        if (!cont(at(2))) { *bad = 2; return 0; }
        if (!cont(at(3))) { *bad = 3; return 0; }
        return 4;
    }
    return 0;
}
inline std::string jstr(const std::string& s) {      // JSON string literal; well-formed UTF-8 passes through, ill-formed bytes (a byte-level model cut mid-character) become U+FFFD
    std::string o = "\"";
    for (size_t i = 0; i < s.size();) {
        const unsigned char c = (unsigned char)s[i];
        if (c >= 0x80) {
            int bad; const int n = utf8_seq(s, i, &bad);
            if (n) { o.append(s, i, (size_t)n); i += (size_t)n; } else { o += "\xEF\xBF\xBD"; i += (size_t)bad; }
            continue;
        }
        switch (c) {
            case '"': o += "\\\""; break; case '\\': o += "\\\\"; break; case '\n': o += "\\n"; break; case '\r': o += "\\r"; break; case '\t': o += "\\t"; break;
            default: if (c < 0x20) { char b[8]; std::snprintf(b, sizeof b, "\\u%04x", c); o += b; } else o += (char)c;
        }
        ++i;
    }
    return o + "\"";
}
inline void json_dump(const Json& j, std::string& o) {     // re-serialise a value (tool-call arguments are returned as a JSON string)
    switch (j.t) {
        case Json::Null: o += "null"; break;
        case Json::Bool: o += j.b ? "true" : "false"; break;
        case Json::Num: { char b[32]; if (j.n == std::floor(j.n) && std::fabs(j.n) < 1e15) std::snprintf(b, sizeof b, "%.0f", j.n); else std::snprintf(b, sizeof b, "%.17g", j.n); o += b; break; }
        case Json::Str: o += jstr(j.s); break;
        case Json::Arr: o += "["; for (size_t i = 0; i < j.a.size(); ++i) { if (i) o += ", "; json_dump(j.a[i], o); } o += "]"; break;
        case Json::Obj: o += "{"; for (size_t i = 0; i < j.o.size(); ++i) { if (i) o += ", "; o += jstr(j.o[i].first) + ": "; json_dump(j.o[i].second, o); } o += "}"; break;
    }
}

}  // namespace oa
