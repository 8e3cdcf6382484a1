// assistants.hpp — C++ host-side mirror of the CALLER of the Chat seam: the reference's ReAct loop
//     func AssistantWithConfig(model, prompts, maxTokens, countTokens, verbose, maxIterations, apiKey, baseUrl) (string, []ChatCompletionMessage, error)
//                                                                                             reference pkg/assistants/simple.go:292-616
// with the pieces it needs from its neighbours:
//     type ToolPrompt struct{ Question, Thought, Action{Name, Input}, Observation, FinalAnswer }          reference pkg/tools/tool.go:29-38
//     json.Marshal / json.Unmarshal of it (encoding/json of go 1.24: go.mod:3)                            simple.go:366, 497, 541
//     isTemplateValue                                                                                     simple.go:624-657
//     llms.NumTokensFromMessages / llms.ConstrictPrompt                                                   reference pkg/llms/tokens.go:60, 128-144
// The reference is Go (no toolchain in this image), so — like localcuda_client.hpp for the seam itself — the host side above the C ABI is written in
// C++ with the same names, argument meaning and error behaviour.  The LLM call and the tools are injected (`ChatFn`, `Tool`): the reference's
// pkg/tools shell out to kubectl / trivy (out of scope) and its client is built from apiKey/baseUrl on every call (simple.go:316),
// which here is LocalCUDAClient::Chat bound to the process-wide engine.  opsagent_b200/assistants.py is the same loop in Python;
// tests/test_host_cpp.py runs both on the same scripted scenarios and demands identical results and histories.
#pragma once
#include <algorithm>
#include <functional>
#include <map>
#include <string>
#include <vector>

#include "../csrc/json_dom.hpp"
#include "localcuda_client.hpp"
#include "perf.hpp"

namespace opsagent {

constexpr int defaultMaxIterations = 5;                                   // simple.go:22

// encoding/json's string encoder with escapeHTML = true (the json.Marshal default), go >= 1.22: '<' '>' '&' become backslash-u 003c / 003e / 0026,
// U+2028 / U+2029 become backslash-u 2028 / 2029, \b \f \n \r \t use the short forms, other control bytes backslash-u 00XX, every invalid UTF-8
// byte becomes backslash-u fffd, everything else stays raw
inline std::string GoJSONString(const std::string& s) {
    std::string o = "\"";
    for (size_t i = 0; i < s.size();) {
        const unsigned char c = (unsigned char)s[i];
        if (c < 0x80) {
            switch (c) {
                case '"': o += "\\\""; break; case '\\': o += "\\\\"; break; case '\n': o += "\\n"; break; case '\r': o += "\\r"; break; case '\t': o += "\\t"; break;
                case '\b': o += "\\b"; break; case '\f': o += "\\f"; break;
                default:
                    if (c < 0x20 || c == '<' || c == '>' || c == '&') { char b[8]; std::snprintf(b, sizeof b, "\\u%04x", c); o += b; } else o += (char)c;
            }
            ++i; continue;
        }
        int bad; const int n = oa::utf8_seq(s, i, &bad);
        if (!n) { o += "\\ufffd"; i += 1; continue; }                     // Go replaces every invalid byte on its own
        if (n == 3 && c == 0xE2 && (unsigned char)s[i + 1] == 0x80 && ((unsigned char)s[i + 2] == 0xA8 || (unsigned char)s[i + 2] == 0xA9)) o += (unsigned char)s[i + 2] == 0xA8 ? "\\u2028" : "\\u2029";
        else o.append(s, i, (size_t)n);
        i += (size_t)n;
    }
    return o + "\"";
}

struct ToolAction { std::string Name, Input; };
struct ToolPrompt {                                                       // tool.go:29-38, JSON keys question / thought / action{name,input} / observation / final_answer
    std::string Question, Thought; ToolAction Action; std::string Observation, FinalAnswer;

    // json.Marshal(toolPrompt), byte for byte: struct field order, no spaces (simple.go:497 sends this string as the next user message)
    std::string Marshal() const {
        return "{\"question\":" + GoJSONString(Question) + ",\"thought\":" + GoJSONString(Thought) + ",\"action\":{\"name\":" + GoJSONString(Action.Name) + ",\"input\":" +
               GoJSONString(Action.Input) + "},\"observation\":" + GoJSONString(Observation) + ",\"final_answer\":" + GoJSONString(FinalAnswer) + "}";
    }
    // json.Unmarshal([]byte(text), &toolPrompt) as encoding/json does it: the whole text must be valid JSON and an object; its keys are visited IN ORDER,
    // each matched to a field exactly or else case-insensitively (duplicates and differently-cased duplicates overwrite each other, the last one in
    // the document wins); unknown keys are ignored; JSON null leaves the field as it is; any other non-string value for a string field, or a non-object
    // `action`, is an UnmarshalTypeError — decoding goes on but Unmarshal returns that error: the caller's "not JSON, assume final answer" /
    // "Summarize…" branches.  Nothing is coerced.
    static bool Unmarshal(const std::string& text, ToolPrompt* out, std::string* err) {
        oa::Json d; std::string perr;
        if (!oa::parse_json(text, d, perr)) { *err = "invalid character: " + perr; return false; }
        if (d.t != oa::Json::Obj) { *err = "json: cannot unmarshal non-object into Go value of type tools.ToolPrompt"; return false; }
        ToolPrompt tp; std::string first_error;
        auto fold = [](const std::string& k) {          // encoding/json's foldName: simple case folding; U+017F and U+212A are the only non-ASCII runes that fold onto ASCII letters
            std::string o;
            for (size_t i = 0; i < k.size(); ++i) {
                const unsigned char c = (unsigned char)k[i];
                if (c == 0xC5 && i + 1 < k.size() && (unsigned char)k[i + 1] == 0xBF) { o += 's'; ++i; }
                else if (c == 0xE2 && i + 2 < k.size() && (unsigned char)k[i + 1] == 0x84 && (unsigned char)k[i + 2] == 0xAA) { o += 'k'; i += 2; }
                else o += (c >= 'A' && c <= 'Z') ? (char)(c - 'A' + 'a') : (char)c;
            }
            return o;
        };
        auto set_string = [&](std::string* dst, const oa::Json& v, const std::string& path) {
            if (v.t == oa::Json::Null) return;
            if (v.t != oa::Json::Str) { if (first_error.empty()) first_error = "json: cannot unmarshal a non-string into Go struct field ToolPrompt." + path + " of type string"; return; }
            *dst = v.s;
        };
        for (auto& kv : d.o) {
            const std::string f = fold(kv.first);
            if (f == "question") set_string(&tp.Question, kv.second, f);
            else if (f == "thought") set_string(&tp.Thought, kv.second, f);
            else if (f == "observation") set_string(&tp.Observation, kv.second, f);
            else if (f == "final_answer") set_string(&tp.FinalAnswer, kv.second, f);
            else if (f == "action") {
                if (kv.second.t == oa::Json::Null) continue;
                if (kv.second.t != oa::Json::Obj) { if (first_error.empty()) first_error = "json: cannot unmarshal non-object into Go struct field ToolPrompt.action"; continue; }
                for (auto& kv2 : kv.second.o) {
                    const std::string f2 = fold(kv2.first);
                    if (f2 == "name") set_string(&tp.Action.Name, kv2.second, "action.name");
                    else if (f2 == "input") set_string(&tp.Action.Input, kv2.second, "action.input");
                }
            }
        }
        if (!first_error.empty()) { *err = first_error; return false; }
        *out = tp; return true;
    }
};

// simple.go:624-657: a final_answer that is still the prompt's placeholder text does not count as an answer
inline bool isTemplateValue(const std::string& value) {
    if (value.size() < 10) return true;                                   // len() of a Go string is bytes
    static const char* patterns[] = {"<最终答案", "<final_answer", "<Final answer", "<最终回答", "<回答", "<答案", "使用 Markdown 格式", "使用Markdown格式", "换行符用 \\n 表示", "换行符用\\n表示"};
    for (const char* p : patterns) if (value.find(p) != std::string::npos) return true;
    return value.find('<') != std::string::npos && value.find('>') != std::string::npos;
}

// `count` = tokens of a message list under the ENGINE's tokenizer + chat template (oa_count_tokens).  The reference counts with tiktoken for OpenAI
// model names and returns 0 for anything else (tokens.go:61-66), which silently disables truncation for local models; an empty `count` keeps that.
using CountTokensFn = std::function<int(const std::vector<ChatCompletionMessage>&)>;
inline int NumTokensFromMessages(const std::vector<ChatCompletionMessage>& messages, const std::string& /*model*/, const CountTokensFn& count) { return count ? count(messages) : 0; }

// tokens.go:26-56: context windows by OpenAI model name (the reference's lookup data; lower-cased key, 4096 when unknown).  `engine_limit` (the local
// engine's max_seq_len) wins when given: the engine answers to any model name, and what bounds a request is ITS window
inline int GetTokenLimits(std::string model, int engine_limit = 0) {
    if (engine_limit > 0) return engine_limit;
    static const std::map<std::string, int> tokenLimitsPerModel = {
        {"code-davinci-002", 4096}, {"gpt-3.5-turbo-0301", 4096}, {"gpt-3.5-turbo-0613", 4096}, {"gpt-3.5-turbo-1106", 16385}, {"gpt-3.5-turbo-16k-0613", 16385}, {"gpt-3.5-turbo-16k", 16385},
        {"gpt-3.5-turbo-instruct", 4096}, {"gpt-3.5-turbo", 4096}, {"gpt-4-0314", 8192}, {"gpt-4-0613", 8192}, {"gpt-4-1106-preview", 128000}, {"gpt-4-32k-0314", 32768},
        {"gpt-4-32k-0613", 32768}, {"gpt-4-32k", 32768}, {"gpt-4-vision-preview", 128000}, {"gpt-4", 8192}, {"text-davinci-002", 4096}, {"text-davinci-003", 4096}, {"qwen-plus", 4096}};
    for (auto& ch : model) if (ch >= 'A' && ch <= 'Z') ch = (char)(ch - 'A' + 'a');
    auto it = tokenLimitsPerModel.find(model);
    return it == tokenLimitsPerModel.end() ? 4096 : it->second;
}
// tokens.go:110-125: nil (*ok = true, empty result, *none = true) when maxTokens alone exceeds the window; otherwise drop the oldest message after the
// first until prompt + maxTokens fit.  A single message that does not fit panics in the reference (messages[2:] out of range): *ok = false here.
inline std::vector<ChatCompletionMessage> ConstrictMessages(std::vector<ChatCompletionMessage> messages, const std::string& model, int maxTokens, const CountTokensFn& count,
                                                            bool* ok, bool* none = nullptr, int engine_limit = 0) {
    *ok = true; if (none) *none = false;
    const int tokenLimits = GetTokenLimits(model, engine_limit);
    if (maxTokens >= tokenLimits) { if (none) *none = true; return {}; }
    for (;;) {
        if (NumTokensFromMessages(messages, model, count) + maxTokens < tokenLimits) return messages;
        if (messages.size() < 2) { *ok = false; return {}; }
        messages.erase(messages.begin() + 1);
    }
}

inline std::string TrimSpace(const std::string& s);
// tokens.go:128-144: while the prompt does not fit, drop its first ceil(n/3) lines
inline std::string ConstrictPrompt(std::string prompt, const std::string& model, int tokenLimits, const CountTokensFn& count) {
    for (;;) {
        if (NumTokensFromMessages({ChatCompletionMessage{"", prompt}}, model, count) < tokenLimits) return prompt;
        std::vector<std::string> lines;
        size_t b = 0;
        for (;;) { const size_t e = prompt.find('\n', b); if (e == std::string::npos) { lines.push_back(prompt.substr(b)); break; } lines.push_back(prompt.substr(b, e - b)); b = e + 1; }
        const size_t drop = (lines.size() + 2) / 3;
        std::string next;
        for (size_t i = drop; i < lines.size(); ++i) { if (i > drop) next += '\n'; next += lines[i]; }
        prompt = next;
        if (TrimSpace(prompt).empty()) return "";
    }
}

// strings.TrimSpace: strips unicode.IsSpace code points — \t \n \v \f \r ' ' U+0085 U+00A0 and the Z category (U+1680, U+2000-200A, U+2028, U+2029,
// U+202F, U+205F, U+3000) — from both ends of a UTF-8 string
inline size_t go_space_at(const std::string& s, size_t i) {               // bytes of the white-space code point starting at i, 0 = not white space
    const unsigned char c = (unsigned char)s[i];
    if (c == ' ' || (c >= '\t' && c <= '\r')) return 1;
    auto at = [&](size_t k) { return i + k < s.size() ? (unsigned char)s[i + k] : 0u; };
    if (c == 0xC2 && (at(1) == 0x85 || at(1) == 0xA0)) return 2;
    if (c == 0xE1 && at(1) == 0x9A && at(2) == 0x80) return 3;                                            // U+1680
    if (c == 0xE2 && at(1) == 0x80 && ((at(2) >= 0x80 && at(2) <= 0x8A) || at(2) == 0xA8 || at(2) == 0xA9 || at(2) == 0xAF)) return 3;      // U+2000-200A, 2028, 2029, 202F
    if (c == 0xE2 && at(1) == 0x81 && at(2) == 0x9F) return 3;                                            // U+205F
    if (c == 0xE3 && at(1) == 0x80 && at(2) == 0x80) return 3;                                            // U+3000
    return 0;
}
inline std::string TrimSpace(const std::string& s) {
    size_t b = 0, e = s.size();
    while (b < e) { const size_t n = go_space_at(s, b); if (!n) break; b += n; }
    while (e > b) {
        size_t k = e - 1;
        while (k > b && ((unsigned char)s[k] & 0xC0) == 0x80) --k;        // start of the last code point
        const size_t n = go_space_at(s, k);
        if (!n || k + n != e) break;
        e = k;
    }
    return s.substr(b, e - b);
}

using ChatFn = std::function<std::string(const std::string& model, int maxTokens, const std::vector<ChatCompletionMessage>& prompts, Error* err)>;
using Tool = std::function<std::string(const std::string& input, std::string* err)>;      // *err non-empty = the tool failed (Go: (string, error))

struct AssistantResult { std::string Result; std::vector<ChatCompletionMessage> ChatHistory; Error Err; };

// -> (result, chatHistory, error).  Control flow and strings of simple.go:292-616:
//   first Chat -> Unmarshal into ToolPrompt; a reply that is not JSON is returned verbatim                                       (343-382)
//   loop (<= maxIterations, default 5): final_answer set, not a placeholder, and an observation present -> return it            (391-419)
//     action.name set -> run the tool; errors / unknown tools become the observation text                                       (421-481)
//     observation = ConstrictPrompt(observation, model, 1024); the whole ToolPrompt marshalled as a USER message                 (495-501)
//     Chat again; reply with final_answer -> return it; unparsable reply -> "Summarize all the chat history…" Chat              (515-600)
inline AssistantResult AssistantWithConfig(const std::string& model, const std::vector<ChatCompletionMessage>& prompts, int maxTokens, bool /*countTokens*/, bool /*verbose*/,
                                           int maxIterations, const ChatFn& chat, const std::map<std::string, Tool>& tools, const CountTokensFn& count = nullptr) {
    AssistantResult R;
    PerfStats& perf = GetPerfStats();
    struct Total { std::function<void()> done; ~Total() { done(); } } total{perf.TraceFunc("assistant_total")};                    // simple.go:296
    if (prompts.empty()) { R.Err = Error{0, "prompts cannot be empty"}; return R; }                                              // simple.go:312
    R.ChatHistory = prompts;
    auto timed_chat = [&](const char* op, std::string* resp) {
        perf.StartTimer(op);
        Error e; *resp = chat(model, maxTokens, R.ChatHistory, &e);
        perf.StopTimer(op);
        if (!e.ok()) { R.Err = Error{e.HTTPStatusCode, "chat completion error: " + e.Message}; return false; }
        return true;
    };
    std::string resp;
    if (!timed_chat("assistant_first_chat", &resp)) return R;                                                                      // simple.go:341-346
    R.ChatHistory.push_back({"assistant", resp});
    ToolPrompt tp; std::string perr;
    perf.StartTimer("assistant_parse_tool_prompt");                                                                                // simple.go:364-385
    const bool parsed = ToolPrompt::Unmarshal(resp, &tp, &perr);
    perf.StopTimer("assistant_parse_tool_prompt");
    if (!parsed) { R.Result = resp; return R; }                                                                                    // not JSON: assume final answer
    if (maxIterations <= 0) maxIterations = defaultMaxIterations;
    for (int iterations = 1;; ++iterations) {
        if (iterations > maxIterations) { R.Result = tp.FinalAnswer; return R; }
        if (!tp.FinalAnswer.empty() && !isTemplateValue(tp.FinalAnswer) && !tp.Observation.empty()) { R.Result = tp.FinalAnswer; return R; }
        if (tp.Action.Name.empty()) continue;                                    // nothing to run: the iteration budget ends the loop (simple.go:421)
        std::string observation;
        auto it = tools.find(tp.Action.Name);
        if (it != tools.end()) {
            perf.StartTimer("assistant_tool_" + tp.Action.Name);                                                                   // simple.go:440-475
            std::string terr; const std::string out = it->second(tp.Action.Input, &terr);
            perf.StopTimer("assistant_tool_" + tp.Action.Name);
            observation = terr.empty() ? TrimSpace(out) : "Tool " + tp.Action.Name + " failed with error " + terr + ". Considering refine the inputs for the tool.";
        } else observation = "Tool " + tp.Action.Name + " is not available. Considering switch to other supported tools.";
        perf.StartTimer("assistant_construct_message");                                                                            // simple.go:491-507
        tp.Observation = ConstrictPrompt(observation, model, 1024, count);
        R.ChatHistory.push_back({"user", tp.Marshal()});
        perf.StopTimer("assistant_construct_message");
        if (!timed_chat("assistant_intermediate_chat", &resp)) return R;                                                           // simple.go:513-518
        R.ChatHistory.push_back({"assistant", resp});
        ToolPrompt next;
        perf.StartTimer("assistant_parse_intermediate");                                                                           // simple.go:541-603
        const bool ok = ToolPrompt::Unmarshal(resp, &next, &perr);
        perf.StopTimer("assistant_parse_intermediate");
        if (!ok) {
            R.ChatHistory.push_back({"user", "Summarize all the chat history and respond to original question with final answer"});
            if (!timed_chat("assistant_summarize", &resp)) return R;                                                               // simple.go:564-569
            R.Result = resp; return R;
        }
        tp = next;
        if (!tp.FinalAnswer.empty()) { R.Result = tp.FinalAnswer; return R; }
    }
}

}  // namespace opsagent
