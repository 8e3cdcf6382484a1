// Runs opsagent::AssistantWithConfig (opsagent_b200/host/assistants.hpp) on scripted scenarios: argv[1] = JSON file
//   [{"prompts": [[role, content], ...], "replies": [text, ...], "tools": {name: {"outputs": [...]} | {"error": msg}}, "maxIterations": n, "count_div": d}, ...]
// and prints one JSON line per scenario: {"result", "error", "history": [[role, content], ...], "chat_calls"}.  A reply "!ERR:<msg>" is a failing Chat
// (status 400); running out of replies is a failing Chat (status 500).  count_div > 0: a message list costs 4 + bytes/count_div tokens per message.
// tests/test_host_cpp.py feeds the same scenarios to the Python mirror (opsagent_b200/assistants.py) and compares.
#include <cstdio>
#include <fstream>
#include <sstream>

#include "opsagent_b200/host/assistants.hpp"

using namespace opsagent;

int main(int argc, char** argv) {
    if (argc < 2) return 2;
    std::ifstream f(argv[1], std::ios::binary); std::stringstream ss; ss << f.rdbuf();
    oa::Json root; std::string err;
    if (!oa::parse_json(ss.str(), root, err) || root.t != oa::Json::Arr) { std::fprintf(stderr, "bad scenario file: %s\n", err.c_str()); return 2; }
    for (const oa::Json& sc : root.a) {
        if (const oa::Json* texts = sc.get("unmarshal")) {          // {"unmarshal": [text, ...]} -> one line: [marshal(unmarshal(text)) or null, ...]
            std::string line = "[";
            for (size_t i = 0; i < texts->a.size(); ++i) {
                ToolPrompt tp; std::string e;
                line += (i ? ", " : "") + (ToolPrompt::Unmarshal(texts->a[i].s, &tp, &e) ? oa::jstr(tp.Marshal()) : std::string("null"));
            }
            std::printf("%s]\n", line.c_str());
            continue;
        }
        std::vector<ChatCompletionMessage> prompts;
        for (const oa::Json& m : sc.get("prompts")->a) prompts.push_back({m.a[0].s, m.a[1].s});
        std::vector<std::string> replies;
        for (const oa::Json& r : sc.get("replies")->a) replies.push_back(r.s);
        size_t next = 0; int calls = 0;
        ChatFn chat = [&](const std::string&, int, const std::vector<ChatCompletionMessage>&, Error* e) -> std::string {
            ++calls;
            if (next >= replies.size()) { *e = Error{500, "no more replies"}; return ""; }
            const std::string r = replies[next++];
            if (r.rfind("!ERR:", 0) == 0) { *e = Error{400, r.substr(5)}; return ""; }
            *e = Error{}; return r;
        };
        std::map<std::string, Tool> tools;
        std::map<std::string, size_t> cursor;
        if (const oa::Json* t = sc.get("tools")) for (auto& kv : t->o) {
            const oa::Json spec = kv.second; const std::string name = kv.first;
            tools[name] = [spec, name, &cursor](const std::string&, std::string* terr) -> std::string {
                if (const oa::Json* e = spec.get("error")) { *terr = e->s; return ""; }
                const auto& outs = spec.get("outputs")->a;
                return outs[cursor[name]++ % outs.size()].s;
            };
        }
        const int div = (int)(sc.get("count_div") ? sc.get("count_div")->n : 0);
        CountTokensFn count;
        if (div > 0) count = [div](const std::vector<ChatCompletionMessage>& ms) { int n = 0; for (auto& m : ms) n += 4 + (int)m.Content.size() / div; return n; };
        GetPerfStats().Reset();
        const AssistantResult R = AssistantWithConfig("m", prompts, 256, false, false, (int)sc.get("maxIterations")->n, chat, tools, count);
        std::string line = "{\"result\": " + oa::jstr(R.Result) + ", \"error\": " + oa::jstr(R.Err.Message) + ", \"chat_calls\": " + std::to_string(calls) + ", \"callCounts\": {";
        { bool first = true; for (auto& kv : GetPerfStats().GetStats().callCounts) { line += (first ? "" : ", ") + oa::jstr(kv.first) + ": " + std::to_string(kv.second); first = false; } }
        line += "}, \"history\": [";
        for (size_t i = 0; i < R.ChatHistory.size(); ++i) line += (i ? ", [" : "[") + oa::jstr(R.ChatHistory[i].Role) + ", " + oa::jstr(R.ChatHistory[i].Content) + "]";
        std::printf("%s]}\n", line.c_str());
    }
    return 0;
}
