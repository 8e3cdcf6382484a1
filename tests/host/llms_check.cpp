// GetTokenLimits / ConstrictMessages / ConstrictPrompt / TrimSpace / isTemplateValue / GoJSONString of host/assistants.hpp on fixed cases (the
// reference-held ones from pkg/llms/tokens_test.go:20-51 and the cases tests/test_host_logic.py runs against the Python mirrors).
#include <cstdio>

#include "opsagent_b200/host/assistants.hpp"

using namespace opsagent;
#define CHECK(c) do { if (!(c)) { std::printf("FAIL line %d: %s\n", __LINE__, #c); return 1; } } while (0)

int main() {
    CHECK(GetTokenLimits("gpt-3.5-turbo-0613") == 4096 && GetTokenLimits("gpt-4") == 8192);               // the reference's own golden cases
    CHECK(GetTokenLimits("GPT-4-32K") == 32768 && GetTokenLimits("llama-3-8b") == 4096 && GetTokenLimits("gpt-4", 16384) == 16384);
    CountTokensFn count = [](const std::vector<ChatCompletionMessage>& ms) { int n = 3; for (auto& m : ms) n += 3 + (int)m.Content.size(); return n; };
    std::vector<ChatCompletionMessage> msgs = {{"system", std::string(100, 's')}, {"user", std::string(300, 'a')}, {"assistant", std::string(300, 'b')}, {"user", std::string(300, 'c')}};
    bool ok = false, none = false;
    CHECK(ConstrictMessages(msgs, "m", 5000, count, &ok, &none).empty() && ok && none);
    CHECK(ConstrictMessages(msgs, "m", 100, count, &ok, &none, 2000).size() == 4 && ok && !none);
    auto got = ConstrictMessages(msgs, "m", 100, count, &ok, &none, 600);
    CHECK(ok && got.size() == 2 && got[0].Content[0] == 's' && got[1].Content[0] == 'c');
    ConstrictMessages({{"system", std::string(1000, 's')}}, "m", 100, count, &ok, &none, 600);
    CHECK(!ok);                                                                                             // the reference panics here
    CHECK(TrimSpace(" \t\n\xe2\x80\xa8\xe3\x80\x80\xc2\xa0\xc2\x85 x y \xe2\x80\x8a\r") == "x y");
    CHECK(TrimSpace("\x1c x \x1f") == "\x1c x \x1f" && TrimSpace("\xe2\x80\x8b x") == "\xe2\x80\x8b x" && TrimSpace("") == "" && TrimSpace(" \n ") == "");
    CHECK(GoJSONString("a\b\f\x0b\x7f<>&\xe2\x80\xa8\xff") == "\"a\\b\\f\\u000b\x7f\\u003c\\u003e\\u0026\\u2028\\ufffd\"");
    CHECK(isTemplateValue("short") && isTemplateValue("<final_answer goes here>") && !isTemplateValue("The pod is crashing because of OOM"));
    // ConstrictPrompt: 1 token per 4 bytes, limit 30: drops the first third of the lines until it fits (tokens.go:128-144)
    CountTokensFn c4 = [](const std::vector<ChatCompletionMessage>& ms) { int n = 0; for (auto& m : ms) n += (int)m.Content.size() / 4; return n; };
    std::string table; for (int i = 0; i < 30; ++i) table += "pod-" + std::to_string(i) + " Running\n";
    const std::string cut = ConstrictPrompt(table, "m", 30, c4);
    CHECK(!cut.empty() && cut.size() / 4 < 30 && table.size() >= cut.size() && table.compare(table.size() - cut.size(), cut.size(), cut) == 0);      // a suffix of the input
    CHECK(ConstrictPrompt("  \n \n", "m", 0, c4) == "");
    {   // Unmarshal visits keys in document order like encoding/json: folded matches overwrite each other, null keeps, a wrong type anywhere is an error
        ToolPrompt tp; std::string e;
        CHECK(ToolPrompt::Unmarshal("{\"que\xc5\xbftion\":\"a\",\"THOUGHT\":\"t\",\"thought\":null,\"action\":{\"Name\":\"x\"},\"ACTION\":{\"input\":\"y\"}}", &tp, &e));
        CHECK(tp.Question == "a" && tp.Thought == "t" && tp.Action.Name == "x" && tp.Action.Input == "y");
        CHECK(ToolPrompt::Unmarshal("{\"thought\":\"b\",\"Thought\":\"a\"}", &tp, &e) && tp.Thought == "a");
        CHECK(!ToolPrompt::Unmarshal("{\"Thought\": 7, \"Thought\": \"ok\"}", &tp, &e) && !ToolPrompt::Unmarshal("{\"action\": \"kubectl\"}", &tp, &e) && !ToolPrompt::Unmarshal("[1]", &tp, &e));
    }
    std::printf("ok\n");
    return 0;
}
