// localcuda_client.hpp — C++ host-side mirror of the reference's LLM seam above the C ABI.
//
// The reference is Go (no toolchain in this image), so the host side above include/opsagent_b200.h is written in
// C++ with the same names, argument meaning and error behaviour as
//     type OpenAIClient struct { *openai.Client; Retries int; Backoff time.Duration }      reference pkg/llms/openai.go:29-35
//     func NewOpenAIClient(apiKey, baseURL string) (*OpenAIClient, error)                    reference pkg/llms/openai.go:38-63
//     func (c *OpenAIClient) Chat(model string, maxTokens int, prompts []ChatCompletionMessage) (string, error)   :69-104
// integration/go/localcuda.go is the cgo file a maintainer drops into pkg/llms; this header is its compiled twin and
// is what tests/test_host_cpp.py builds and runs.
#pragma once
#include <chrono>
#include <functional>
#include <string>
#include <thread>
#include <vector>

#include "../../include/opsagent_b200.h"

namespace opsagent {

struct ChatCompletionMessage { std::string Role, Content; };   // go-openai ChatCompletionMessage{Role, Content}

struct Error {                                                  // nil error <=> ok()
    int HTTPStatusCode = 0; std::string Message;
    bool ok() const { return HTTPStatusCode == 0 && Message.empty(); }
};

class LocalCUDAClient {
public:
    int Retries = 5;                                            // openai.go:58
    std::chrono::milliseconds Backoff{1000};                    // openai.go:59
    std::function<void(std::chrono::milliseconds)> Sleep = [](std::chrono::milliseconds d) { std::this_thread::sleep_for(d); };

    // NewOpenAIClient: empty key -> "OPENAI_API_KEY is not set" (openai.go:40-42).  The engine handle is process-global
    // (the reference builds a client per request, pkg/assistants/simple.go:316), so it is passed in, not owned.
    static Error New(const std::string& apiKey, oa_engine* engine, LocalCUDAClient* out) {
        if (apiKey.empty()) return Error{0, "OPENAI_API_KEY is not set"};
        out->engine_ = engine;
        return Error{};
    }

    // Chat: one ReAct step.  Returns Choices[0].Message.Content (openai.go:82).
    std::string Chat(const std::string& model, int maxTokens, const std::vector<ChatCompletionMessage>& prompts, Error* err) {
        std::vector<oa_msg> msgs(prompts.size());
        for (size_t i = 0; i < prompts.size(); ++i) { msgs[i].role = prompts[i].Role.c_str(); msgs[i].content = prompts[i].Content.c_str(); }
        oa_chat_req req{};
        req.model = model.c_str(); req.msgs = msgs.data(); req.n_msgs = (int32_t)msgs.size(); req.max_tokens = maxTokens;
        req.temperature = 1.401298464324817e-45f;               // math.SmallestNonzeroFloat32 (openai.go:73)
        auto backoff = Backoff;
        for (int attempt = 0; attempt < Retries; ++attempt) {
            oa_chat_resp resp{};
            const int rc = engine_ ? oa_chat_complete(engine_, &req, &resp) : OA_ERR_INTERNAL;
            if (rc == OA_OK) {
                std::string content(resp.content, (size_t)resp.content_len);
                oa_free_resp(&resp);
                *err = Error{};
                return content;
            }
            const std::string msg = engine_ ? oa_last_error() : "engine handle is null";
            switch (rc) {
                case 401: *err = Error{rc, msg}; return "";                       // openai.go:88-90
                case 429: case 500: Sleep(backoff); backoff *= 2; continue;       // openai.go:91-94
                default: *err = Error{rc, msg}; return "";                        // openai.go:95-97
            }
        }
        *err = Error{0, "OpenAI request throttled after retrying " + std::to_string(Retries) + " times"};   // openai.go:103
        return "";
    }

private:
    oa_engine* engine_ = nullptr;
};

}  // namespace opsagent
