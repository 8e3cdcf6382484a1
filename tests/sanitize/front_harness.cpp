// Race / memory-error harness for the native HTTP front: csrc/http_server.cpp compiled together with a STUB engine (no CUDA) under
// -fsanitize=thread or -fsanitize=address,undefined, hammered by concurrent in-process HTTP clients — well-formed chats, function calls, bad JSON,
// clients that hang up mid-request — then stopped while connections are still open.  tests/test_native_front.py builds and runs it; exit 0 = the
// wire checks passed and the sanitizer reported nothing (TSAN exits 66, ASAN aborts).
#include <arpa/inet.h>
#include <netinet/in.h>
#include <sys/socket.h>
#include <unistd.h>

#include <atomic>
#include <chrono>
#include <condition_variable>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <deque>
#include <mutex>
#include <random>
#include <string>
#include <thread>
#include <unordered_map>
#include <vector>

#include "../../include/opsagent_b200.h"

// ---- stub engine: answers after a short random delay from its own thread, like the scheduler thread of the real one ---------------------
struct Job { uint64_t ticket; std::string text; uint32_t flags; std::string functions; int prompt_tokens; };
struct oa_engine {
    int id = 0;
    std::mutex mu; std::condition_variable cv_work, cv_done;
    std::deque<Job> queue; std::unordered_map<uint64_t, Job> done;
    uint64_t next = 1; bool stop = false; long served = 0, cancelled = 0; std::unordered_map<uint64_t, bool> cancel_req;
    std::thread worker;
    void run() {
        std::mt19937 rng(1234u + (unsigned)id);
        for (;;) {
            Job j;
            {
                std::unique_lock<std::mutex> lk(mu);
                cv_work.wait(lk, [&] { return stop || !queue.empty(); });
                if (queue.empty()) return;
                j = std::move(queue.front()); queue.pop_front();
            }
            std::this_thread::sleep_for(std::chrono::microseconds(rng() % 1500));
            if (j.text == "slow please") {          // a long completion: runs until it is cancelled (or 3 s)
                bool dropped = false;
                for (int k = 0; k < 60 && !dropped; ++k) {
                    std::this_thread::sleep_for(std::chrono::milliseconds(50));
                    std::lock_guard<std::mutex> lk(mu); dropped = cancel_req.count(j.ticket) != 0;
                }
                if (dropped) { std::lock_guard<std::mutex> lk(mu); cancel_req.erase(j.ticket); ++cancelled; continue; }
            }
            if (j.flags & OA_FLAG_JSON_FUNCTION) {
                const std::string f = j.functions.substr(0, j.functions.find(','));
                const size_t c = f.find(':');
                j.text = "{\"name\":\"" + f.substr(0, c) + "\",\"arguments\":{\"" + f.substr(c + 1) + "\":\"echo " + std::to_string(j.text.size()) + "\"}}";
            } else j.text = "replica " + std::to_string(id) + " got " + std::to_string(j.text.size()) + " bytes: " + j.text.substr(0, 24) + "\xe4\xb8\n";      // ends mid-character on purpose
            { std::lock_guard<std::mutex> lk(mu); ++served; done[j.ticket] = std::move(j); }
            cv_done.notify_all();
        }
    }
};
extern "C" {
int oa_model_info(oa_engine*, char* buf, size_t n) { std::snprintf(buf, n, "{\"model\": \"stub-model\"}"); return 0; }
int oa_engine_stats(oa_engine* e, char* buf, size_t n) { std::lock_guard<std::mutex> lk(e->mu); std::snprintf(buf, n, "{\"completed\": %ld, \"busy_ms\": 1.5}", e->served); return 0; }
int oa_chat_submit_ex(oa_engine* e, const oa_chat_req* r, uint64_t* ticket, char* eb, size_t ec) {
    if (r->n_msgs <= 0) { std::snprintf(eb, ec, "no messages"); return 400; }
    Job j; j.text = r->msgs[r->n_msgs - 1].content; j.flags = r->flags; j.functions = r->functions ? r->functions : ""; j.prompt_tokens = (int)j.text.size();
    if (j.text == "please fail") { std::snprintf(eb, ec, "engine says no"); return 500; }
    { std::lock_guard<std::mutex> lk(e->mu); j.ticket = *ticket = e->next++; e->queue.push_back(std::move(j)); }
    e->cv_work.notify_one();
    return 0;
}
int oa_chat_cancel(oa_engine* e, uint64_t ticket) { std::lock_guard<std::mutex> lk(e->mu); e->cancel_req[ticket] = true; return 0; }
int oa_chat_wait_ex(oa_engine* e, uint64_t ticket, int32_t timeout_ms, oa_chat_resp* out, char*, size_t) {
    std::unique_lock<std::mutex> lk(e->mu);
    if (timeout_ms < 0) e->cv_done.wait(lk, [&] { return e->done.count(ticket) != 0; });
    else if (!e->cv_done.wait_for(lk, std::chrono::milliseconds(timeout_ms), [&] { return e->done.count(ticket) != 0; })) return OA_ERR_TIMEOUT;
    Job j = std::move(e->done[ticket]); e->done.erase(ticket);
    out->content = (char*)std::malloc(j.text.size() + 1); std::memcpy(out->content, j.text.c_str(), j.text.size() + 1);
    out->content_len = (int32_t)j.text.size(); out->prompt_tokens = j.prompt_tokens; out->completion_tokens = 7; out->finish_reason = 0; out->token_ids = nullptr;
    return 0;
}
void oa_free_resp(oa_chat_resp* r) { std::free(r->content); r->content = nullptr; }
}

// ---- a minimal blocking HTTP client ---------------------------------------------------------------------------------------------------------
static int dial(int port) {
    const int s = socket(AF_INET, SOCK_STREAM, 0);
    sockaddr_in a{}; a.sin_family = AF_INET; a.sin_port = htons((uint16_t)port); inet_pton(AF_INET, "127.0.0.1", &a.sin_addr);
    if (connect(s, (sockaddr*)&a, sizeof a) != 0) { close(s); return -1; }
    return s;
}
static bool roundtrip(int s, const std::string& method, const std::string& path, const std::string& body, bool auth, int* status, std::string* resp) {
    std::string rq = method + " " + path + " HTTP/1.1\r\nHost: x\r\n" + (auth ? "Authorization: Bearer k\r\n" : "") + "Content-Length: " + std::to_string(body.size()) + "\r\n\r\n" + body;
    size_t off = 0;
    while (off < rq.size()) { const ssize_t n = send(s, rq.data() + off, rq.size() - off, MSG_NOSIGNAL); if (n <= 0) return false; off += (size_t)n; }
    std::string buf; char tmp[4096]; size_t hdr_end;
    while ((hdr_end = buf.find("\r\n\r\n")) == std::string::npos) { const ssize_t n = recv(s, tmp, sizeof tmp, 0); if (n <= 0) return false; buf.append(tmp, (size_t)n); }
    *status = std::atoi(buf.c_str() + 9);
    const size_t cl = buf.find("Content-Length: ");
    const size_t need = hdr_end + 4 + (size_t)std::atol(buf.c_str() + cl + 16);
    while (buf.size() < need) { const ssize_t n = recv(s, tmp, sizeof tmp, 0); if (n <= 0) return false; buf.append(tmp, (size_t)n); }
    *resp = buf.substr(hdr_end + 4);
    return true;
}

static std::atomic<int> failures{0};
#define CHECK(c) do { if (!(c)) { std::fprintf(stderr, "CHECK failed line %d: %s\n", __LINE__, #c); failures.fetch_add(1); } } while (0)

int main() {
    constexpr int N_ENGINES = 3, N_CLIENTS = 24, N_ROUNDS = 40;
    std::vector<oa_engine*> engines;
    for (int i = 0; i < N_ENGINES; ++i) { auto* e = new oa_engine; e->id = i; e->worker = std::thread([e] { e->run(); }); engines.push_back(e); }
    oa_http* front = nullptr;
    if (oa_http_start(engines.data(), N_ENGINES, "{\"port\": 0, \"require_key\": 1, \"max_inflight\": 4, \"tool_steps\": 1}", &front) != 0) { std::fprintf(stderr, "start: %s\n", oa_http_last_error()); return 2; }
    const int port = oa_http_port(front);
    std::atomic<int> n200{0}, n429{0};
    std::vector<std::thread> clients;
    for (int c = 0; c < N_CLIENTS; ++c) clients.emplace_back([&, c] {
        std::mt19937 rng(99u + (unsigned)c);
        int s = dial(port);
        const std::string sys = "{\"role\": \"system\", \"content\": \"conversation " + std::to_string(c) + " \\u4e2d\\ud83d\\ude00\"}, {\"role\": \"user\", \"content\": \"first question of " + std::to_string(c) + "\"}";
        for (int r = 0; r < N_ROUNDS; ++r) {
            if (s < 0) s = dial(port);
            CHECK(s >= 0);
            int st = 0; std::string resp;
            switch (rng() % 8) {
                case 0: CHECK(roundtrip(s, "POST", "/v1/chat/completions", "{bad json", true, &st, &resp) && st == 400); break;
                case 1: CHECK(roundtrip(s, "GET", "/v1/models", "", true, &st, &resp) && st == 200 && resp.find("stub-model") != std::string::npos); break;
                case 2: CHECK(roundtrip(s, "GET", "/api/perf/stats", "", true, &st, &resp) && st == 200 && resp.find("callCounts") != std::string::npos); break;
                case 3: CHECK(roundtrip(s, "POST", "/v1/chat/completions", "{\"messages\": []}", false, &st, &resp) && st == 401); break;
                case 4: {   // hang up in the middle of a body; the server must drop the connection and carry on
                    const std::string half = "POST /v1/chat/completions HTTP/1.1\r\nAuthorization: Bearer k\r\nContent-Length: 500\r\n\r\n{\"messages\": [";
                    send(s, half.data(), half.size(), MSG_NOSIGNAL); close(s); s = -1; break;
                }
                case 5: {   // function calling
                    const std::string body = "{\"model\": \"m\", \"messages\": [" + sys + "], \"tools\": [{\"type\": \"function\", \"function\": {\"name\": \"kubectl\", \"parameters\": {\"properties\": {\"command\": {}}}}}]}";
                    CHECK(roundtrip(s, "POST", "/v1/chat/completions", body, true, &st, &resp));
                    if (st == 200) { n200.fetch_add(1); CHECK(resp.find("\"tool_calls\"") != std::string::npos && resp.find("kubectl") != std::string::npos && resp.find("{\\\"command\\\": \\\"echo ") != std::string::npos); }
                    else { CHECK(st == 429); n429.fetch_add(1); }
                    break;
                }
                case 6: CHECK(roundtrip(s, "POST", "/v1/chat/completions", "{\"messages\": [{\"role\": \"user\", \"content\": \"please fail\"}]}", true, &st, &resp) && (st == 500 || st == 429)); break;
                default: {
                    const std::string body = "{\"model\": \"m\", \"max_tokens\": 16, \"messages\": [" + sys + ", {\"role\": \"user\", \"content\": \"step " + std::to_string(r) + "\"}]}";
                    CHECK(roundtrip(s, "POST", "/v1/chat/completions", body, true, &st, &resp));
                    if (st == 200) { n200.fetch_add(1); CHECK(resp.find("\"finish_reason\": \"stop\"") != std::string::npos && resp.find("got ") != std::string::npos && resp.find("\xef\xbf\xbd") != std::string::npos); }
                    else { CHECK(st == 429); n429.fetch_add(1); }
                }
            }
        }
        if (s >= 0 && c % 2 == 0) close(s);          // odd clients leave their keep-alive connection open: stop() has to deal with it
    });
    clients.emplace_back([&] {          // abandons three long completions: the front must notice the hang-up and cancel them in the engine
        for (int k = 0; k < 3; ++k) {
            const int s = dial(port);
            const std::string body = "{\"messages\": [{\"role\": \"system\", \"content\": \"abandoner " + std::to_string(k) + "\"}, {\"role\": \"user\", \"content\": \"slow please\"}]}";
            const std::string rq = "POST /v1/chat/completions HTTP/1.1\r\nAuthorization: Bearer k\r\nContent-Length: " + std::to_string(body.size()) + "\r\n\r\n" + body;
            send(s, rq.data(), rq.size(), MSG_NOSIGNAL);
            std::this_thread::sleep_for(std::chrono::milliseconds(300));
            close(s);
        }
    });
    for (auto& t : clients) t.join();
    std::this_thread::sleep_for(std::chrono::milliseconds(800));          // one wait slice + one stub poll tick
    char buf[1 << 14];
    CHECK(oa_http_stats(front, buf, sizeof buf) == 0);
    std::string st(buf);
    CHECK(st.find("\"replicas\": 3") != std::string::npos);
    long cancelled = 0;
    for (auto* e : engines) { std::lock_guard<std::mutex> lk(e->mu); cancelled += e->cancelled; }
    CHECK(cancelled >= 1);                              // (a request rejected with 429 never reached an engine)
    CHECK(st.find("\"cancelled\": 0") == std::string::npos);
    long served = 0;
    for (auto* e : engines) { std::lock_guard<std::mutex> lk(e->mu); served += e->served; }
    CHECK(served == n200.load());
    CHECK(n200.load() > 0);
    oa_http_stop(front);
    for (auto* e : engines) { { std::lock_guard<std::mutex> lk(e->mu); e->stop = true; } e->cv_work.notify_all(); e->worker.join(); delete e; }
    std::printf("ok: %d completions, %d rejected with 429, %d check failures\n", n200.load(), n429.load(), failures.load());
    return failures.load() ? 1 : 0;
}
