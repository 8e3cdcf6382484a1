// http_server.cpp — native OpenAI-compatible front: POST /v1/chat/completions (incl. `tools` -> grammar-forced `tool_calls`), GET /v1/models,
// GET /api/perf/stats, in front of ONE OR MORE engines of this process (data-parallel replicas), in C++ on top of the C ABI.
//
// The reference serves requests from one Go process, one goroutine per request (pkg/api/router.go:95); its LLM calls leave through go-openai
// (pkg/llms/openai.go:70-82) or openai-go (swarm flows, pkg/workflows/swarm.go:83).  This is the endpoint those clients can be pointed at
// (`baseUrl` of POST /api/execute, OPENAI_API_BASE) — the same wire behaviour as opsagent_b200/http_front.py (its Python twin, kept as the
// executable specification the tests compare against), without an interpreter between the socket and the engines:
//   * one OS thread per connection (HTTP/1.1 keep-alive), blocking in oa_chat_complete; batching happens inside the engines;
//   * N engines: a conversation is keyed by its first two messages and sticks to the replica that holds its prefix KV pages (the ReAct loop
//     resends its history, pkg/assistants/simple.go:498-501); new conversations go to the replica with the fewest requests in flight; a
//     replica over `max_inflight` answers 429, which the caller's retry loop backs off on (openai.go:91-94)  [opsagent_b200/router.py];
//   * status codes are the ones Chat switches on: 400 fail fast, 401, 429 / 500 retry (openai.go:85-101); errors use the OpenAI error body.
#ifndef _GNU_SOURCE
#define _GNU_SOURCE      // POLLRDHUP
#endif
#include <arpa/inet.h>
#include <netinet/in.h>
#include <netinet/tcp.h>
#include <poll.h>
#include <sys/socket.h>
#include <unistd.h>

#include <algorithm>
#include <atomic>
#include <chrono>
#include <cmath>
#include <cstdio>
#include <cstring>
#include <memory>
#include <ctime>
#include <map>
#include <mutex>
#include <string>
#include <thread>
#include <unordered_map>
#include <utility>
#include <vector>

#include "../../include/opsagent_b200.h"
#include "config.hpp"
#include "json_dom.hpp"

namespace oa {

// ---- the front -------------------------------------------------------------------------------------------------------------------------
class HttpFront {
public:
    HttpFront(std::vector<oa_engine*> engines, const JsonFlat& opt) : engines_(std::move(engines)), inflight_(engines_.size()), routed_(engines_.size()) {
        require_key_ = opt.i("require_key", 1) != 0; api_key_ = opt.s("api_key", ""); tool_steps_ = (int)opt.i("tool_steps", 3);
        max_inflight_ = (int)opt.i("max_inflight", 256); max_conn_ = (int)opt.i("max_connections", 8192); max_body_ = (size_t)opt.i("max_body_bytes", 64 << 20);
        idle_timeout_ms_ = (int)std::max<long long>(200, opt.i("idle_timeout_s", 120) * 1000);
        for (auto& a : inflight_) a.store(0);
        for (auto& a : routed_) a.store(0);
        const std::string host = opt.s("host", "127.0.0.1");
        lsock_ = socket(AF_INET, SOCK_STREAM, 0);
        if (lsock_ < 0) throw std::runtime_error("socket() failed");
        int one = 1; setsockopt(lsock_, SOL_SOCKET, SO_REUSEADDR, &one, sizeof one);
        sockaddr_in a{}; a.sin_family = AF_INET; a.sin_port = htons((uint16_t)opt.i("port", 0));
        if (inet_pton(AF_INET, host.c_str(), &a.sin_addr) != 1) { close(lsock_); throw std::runtime_error("bad listen address " + host); }
        if (bind(lsock_, reinterpret_cast<sockaddr*>(&a), sizeof a) != 0 || listen(lsock_, 4096) != 0) { close(lsock_); throw std::runtime_error("bind/listen failed on " + host); }
        socklen_t len = sizeof a; getsockname(lsock_, reinterpret_cast<sockaddr*>(&a), &len); port_ = ntohs(a.sin_port);
        if (!engines_.empty()) { char buf[2048]; if (oa_model_info(engines_[0], buf, sizeof buf) == 0) { Json j; std::string e; if (parse_json(buf, j, e)) model_ = j.str("model"); } }
        acceptor_ = std::thread([this] { accept_loop(); });
    }
    ~HttpFront() { stop(); }
    // stop accepting and wait for the connection threads (they notice stop_ at their next poll tick; one blocked in a long completion gets 30 s).
    // false: a connection thread is still inside the engine — the caller must not free this object
    bool stop() {
        if (!stop_.exchange(true)) {
            shutdown(lsock_, SHUT_RDWR);                       // wakes the acceptor's poll; the descriptor is closed only after the thread is gone
            if (acceptor_.joinable()) acceptor_.join();
            close(lsock_);
        }
        const auto t0 = std::chrono::steady_clock::now();
        while (n_conn_.load() > 0 && std::chrono::duration<double>(std::chrono::steady_clock::now() - t0).count() < 30.0) std::this_thread::sleep_for(std::chrono::milliseconds(5));
        return n_conn_.load() == 0;
    }
    int port() const { return port_; }
    std::string stats_json() {
        std::string s = "{\"replicas\": " + std::to_string(engines_.size()) + ", \"requests\": " + std::to_string(n_req_.load()) + ", \"chat_completions\": " + std::to_string(n_chat_.load()) +
                        ", \"rejected_429\": " + std::to_string(n_429_.load()) + ", \"sticky_hits\": " + std::to_string(n_sticky_.load()) + ", \"cancelled\": " + std::to_string(n_cancelled_.load()) + ", \"connections\": " + std::to_string(n_conn_.load()) + ", \"routed\": [";
        for (size_t i = 0; i < routed_.size(); ++i) s += (i ? ", " : "") + std::to_string(routed_[i].load());
        s += "], \"inflight\": [";
        for (size_t i = 0; i < inflight_.size(); ++i) s += (i ? ", " : "") + std::to_string(inflight_[i].load());
        s += "], \"engines\": [";
        for (size_t i = 0; i < engines_.size(); ++i) { char buf[4096]; buf[0] = 0; if (oa_engine_stats(engines_[i], buf, sizeof buf) != 0) std::snprintf(buf, sizeof buf, "null"); s += (i ? ", " : "") + std::string(buf); }
        return s + "]}";
    }

private:
    struct Request { std::string method, path, auth, body; bool keep_alive = true; int fd = -1; };
    void accept_loop() {
        while (!stop_.load()) {
            pollfd p{lsock_, POLLIN, 0};
            if (poll(&p, 1, 200) <= 0) continue;
            const int c = accept(lsock_, nullptr, nullptr);
            if (c < 0) continue;
            if (n_conn_.load() >= max_conn_) { respond(c, 429, error_body(429, "too many connections"), false); close(c); continue; }
            int one = 1; setsockopt(c, IPPROTO_TCP, TCP_NODELAY, &one, sizeof one);
            n_conn_.fetch_add(1);
            try { std::thread([this, c] { serve_connection(c); close(c); n_conn_.fetch_sub(1); }).detach(); }
            catch (const std::exception&) { n_conn_.fetch_sub(1); respond(c, 429, error_body(429, "cannot start a connection thread"), false); close(c); }
        }
    }
    // read one request; false = connection closed / malformed (a reply has been sent where one was possible)
    bool read_request(int c, std::string& buf, Request& rq) {
        size_t hdr_end;
        while ((hdr_end = buf.find("\r\n\r\n")) == std::string::npos) {
            if (buf.size() > (64u << 10)) { respond(c, 400, error_body(400, "header section too large"), false); return false; }
            if (!fill(c, buf)) return false;
        }
        const std::string head = buf.substr(0, hdr_end);
        size_t le = head.find("\r\n");
        const std::string line = head.substr(0, le);
        const size_t s1 = line.find(' '), s2 = line.rfind(' ');
        if (s1 == std::string::npos || s2 <= s1) { respond(c, 400, error_body(400, "malformed request line"), false); return false; }
        rq.method = line.substr(0, s1); rq.path = line.substr(s1 + 1, s2 - s1 - 1);
        const size_t q = rq.path.find('?'); if (q != std::string::npos) rq.path.resize(q);
        while (rq.path.size() > 1 && rq.path.back() == '/') rq.path.pop_back();
        rq.keep_alive = line.substr(s2 + 1) != "HTTP/1.0";
        long long clen = 0; rq.auth.clear();
        size_t pos = le == std::string::npos ? head.size() : le + 2;
        while (pos < head.size()) {
            size_t eol = head.find("\r\n", pos); if (eol == std::string::npos) eol = head.size();
            const std::string h = head.substr(pos, eol - pos); pos = eol + 2;
            const size_t colon = h.find(':'); if (colon == std::string::npos) continue;
            std::string k = h.substr(0, colon), v = h.substr(colon + 1);
            for (auto& ch : k) ch = (char)std::tolower((unsigned char)ch);
            while (!v.empty() && (v.front() == ' ' || v.front() == '\t')) v.erase(v.begin());
            while (!v.empty() && (v.back() == ' ' || v.back() == '\t')) v.pop_back();
            if (k == "content-length") { char* endp = nullptr; clen = std::strtoll(v.c_str(), &endp, 10); if (endp == v.c_str() || *endp || clen < 0) clen = -1; }
            else if (k == "authorization") rq.auth = v;
            else if (k == "connection") { for (auto& ch : v) ch = (char)std::tolower((unsigned char)ch); if (v == "close") rq.keep_alive = false; }
            else if (k == "transfer-encoding") { respond(c, 400, error_body(400, "chunked request bodies are not supported"), false); return false; }
        }
        if (clen < 0 || (size_t)clen > max_body_) { respond(c, clen < 0 ? 400 : 413, error_body(clen < 0 ? 400 : 413, "bad Content-Length"), false); return false; }
        buf.erase(0, hdr_end + 4);
        while (buf.size() < (size_t)clen) if (!fill(c, buf)) return false;
        rq.body = buf.substr(0, (size_t)clen); buf.erase(0, (size_t)clen);      // always drained: early 401 / 404 replies keep the connection in sync
        return true;
    }
    bool fill(int c, std::string& buf) {
        int idle_ms = 0;
        while (!stop_.load()) {
            pollfd p{c, POLLIN, 0};
            const int r = poll(&p, 1, 200);
            if (r < 0) return false;
            if (r == 0) { if ((idle_ms += 200) >= idle_timeout_ms_) return false; continue; }      // silent for too long (idle keep-alive or a stalled sender): drop it
            char tmp[16384];
            const ssize_t n = recv(c, tmp, sizeof tmp, 0);
            if (n <= 0) return false;
            buf.append(tmp, (size_t)n); return true;
        }
        return false;
    }
    static const char* reason(int s) { switch (s) { case 200: return "OK"; case 400: return "Bad Request"; case 401: return "Unauthorized"; case 404: return "Not Found"; case 413: return "Payload Too Large"; case 429: return "Too Many Requests"; default: return "Internal Server Error"; } }
    static void respond(int c, int status, const std::string& body, bool keep_alive) {
        std::string h = "HTTP/1.1 " + std::to_string(status) + " " + reason(status) + "\r\nContent-Type: application/json\r\nContent-Length: " + std::to_string(body.size()) +
                        (keep_alive ? "\r\n\r\n" : "\r\nConnection: close\r\n\r\n");
        h += body;
        size_t off = 0;
        while (off < h.size()) { const ssize_t n = send(c, h.data() + off, h.size() - off, MSG_NOSIGNAL); if (n <= 0) return; off += (size_t)n; }
    }
    static std::string error_body(int status, const std::string& msg) {
        const char* type = status == 400 ? "invalid_request_error" : status == 401 ? "authentication_error" : status == 429 ? "rate_limit_error" : "server_error";
        return "{\"error\": {\"message\": " + jstr(msg) + ", \"type\": \"" + type + "\", \"code\": " + std::to_string(status) + "}}";
    }
    // the reference's PerfStats shape (pkg/utils/perf.go:296-320 as marshalled by pkg/handlers/perf.go:12-25): summed durations in ns + call counts
    // per operation, and the last reset time (RFC 3339, UTC)
    void perf_record(const char* op, long long ns) { std::lock_guard<std::mutex> lk(perf_mu_); perf_ns_[op] += ns; perf_n_[op] += 1; }
    std::string perf_json() {
        std::lock_guard<std::mutex> lk(perf_mu_);
        std::string t = "\"timers\": {", n = "\"callCounts\": {";
        bool first = true;
        for (auto& kv : perf_ns_) { t += (first ? "" : ", ") + jstr(kv.first) + ": " + std::to_string(kv.second); n += (first ? "" : ", ") + jstr(kv.first) + ": " + std::to_string(perf_n_[kv.first]); first = false; }
        const auto us = std::chrono::duration_cast<std::chrono::microseconds>(perf_reset_.time_since_epoch()).count();
        const time_t secs = (time_t)(us / 1000000); struct tm tmv; gmtime_r(&secs, &tmv);
        char ts[64]; std::snprintf(ts, sizeof ts, "%04d-%02d-%02dT%02d:%02d:%02d.%06lldZ", tmv.tm_year + 1900, tmv.tm_mon + 1, tmv.tm_mday, tmv.tm_hour, tmv.tm_min, tmv.tm_sec, (long long)(us % 1000000));
        return t + "}, " + n + "}, \"lastResetTime\": \"" + ts + "\"";
    }
    // oa_engine_stats summed over the replicas (one engine: its own counters) — `stats.engine`, as http_front.py serves it
    std::string engine_totals_json() {
        std::vector<std::pair<std::string, double>> tot;
        for (oa_engine* e : engines_) {
            char buf[4096]; buf[0] = 0; Json j; std::string err;
            if (oa_engine_stats(e, buf, sizeof buf) != 0 || !parse_json(buf, j, err) || j.t != Json::Obj) continue;
            for (auto& kv : j.o) {
                if (kv.second.t != Json::Num) continue;
                auto it = std::find_if(tot.begin(), tot.end(), [&](const std::pair<std::string, double>& p) { return p.first == kv.first; });
                if (it == tot.end()) tot.emplace_back(kv.first, kv.second.n); else it->second += kv.second.n;
            }
        }
        std::string o = "{";
        for (size_t i = 0; i < tot.size(); ++i) { Json n; n.t = Json::Num; n.n = tot[i].second; o += (i ? ", " : "") + jstr(tot[i].first) + ": "; json_dump(n, o); }
        return o + "}";
    }
    bool authorised(const Request& rq) const {
        if (!require_key_) return true;
        if (rq.auth.rfind("Bearer ", 0) != 0 || rq.auth.size() <= 7) return false;          // the reference always sends its apiKey (openai.go:44)
        return api_key_.empty() || rq.auth.substr(7) == api_key_;
    }
    void serve_connection(int c) {
        std::string buf;
        while (!stop_.load()) {
            Request rq; rq.fd = c;
            if (!read_request(c, buf, rq)) return;
            n_req_.fetch_add(1);
            int status = 200; std::string body;
            try { handle(rq, status, body); }
            catch (const std::exception& e) { status = 500; body = error_body(500, std::string("internal error: ") + e.what()); }
            if (status == 499) return;                      // the client hung up while its completion was running: nobody to answer
            respond(c, status, body, rq.keep_alive);
            if (!rq.keep_alive) return;
        }
    }
    void handle(const Request& rq, int& status, std::string& body) {
        auto ends = [&](const char* suf) { const size_t n = std::strlen(suf); return rq.path.size() >= n && rq.path.compare(rq.path.size() - n, n, suf) == 0; };
        if (rq.method == "GET" && ends("/models")) { body = "{\"object\": \"list\", \"data\": [{\"id\": " + jstr(model_) + ", \"object\": \"model\", \"owned_by\": \"opsagent_b200\"}]}"; return; }
        if (rq.method == "GET" && ends("/perf/stats")) {
            if (!authorised(rq)) { status = 401; body = error_body(401, "missing bearer token"); return; }
            body = "{\"stats\": {" + perf_json() + ", \"engine\": " + engine_totals_json() + ", \"front\": " + stats_json() + "}, \"status\": \"success\"}"; return;
        }
        if (rq.method == "POST" && ends("/perf/reset")) {          // pkg/api/router.go:105, pkg/handlers/perf.go:28-39
            if (!authorised(rq)) { status = 401; body = error_body(401, "missing bearer token"); return; }
            { std::lock_guard<std::mutex> lk(perf_mu_); perf_ns_.clear(); perf_n_.clear(); perf_reset_ = std::chrono::system_clock::now(); }
            body = "{\"message\": \"performance statistics reset\", \"status\": \"success\"}"; return;
        }
        if (rq.method != "POST" || !ends("/chat/completions")) { status = 404; body = error_body(404, "not found"); return; }
        if (!authorised(rq)) { status = 401; body = error_body(401, "missing bearer token"); return; }
        chat(rq, status, body);
    }
    // sticky least-loaded replica choice; -1 = the chosen replica is over its in-flight limit
    int acquire(const std::vector<std::pair<std::string, std::string>>& msgs) {
        uint64_t h = 1469598103934665603ull;
        for (size_t i = 0; i < msgs.size() && i < 2; ++i)
            for (const std::string* s : {&msgs[i].first, &msgs[i].second}) { for (unsigned char ch : *s) { h ^= ch; h *= 1099511628211ull; } h ^= 0xff; h *= 1099511628211ull; }
        std::lock_guard<std::mutex> lk(mu_);
        int r;
        auto it = home_.find(h);
        if (it != home_.end()) { r = it->second; n_sticky_.fetch_add(1); }
        else {
            r = 0;
            for (size_t i = 1; i < engines_.size(); ++i)
                if (inflight_[i].load() < inflight_[(size_t)r].load() || (inflight_[i].load() == inflight_[(size_t)r].load() && routed_[i].load() < routed_[(size_t)r].load())) r = (int)i;
            if (home_.size() >= (1u << 16)) home_.clear();          // bounded; losing stickiness only costs a prefix-cache miss
            home_[h] = r;
        }
        if (inflight_[(size_t)r].load() >= max_inflight_) { n_429_.fetch_add(1); return -1; }
        inflight_[(size_t)r].fetch_add(1); routed_[(size_t)r].fetch_add(1);
        return r;
    }
    void chat(const Request& rq, int& status, std::string& body) {
        Json req; std::string err;
        if (!parse_json(rq.body.empty() ? std::string("{}") : rq.body, req, err) || req.t != Json::Obj) { status = 400; body = error_body(400, "bad request: " + (err.empty() ? "not a JSON object" : err)); return; }
        const Json* jm = req.get("messages");
        if (!jm || jm->t != Json::Arr) { status = 400; body = error_body(400, "bad request: 'messages'"); return; }
        std::vector<std::pair<std::string, std::string>> msgs;
        int n_tool_results = 0;
        for (const Json& m : jm->a) {
            if (m.t != Json::Obj) { status = 400; body = error_body(400, "bad request: message is not an object"); return; }
            const std::string role = m.str("role");
            if (role == "tool") ++n_tool_results;
            const Json* tc = m.get("tool_calls");
            if (tc && tc->t == Json::Arr && !tc->a.empty()) {       // function-calling turns are flattened into text the chat template can carry
                const Json* f = tc->a[0].get("function");
                std::string args = f ? f->str("arguments", "{}") : "{}";
                if (args.empty()) args = "{}";
                msgs.emplace_back(role, "{\"name\":" + jstr(f ? f->str("name") : "") + ",\"arguments\":" + args + "}");
            } else msgs.emplace_back(role, m.str("content"));
        }
        if (const Json* st = req.get("stream")) if (st->t == Json::Bool && st->b) { status = 400; body = error_body(400, "streaming is not implemented (the reference does not request it)"); return; }
        if (const Json* tj = req.get("temperature")) if (tj->t == Json::Num && tj->n > 1e-3) { status = 400; body = error_body(400, "only greedy decoding is implemented"); return; }
        int max_tokens = 1024;
        for (const char* k : {"max_tokens", "max_completion_tokens"}) if (const Json* mt = req.get(k)) if (mt->t == Json::Num && mt->n >= 1) { max_tokens = (int)std::min(mt->n, 1048576.0); break; }      // the engine clips to max_seq_len; the clamp keeps the double -> int conversion defined
        // OpenAI function calling (swarm-go flows, pkg/workflows/swarm.go:14-78): a grammar-forced call of one offered function while fewer than
        // `tool_steps` tool results are in the history, afterwards one line of text
        uint32_t flags = 0; std::string functions;
        if (const Json* tools = req.get("tools")) if (tools->t == Json::Arr && !tools->a.empty()) {
            for (const Json& t : tools->a) {
                const Json* f = t.get("function");
                std::string name = f ? f->str("name", "fn") : "fn", param = "input";
                if (f) if (const Json* pr = f->get("parameters")) if (const Json* props = pr->get("properties")) if (props->t == Json::Obj && !props->o.empty()) param = props->o[0].first;
                functions += (functions.empty() ? "" : ",") + name + ":" + param;
            }
            flags = n_tool_results < tool_steps_ ? OA_FLAG_JSON_FUNCTION : OA_FLAG_JSON_TEXT;
        }
        if (engines_.empty()) { status = 500; body = error_body(500, "no engine behind this front"); return; }
        const int r = acquire(msgs);
        if (r < 0) { status = 429; body = error_body(429, "replica over its in-flight limit"); return; }
        struct Release { std::atomic<int>& n; bool done = false; void now() { if (!done) { n.fetch_sub(1); done = true; } } ~Release() { now(); } } release{inflight_[(size_t)r]};
        const std::string model = req.str("model");
        std::vector<oa_msg> cm(msgs.size());
        for (size_t i = 0; i < msgs.size(); ++i) { cm[i].role = msgs[i].first.c_str(); cm[i].content = msgs[i].second.c_str(); }
        oa_chat_req creq{}; creq.model = model.c_str(); creq.msgs = cm.data(); creq.n_msgs = (int32_t)cm.size(); creq.max_tokens = max_tokens;
        creq.temperature = 1.401298464324817e-45f; creq.flags = flags; creq.functions = (flags & OA_FLAG_JSON_FUNCTION) ? functions.c_str() : nullptr;
        oa_chat_resp out{}; char ebuf[512]; ebuf[0] = 0; uint64_t ticket = 0;
        const auto t_chat = std::chrono::steady_clock::now();
        int rc = oa_chat_submit_ex(engines_[(size_t)r], &creq, &ticket, ebuf, sizeof ebuf);
        // wait in slices: a caller that hangs up (the reference's HTTP client timing out, a cancelled context) or a front that is shutting down gives
        // its KV pages and decode slot back instead of generating for nobody
        while (rc == 0) {
            rc = oa_chat_wait_ex(engines_[(size_t)r], ticket, 250, &out, ebuf, sizeof ebuf);
            if (rc != OA_ERR_TIMEOUT) break;
            rc = 0;
            pollfd p{rq.fd, POLLRDHUP, 0};
            const bool gone = rq.fd >= 0 && poll(&p, 1, 0) > 0 && (p.revents & (POLLRDHUP | POLLHUP | POLLERR));
            if (gone || stop_.load()) {
                oa_chat_cancel(engines_[(size_t)r], ticket); n_cancelled_.fetch_add(1);
                if (gone) { status = 499; return; }
                status = 500; body = error_body(500, "the front is shutting down"); return;
            }
        }
        release.now();
        perf_record((flags & OA_FLAG_JSON_FUNCTION) ? "chat_completion_tool_call" : "chat_completion", std::chrono::duration_cast<std::chrono::nanoseconds>(std::chrono::steady_clock::now() - t_chat).count());
        if (rc != 0) { status = (rc == 400 || rc == 401 || rc == 429 || rc == 500) ? rc : 500; body = error_body(status, ebuf); return; }
        n_chat_.fetch_add(1);
        const std::string content(out.content ? out.content : "", (size_t)out.content_len);
        const int pt = out.prompt_tokens, ct = out.completion_tokens, fr = out.finish_reason;
        oa_free_resp(&out);
        const long long now_us = std::chrono::duration_cast<std::chrono::microseconds>(std::chrono::system_clock::now().time_since_epoch()).count();
        char idb[48]; std::snprintf(idb, sizeof idb, "%llx", (unsigned long long)(now_us ^ ((long long)n_chat_.load() << 20)));
        const std::string usage = "\"usage\": {\"prompt_tokens\": " + std::to_string(pt) + ", \"completion_tokens\": " + std::to_string(ct) + ", \"total_tokens\": " + std::to_string(pt + ct) + "}";
        const std::string head = "{\"id\": \"chatcmpl-" + std::string(idb) + "\", \"object\": \"chat.completion\", \"created\": " + std::to_string(now_us / 1000000) + ", \"model\": " + jstr(model) + ", \"choices\": [{\"index\": 0, \"message\": ";
        if (flags & OA_FLAG_JSON_FUNCTION) {
            Json call; std::string e2;
            const Json *nm = nullptr, *ar = nullptr;
            if (parse_json(content, call, e2)) { nm = call.get("name"); ar = call.get("arguments"); }
            if (!nm || nm->t != Json::Str || !ar) {      // cut off by max_tokens / the context budget: nothing parseable to return; not retried by the caller
                status = 400; body = error_body(400, "function call truncated after " + std::to_string(ct) + " tokens (finish_reason=" + (fr == 0 ? "stop" : "length") + "): raise max_tokens or shorten the history"); return;
            }
            std::string args; json_dump(*ar, args);
            body = head + "{\"role\": \"assistant\", \"content\": null, \"tool_calls\": [{\"id\": \"call_" + std::string(idb) + "\", \"type\": \"function\", \"function\": {\"name\": " + jstr(nm->s) +
                   ", \"arguments\": " + jstr(args) + "}}]}, \"finish_reason\": \"tool_calls\"}], " + usage + "}";
            return;
        }
        std::string text = content;
        while (!text.empty() && text.back() == '\n') text.pop_back();
        body = head + "{\"role\": \"assistant\", \"content\": " + jstr(text) + "}, \"finish_reason\": \"" + (fr == 0 ? "stop" : "length") + "\"}], " + usage + "}";
    }

    std::vector<oa_engine*> engines_;
    std::vector<std::atomic<int>> inflight_; std::vector<std::atomic<long long>> routed_;
    std::mutex mu_; std::unordered_map<uint64_t, int> home_;
    std::mutex perf_mu_; std::map<std::string, long long> perf_ns_, perf_n_; std::chrono::system_clock::time_point perf_reset_ = std::chrono::system_clock::now();
    bool require_key_ = true; std::string api_key_, model_; int tool_steps_ = 3, max_inflight_ = 256, max_conn_ = 8192, idle_timeout_ms_ = 120000; size_t max_body_ = 64 << 20;
    int lsock_ = -1, port_ = 0; std::thread acceptor_; std::atomic<bool> stop_{false};
    std::atomic<int> n_conn_{0}; std::atomic<long long> n_req_{0}, n_chat_{0}, n_429_{0}, n_sticky_{0}, n_cancelled_{0};
};

}  // namespace oa

struct oa_http { std::unique_ptr<oa::HttpFront> f; };
static thread_local std::string g_http_err;

extern "C" {
int oa_http_start(oa_engine* const* engines, int32_t n_engines, const char* options_json, oa_http** out) {
    if (!out || n_engines < 0 || (n_engines > 0 && !engines)) return OA_ERR_BAD_REQUEST;
    *out = nullptr;
    try {
        const oa::JsonFlat opt = oa::JsonFlat::parse(options_json && options_json[0] ? options_json : "{}");
        auto h = new oa_http; h->f.reset(new oa::HttpFront(std::vector<oa_engine*>(engines, engines + n_engines), opt)); *out = h;
    } catch (const std::exception& e) { g_http_err = e.what(); return OA_ERR_INTERNAL; }
    return OA_OK;
}
int32_t oa_http_port(oa_http* h) { return h ? h->f->port() : -1; }
int oa_http_stats(oa_http* h, char* buf, size_t n) { if (!h || !buf || !n) return OA_ERR_BAD_REQUEST; std::snprintf(buf, n, "%s", h->f->stats_json().c_str()); return OA_OK; }
void oa_http_stop(oa_http* h) { if (!h) return; if (!h->f->stop()) (void)h->f.release(); /* leak rather than free under a live thread */ delete h; }
const char* oa_http_last_error(void) { return g_http_err.c_str(); }
// test hook: parse `in` with the front's JSON reader and write it back with its writer (tests compare both with Python's json on CPU)
int oa_host_json_roundtrip(const uint8_t* in, size_t n_in, char* out, size_t n_out) {
    if (!in || !out || !n_out) return OA_ERR_BAD_REQUEST;
    oa::Json j; std::string err, o;
    if (!oa::parse_json(std::string((const char*)in, n_in), j, err)) { std::snprintf(out, n_out, "%s", err.c_str()); return OA_ERR_BAD_REQUEST; }
    oa::json_dump(j, o);
    if (o.size() + 1 > n_out) return OA_ERR_BAD_REQUEST;
    std::memcpy(out, o.data(), o.size()); out[o.size()] = 0;
    return OA_OK;
}
}
