// ASAN + UBSAN pass over the host-only pieces of the engine that parse UNTRUSTED bytes: request text -> BPE, tokenizer.json and
// config JSON readers, the safetensors header reader, the ToolPrompt byte automaton and its token masks.  Seeded mutations of valid inputs; every
// input must end in a result or a C++ exception, never in a sanitizer report.  argv: <tokenizer.json> <scratch dir> [iterations]
#include <cstdio>
#include <cstring>
#include <fstream>
#include <random>
#include <sstream>
#include <stdexcept>
#include <string>
#include <vector>

#include "grammar.hpp"
#include "token_mask.hpp"
#include "bpe.hpp"
#include "config.hpp"
#include "tokenizer.hpp"
#include "safetensors.hpp"

using namespace oa;

static std::string mutate(const std::string& s, std::mt19937& rng) {
    std::string o = s;
    const int n = 1 + (int)(rng() % 4);
    for (int k = 0; k < n && !o.empty(); ++k) {
        const size_t p = rng() % o.size();
        switch (rng() % 5) {
            case 0: o[p] = (char)(rng() & 255); break;
            case 1: o.erase(p, 1 + rng() % 8); break;
            case 2: o.insert(p, 1 + rng() % 4, "{}[]\",:\\u\xff\xc0\xed\xa0"[rng() % 13]); break;
            case 3: o.resize(p); break;
            default: o.insert(p, o.substr(rng() % o.size(), rng() % 32));
        }
    }
    return o;
}

int main(int argc, char** argv) {
    if (argc < 3) { std::fprintf(stderr, "usage: host_fuzz tokenizer.json scratch_dir [iters]\n"); return 2; }
    const int iters = argc > 3 ? std::atoi(argv[3]) : 300;
    std::mt19937 rng(20260921u);
    long ok = 0, thrown = 0;
    std::ifstream f(argv[1], std::ios::binary); std::stringstream ss; ss << f.rdbuf();
    const std::string tok_json = ss.str();
    auto bpe = BpeTokenizer::from_json(tok_json);

    // 1. BPE on arbitrary bytes: encode must accept anything; decoding what it produced gives the bytes back
    for (int it = 0; it < iters; ++it) {
        std::string text;
        const int len = (int)(rng() % 300);
        for (int i = 0; i < len; ++i) {
            switch (rng() % 6) {
                case 0: text += (char)(rng() & 255); break;
                case 1: text += "\xe4\xb8\xad"; break;
                case 2: text += "\xf0\x9f\x98\x80"; break;
                case 3: text += " \n\t  "[rng() % 6]; break;
                default: text += (char)(' ' + rng() % 95);
            }
        }
        std::vector<int32_t> ids; bpe->encode(text, ids);
        const std::string back = bpe->decode(ids.data(), ids.size());
        if (back != text) { std::fprintf(stderr, "BPE round trip differs (len %zu vs %zu)\n", back.size(), text.size()); return 1; }
        ++ok;
    }
    // 2. tokenizer.json reader on damaged files (only a short prefix region and a few random regions are mutated: the file is ~600 KB)
    for (int it = 0; it < iters / 10; ++it) {
        try { auto t = BpeTokenizer::from_json(mutate(tok_json, rng)); std::vector<int32_t> ids; t->encode("kubectl get pods -n default", ids); ++ok; } catch (const std::exception&) { ++thrown; }
    }
    // 3. flat config reader + model/engine option resolution
    const std::string cfg = "{\"model\": \"custom\", \"hidden\": 64, \"n_layers\": 2, \"n_heads\": 4, \"n_kv_heads\": 2, \"head_dim\": 64, \"ffn\": 128, \"vocab\": 512, \"max_batch\": 8, \"kv_gb\": 0.5, \"tp\": 1, \"model_aliases\": \"a,b,*\", \"rope_theta\": 5e5}";
    for (int it = 0; it < iters * 4; ++it) {
        try { ModelConfig m; EngineOptions o; parse_config(it == 0 ? cfg : mutate(cfg, rng), m, o); ++ok; } catch (const std::exception&) { ++thrown; }
    }
    // 4. safetensors header reader on damaged files
    {
        const std::string hdr = "{\"w\": {\"dtype\": \"BF16\", \"shape\": [4, 8], \"data_offsets\": [0, 64]}, \"b\": {\"dtype\": \"F32\", \"shape\": [8], \"data_offsets\": [64, 96]}, \"__metadata__\": {\"format\": \"pt\"}}";
        for (int it = 0; it < iters; ++it) {
            std::string h = it == 0 ? hdr : mutate(hdr, rng);
            uint64_t n = h.size();
            if (it % 7 == 3) n = rng();                       // a header length that lies
            if (it % 11 == 5) n = ~0ull - (rng() % 16);
            std::string file((const char*)&n, 8); file += h; file.append(it % 5 == 0 ? rng() % 96 : 96, '\x01');
            if (it % 13 == 7) file.resize(rng() % 8);          // shorter than the length field
            const std::string path = std::string(argv[2]) + "/fuzz.safetensors";
            { std::ofstream o(path, std::ios::binary); o.write(file.data(), (std::streamsize)file.size()); }
            try {
                SafeTensors st(path);
                if (st.has("w")) { const StTensor& t = st.get("w"); volatile uint8_t sink = 0; for (size_t i = 0; i < t.bytes; ++i) sink = sink ^ t.data[i]; (void)sink; }
                ++ok;
            } catch (const std::exception&) { ++thrown; }
        }
    }
    // 5. grammar automaton + token masks: random legal walks, every masked-in token must be steppable byte by byte, illegal bytes are refused
    {
        TokenTrie trie; const std::vector<std::string> tb = bpe->text_token_bytes(); trie.build(tb, (int)tb.size());
        std::vector<uint32_t> mask((size_t)trie.words());
        const char* fns[] = {"kubectl:command,trivy:image", "a:b", "", ":::,,,", "python:code,jq:filter,x:y"};
        for (int it = 0; it < iters / 3; ++it) {
            const int kind = 1 + (int)(rng() % 4);       // TOOLCALL / FINAL / FUNCTION / TEXT
            ToolPromptGrammar g(kind, fns[rng() % 5]);
            if (!g.active()) { ++ok; continue; }
            auto c = g.start();
            for (int step = 0; step < 600 && !g.done_at(c); ++step) {
                trie.allowed_tokens(g, c, mask.data());
                std::vector<int> allowed;
                for (int t = 0; t < (int)tb.size(); ++t) if (mask[(size_t)t >> 5] >> (t & 31) & 1u) allowed.push_back(t);
                if (allowed.empty()) { std::fprintf(stderr, "grammar state with no allowed token (kind %d, key %s)\n", kind, g.state_key(c).c_str()); return 1; }
                (void)grammar_mask_key(g, c);
                const int t = allowed[rng() % allowed.size()];
                for (unsigned char b : tb[(size_t)t]) if (!g.step(c, b)) { std::fprintf(stderr, "masked-in token %d is not steppable\n", t); return 1; }
                auto bad = c; uint32_t bytes[8]; g.allowed_at(c, bytes);
                const int b = (int)(rng() & 255);
                if (!g.done_at(c) && !(bytes[b >> 5] >> (b & 31) & 1u) && g.step(bad, b)) { std::fprintf(stderr, "byte %d accepted though masked out\n", b); return 1; }
            }
            ++ok;
        }
    }
    // 6. chat template over both templates with hostile roles / contents
    for (int it = 0; it < iters / 3; ++it) {
        ModelConfig m; EngineOptions o; parse_config(cfg, m, o); m.chat_template = it % 2 ? "llama3" : "chatml";
        Tokenizer tk(m);
        std::vector<ChatMessage> msgs;
        for (int k = (int)(rng() % 4); k >= 0; --k) msgs.push_back({mutate("user", rng), mutate("why is pod web-0 crashing? \xe4\xb8\xad", rng)});
        auto ids = tk.apply_chat_template(msgs);
        for (int32_t id : ids) if (id < 0 || id >= m.vocab) { std::fprintf(stderr, "template id out of range\n"); return 1; }
        (void)tk.detokenize(ids); ++ok;
    }
    std::printf("ok: %ld inputs handled, %ld refused with an exception\n", ok, thrown);
    return 0;
}
